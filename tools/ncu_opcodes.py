#!/usr/bin/env python
"""Executed-instruction histogram by opcode of one kernel in an .ncu-rep (needs --import-source on / --set full).
Usage: python tools/ncu_opcodes.py file.ncu-rep [top]"""
import csv
import subprocess
import sys
from collections import Counter


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = next(r for r in rows if 'Instructions Executed' in r)
    data = rows[rows.index(hdr) + 1:]
    ia, isrc, isamp = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('# Samples')
    c, samp, tot, bar = Counter(), Counter(), 0, 0
    for r in data:
        try:
            n = int(r[ia])
        except (ValueError, IndexError):
            continue
        op = r[isrc].strip().split()
        if not op:
            continue
        o = op[1] if op[0].startswith('@') else op[0]
        o = o.split('.')[0]
        c[o] += n; tot += n; samp[o] += int(r[isamp])
        if o == 'BAR':
            bar += n
    print(f"total warp instructions {tot}; BAR executions {bar}; instructions per BAR {tot / max(bar, 1):.1f}")
    for o, n in c.most_common(top):
        print(f"{o:10s} {n:14d} {100 * n / tot:5.1f}%  per-BAR {n / max(bar, 1):6.1f}  samples {samp[o]}")


if __name__ == '__main__':
    main()
