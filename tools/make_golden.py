#!/usr/bin/env python
"""Generates tests/golden/reference_goldens.json by running the UNMODIFIED reference
(oracle/_ref/kripke_ref, built from /root/reference by `make -C oracle ref`) with --arch Sequential.

For every case it stores the command line, the per-iteration particle counts (17 significant
digits) and the L2 norm / sum of phi, phi_out, rhs, psi after the last iteration.  The reference
itself carries no numerical fixtures, so these outputs of the reference ARE the
golden vectors that pin both the oracle and the CUDA path.  Re-run only where /root/reference exists.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "kripke_ref")

CASES = {
    # BASELINE config 1 (reference default problem), SURVEY G1
    "G1_default": "--zones 16,16,16 --groups 32 --quad 96 --legendre 4 --layout DGZ --niter 10",
    "G1z_zset222": "--zones 16,16,16 --groups 32 --quad 96 --legendre 4 --layout DGZ --niter 10 --zset 2,2,2",
    # config-3 proxy (high scattering order), SURVEY G3
    "G3_legendre9": "--zones 8,8,8 --groups 16 --quad 16 --legendre 9 --niter 5",
    # config-4 proxy: procs 2,2,2 x zset 2,2,2 == single rank zset 4,4,4
    "G4_kba_proxy": "--zones 16,16,16 --zset 4,4,4 --gset 4 --dset 8 --groups 16 --quad 48 --niter 3",
    "G4b_undecomposed": "--zones 16,16,16 --zset 1,1,1 --gset 1 --dset 8 --groups 16 --quad 48 --niter 3",
    # config-5 proxy (block Jacobi); plane chunks big enough to come from fresh zero pages
    "G5_block_jacobi": "--pmethod bj --zones 32,32,32 --zset 2,2,2 --groups 32 --quad 96 --niter 3",
    # Gauss-Legendre quadrature, <=16 and >16 directions (std::sort paths), SURVEY G6
    "G6_gauss_legendre_4x4": "--quad 4:4 --zones 8,8,8 --groups 4 --legendre 2 --niter 3",
    "G6b_gauss_legendre_8x8": "--quad 8:8 --zones 8,8,8 --groups 4 --legendre 3 --niter 3",
    # ragged: non-cubic zones, odd extents, one group set
    "R1_ragged": "--zones 12,6,10 --groups 6 --quad 24 --legendre 1 --gset 3 --zset 3,1,2 --niter 3",
    "R2_legendre0": "--zones 6,6,6 --groups 2 --quad 8 --legendre 0 --gset 1 --niter 2",
    "R3_custom_xs": "--zones 8,8,8 --groups 4 --quad 16 --legendre 2 --sigt 0.2,0.001,0.3 --sigs 0.1,0.0005,0.02 --niter 4",
}
# BASELINE config 3 at full size (15 GB of fields, minutes on the host) and the largest slab of BASELINE config 2
# this container's 62 GB host can hold (config 2's groups, directions and 64 x 64 zone cross-section, half its k extent)
BIG_CASES = {
    "G7_config3_full": "--zones 32,32,32 --groups 128 --quad 128 --legendre 9 --niter 2",
    "G8_config2_half_slab": "--zones 64,64,32 --groups 64 --quad 192 --legendre 4 --niter 2",
}
# every storage order on one decomposed problem
for lay in ["DGZ", "DZG", "GDZ", "GZD", "ZDG", "ZGD"]:
    CASES[f"L_{lay}"] = f"--zones 12,8,8 --groups 8 --quad 32 --legendre 3 --gset 2 --dset 8 --zset 2,1,2 --layout {lay} --niter 3"


def run_case(args):
    out = subprocess.check_output([REF, "--arch", "Sequential"] + args.split(), text=True)
    particles, norms = [], {}
    for line in out.splitlines():
        if line.startswith("ITER"):
            particles.append(float(line.split("particles=")[1]))
        elif line.startswith("NORM"):
            _, name, n, l2, s = line.split()
            norms[name] = {"n": int(n[2:]), "l2": float(l2[3:]), "sum": float(s[4:])}
        elif "Material Volumes" in line:
            norms["material_volumes_line"] = line.strip()
    return {"args": args, "particles": particles, "norms": norms}


def main():
    """no arguments: regenerate the small cases; `make_golden.py NAME...`: (re)generate only the named cases (the big ones
    are only run by name) and merge them into the existing file."""
    if not os.path.exists(REF):
        sys.exit("build the reference first: make -C oracle ref")
    path = os.path.join(ROOT, "tests", "golden", "reference_goldens.json")
    allc = dict(CASES, **BIG_CASES)
    if len(sys.argv) > 1:
        with open(path) as f:
            out = json.load(f)
        for name in sys.argv[1:]:
            out[name] = run_case(allc[name])
    else:
        out = {}
        if os.path.exists(path):
            with open(path) as f:
                out = {k: v for k, v in json.load(f).items() if k in BIG_CASES}
        out.update({name: run_case(args) for name, args in CASES.items()})
    for name, r in out.items():
        print(name, "%.17g" % r["particles"][-1])
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
