#!/usr/bin/env python
"""One-shot GPU probe: fp64 / copy peaks, and per-kernel device times of one source iteration for
a list of (workload, layout) pairs.  Usage: python tools/gpu_probe.py [workload:layout ...]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kripke_b200 as kb  # noqa: E402
from bench import WORKLOADS, kripke_args  # noqa: E402


def main():
    A, H = kb.abi(), kb.host()
    kb.init_device(0)
    H.kripke_b200_timer_sync(0)
    out = {}
    g = C.c_double()
    A.kb200_peak_fp64_gflops(0, 4000, C.byref(g)); out["dfma_gflops"] = g.value
    A.kb200_peak_fp64_gflops(1, 4000, C.byref(g)); out["dmma_gflops"] = g.value
    A.kb200_peak_copy_gbs(4 << 30, 5, C.byref(g)); out["copy_gbs"] = g.value
    # write-only stream (kb200_fill_f64 over 8 GB): what a store-dominated kernel such as LPlusTimes can hope for
    buf, e0, e1, ms = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_float()
    nfill = 1 << 30
    if A.kb200_alloc(nfill * 8, C.byref(buf)) == 0:
        A.kb200_event_create(C.byref(e0)); A.kb200_event_create(C.byref(e1))
        best = 1e30
        for _ in range(4):
            A.kb200_event_record(e0, None)
            A.kb200_fill_f64(buf, 1.0, nfill, None)
            A.kb200_event_record(e1, None)
            A.kb200_event_sync(e1)
            A.kb200_event_elapsed_ms(e0, e1, C.byref(ms))
            best = min(best, ms.value)
        out["fill_gbs"] = nfill * 8 / (best * 1e-3) * 1e-9
        A.kb200_free(buf)
    print(json.dumps(out), flush=True)
    pairs = sys.argv[1:] or ["small:DGZ", "small:GZD", "small:ZGD"]
    kernels = ["LTimes", "scattering", "source", "LPlusTimes", "SweepSolver", "population"]
    for pr in pairs:
        wl, lay = pr.split(":")
        kargs, unknowns = kripke_args(wl, 1, lay, 1)
        p = kb.Problem(kargs)
        G, D, L = WORKLOADS[wl][1], WORKLOADS[wl][2], WORKLOADS[wl][3]
        M = (L + 1) ** 2
        N_u = float(unknowns); N_m = N_u / D * M
        alg = {"LTimes": 8 * N_u + 8 * N_m, "LPlusTimes": 8 * N_u + 8 * N_m, "scattering": 16 * N_m, "SweepSolver": 16 * N_u,
               "population": 8 * N_u, "source": 1.0}
        evs = {k: (C.c_void_p(), C.c_void_p()) for k in kernels}
        for a, b in evs.values():
            A.kb200_event_create(C.byref(a)); A.kb200_event_create(C.byref(b))
        p.call("zero:psi")
        res = {}
        for it in range(3):
            for z, k in (("phi", "LTimes"), ("phi_out", "scattering"), (None, "source"), ("rhs", "LPlusTimes"),
                         (None, "SweepSolver"), (None, "population")):
                if z:
                    p.call("zero:" + z)
                A.kb200_event_record(evs[k][0], None)
                r = p.call(k)
                A.kb200_event_record(evs[k][1], None)
            A.kb200_device_sync()
            for k in kernels:
                ms = C.c_float(); A.kb200_event_elapsed_ms(evs[k][0], evs[k][1], C.byref(ms))
                res[k] = ms.value
        total = sum(res.values())
        print(f"{pr}: unknowns={unknowns:.3e} total={total:.3f} ms grind={1e6*total/unknowns:.4f} ns  particles={r:.10e}")
        for k in kernels:
            print(f"   {k:12s} {res[k]:9.3f} ms  {alg[k]/res[k]/1e6:9.1f} GB/s(alg)")
        sys.stdout.flush()
        p.close()


if __name__ == "__main__":
    main()
