#!/usr/bin/env python
"""bench.py -- Kripke source-iteration hot path on B200: grind time per unknown per iteration.

  python bench.py --gpus 1 --steps K --warmup W              our arm (CUDA, sm_100a)
  python bench.py --impl reference --gpus N --steps K ...     the unmodified reference on the host CPUs
  torchrun --nproc-per-node N ... bench.py --gpus N ...       one rank per GPU (NCCL over NVLink)

A "step" is one source iteration (LTimes, Scattering, Source, LPlusTimes, SweepSolver, Population)
of BASELINE.json configs[1] per GPU: 64^3 zones x 64 groups x 192 directions, Legendre order 4,
pmethod sweep (3.22e9 unknowns, 59.5 GB of fields: far larger than the 126 MB L2, so no flush is
needed between iterations).  For N>1 the same per-GPU block is tiled with --procs (weak scaling,
KBA sweep with face exchange over NCCL).  Rank 0 prints ONE JSON line.
"""
import argparse
import ctypes as C
import json
import os
import re
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PROCS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
WORKLOADS = {
    # name: (zones per GPU, groups, directions, legendre, extra flags)
    "config2": ((64, 64, 64), 64, 192, 4, ""),                 # BASELINE configs[1] (headline)
    "config1": ((16, 16, 16), 32, 96, 4, ""),                  # configs[0], the reference's default
    "config3": ((32, 32, 32), 128, 128, 9, ""),                # configs[2], high scattering order
    "config4": ((64, 64, 64), 64, 96, 4, "--zset 4,4,4 --gset 4 --dset 8"),  # configs[3], per-GPU share of the 8-GPU KBA sweep (16^3 subdomains)
    "config5": ((64, 64, 64), 32, 96, 4, "--pmethod bj"),      # configs[4], block Jacobi weak scaling
    "small": ((32, 32, 32), 32, 96, 4, ""),
}


def kripke_args(workload, n_gpus, layout, niter, zones_override=None):
    zones, groups, dirs, leg, extra = WORKLOADS[workload]
    if zones_override:
        zones = zones_override
    px, py, pz = PROCS[n_gpus]
    gz = (zones[0] * px, zones[1] * py, zones[2] * pz)
    a = f"--zones {gz[0]},{gz[1]},{gz[2]} --groups {groups} --quad {dirs} --legendre {leg} --layout {layout} " \
        f"--procs {px},{py},{pz} --niter {niter} {extra}"
    return a.split(), groups * dirs * gz[0] * gz[1] * gz[2]


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks / throttle reasons while the timed region runs"""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.max_mhz, self.reasons = [], None, set()
        self.stop_flag = threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag.is_set():
            try:
                out = subprocess.check_output(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                               "-i", str(self.index)], text=True, timeout=5).strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def result(self):
        self.stop_flag.set()
        self.join(timeout=3)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload, layout, kernel):
    """DRAM bytes (read + write) per call of the entry point (one launch, or the eleven tile-diagonal launches of the pencil
    sweep) from the committed `ncu --set full` capture of the
    same workload (profiles/ncu_traffic.json, written by tools/ncu_traffic.py), and where that number comes from: the
    capture file and the commit it was taken at.  (None, None) if no capture of this workload/layout exists -- a number
    measured under a profiler cannot be produced inside a timing run."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            e = json.load(f)[f"{workload}:{layout}"]
        k = e["per_entry_point"][kernel]
        # the capture holds exactly one source iteration: the sum over the entry point's launches is its traffic per call
        return k["dram_bytes"], {"file": "profiles/ncu_traffic.json", "capture": e.get("source"), "commit": e.get("commit"),
                                 "launches_per_call": k["launches"]}
    except Exception:
        return None, None


def host_mem_available_gb():
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


def reference_sample_zones(workload, steps, warmup, budget_s=240.0):
    """The zone block the reference arm runs: the workload's own per-GPU block if the host can hold it (fields take
    ~18.5 bytes x unknowns plus moments) and steps+warmup iterations fit the time budget at the CPU rate measured by
    earlier runs (~5.4 ns per unknown and iteration on 16 threads), else the largest halving of it that does."""
    zones, groups, dirs, leg, _ = WORKLOADS[workload]
    M = (leg + 1) ** 2
    cores = os.cpu_count() or 1
    rate = 5.4e-9 * 16.0 / max(cores, 1)
    mem = host_mem_available_gb() * 1e9
    z = list(zones)
    k = 2
    while True:
        nz = z[0] * z[1] * z[2]
        bytes_needed = nz * groups * (16.0 * dirs + 16.0 * M + 8.0) * 1.15
        secs = nz * groups * dirs * rate * (steps + warmup)
        if (bytes_needed <= 0.8 * mem and secs <= budget_s) or nz <= 16 ** 3:
            return tuple(z)
        if z[k] > 8:
            z[k] //= 2
        k = (k - 1) % 3


def run_reference_cpu(workload, layout, steps, warmup, zones):
    """times the unmodified reference (oracle/_ref/kripke_ref, OpenMP, all host threads) on a bounded
    sample of the workload: same groups/directions/legendre/layout, fewer zones."""
    ref = os.path.join(ROOT, "oracle", "_ref", "kripke_ref")
    cores = os.cpu_count() or 1
    groups, dirs, leg, extra = WORKLOADS[workload][1:]
    if not os.path.exists(ref):
        return None
    env = dict(os.environ, OMP_NUM_THREADS=str(cores), OMP_PROC_BIND="spread", OMP_PLACES="cores")
    cmd = [ref, "--arch", "OpenMP", "--layout", layout, "--zones", "%d,%d,%d" % zones, "--groups", str(groups),
           "--quad", str(dirs), "--legendre", str(leg), "--niter", str(steps + warmup), "--time"] + extra.split()
    out = subprocess.check_output(cmd, text=True, env=env)
    command = "kripke " + " ".join(cmd[1:])
    times = [float(l.split()[2]) for l in out.splitlines() if l.startswith("ITER_TIME")]
    timers = {l.split()[1]: float(l.split()[2]) for l in out.splitlines() if l.startswith("TIMER ")}
    unknowns = groups * dirs * zones[0] * zones[1] * zones[2]
    t = sum(times[warmup:]) / max(1, len(times[warmup:]))
    return {"value": 1e9 * t / unknowns, "unit": "ns/(unknown*iter)", "cores": cores, "kind": "reference",
            "sample": f"unmodified reference (OpenMP, {cores} threads), zones {zones[0]}x{zones[1]}x{zones[2]} of the "
                      f"workload's {groups} groups x {dirs} directions, {len(times[warmup:])} timed iterations after {warmup} warm-up",
            "s_per_iter": t, "unknowns": unknowns, "kernel_seconds_total": timers, "command": command, "zones": list(zones)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--layout", default=os.environ.get("KB200_BENCH_LAYOUT", "DGZ"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-zones", default="auto", help="zones of the CPU sample for --impl reference (auto: the largest "
                    "halving of the workload's block that fits host memory and a few minutes)")
    args = ap.parse_args()
    warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = args.gpus
    if world > 1 and world != n_gpus:
        sys.exit(f"--gpus {n_gpus} but WORLD_SIZE={world}")

    kargs, unknowns = kripke_args(args.workload, n_gpus, args.layout, args.steps + warmup)
    wz, wg, wd, wl, wx = WORKLOADS[args.workload]
    # the named workload; the command line each arm actually ran is the line's own "command" key
    config = {"workload": f"Kobayashi-3i synthetic, BASELINE {args.workload}: {wz[0]}x{wz[1]}x{wz[2]} zones per GPU, {wg} groups, "
                          f"{wd} directions, legendre {wl}{(' ' + wx) if wx else ''}, one source iteration per step",
              "layout": args.layout, "unknowns": unknowns,
              "l2_policy": "inputs larger than L2 (%.1f GB of fields per GPU)" % (unknowns / n_gpus * (16.0 + 16.0 * (wl + 1) ** 2 / wd) / 1e9),
              "parallelism": f"kba{n_gpus}" if n_gpus > 1 else "single"}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        zones = reference_sample_zones(args.workload, args.steps, warmup) if args.ref_zones == "auto" \
            else tuple(int(x) for x in args.ref_zones.split(","))
        r = run_reference_cpu(args.workload, args.layout, args.steps, warmup, zones)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/kripke_ref was not built"}))
            return
        line = {"impl": "reference", "metric": "grind_time", "value": r["value"], "unit": r["unit"], "n_gpus": n_gpus,
                "steps": args.steps, "warmup": warmup, "ms_per_step": 1e3 * r["s_per_iter"], "higher_is_better": False,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "throughput_unknowns_per_s": 1e9 / r["value"],
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": r["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "command": r["command"], "sample_zones": r["zones"], "sample_unknowns": r["unknowns"],
                "sample_is_full_workload": list(r["zones"]) == list(WORKLOADS[args.workload][0]) and n_gpus == 1,
                "host_mem_available_gb": host_mem_available_gb(),
                "kernel_seconds_total": r["kernel_seconds_total"]}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm
    import numpy as np
    import kripke_b200 as kb
    A, H = kb.abi(), kb.host()
    kb.init_device(local_rank)
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", rank=rank, world_size=world)
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_ubyte * 128)()
            kb.api.check(A.kb200_comm_unique_id(buf), "kb200_comm_unique_id")
            uid = torch.tensor(list(buf), dtype=torch.uint8)
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        raw = bytes(uid.cpu().tolist())
        kb.api.check(A.kb200_comm_init(rank, world, raw), "kb200_comm_init")
        H.kripke_b200_set_world(rank, world)

    def barrier():
        A.kb200_device_sync()
        if dist is not None:
            dist.barrier()

    H.kripke_b200_timer_sync(0)

    # ---- parity of the multi-GPU path on THIS box, before anything is timed (N > 1): the KBA sweep and the block-Jacobi
    # exchange over NCCL must reproduce the particle counts of the unmodified reference for the equivalent single-rank
    # zone-set decomposition (--procs px,py,pz --zset a,b,c == one rank with --zset px*a,py*b,pz*c)
    parity = None
    if world > 1:
        with open(os.path.join(ROOT, "tests", "golden", "reference_goldens.json")) as f:
            gold = json.load(f)
        px, py, pz = PROCS[n_gpus]
        parity = {"rtol": 1e-12, "cases": {}, "max_rel_err": 0.0}
        for name in ("G4_kba_proxy", "G5_block_jacobi", "L_GZD", "L_ZGD"):
            a = gold[name]["args"].split()
            zs = [1, 1, 1]
            if "--zset" in a:
                i = a.index("--zset")
                zs = [int(x) for x in a[i + 1].split(",")]
                del a[i:i + 2]
            if zs[0] % px or zs[1] % py or zs[2] % pz:
                continue
            a += ["--zset", "%d,%d,%d" % (zs[0] // px, zs[1] // py, zs[2] // pz), "--procs", "%d,%d,%d" % (px, py, pz)]
            q = kb.Problem(a)
            got = q.solve()
            q.close()
            ref = gold[name]["particles"]
            err = max(abs(x - y) / abs(y) for x, y in zip(got, ref)) if len(got) == len(ref) else float("inf")
            parity["cases"][name] = {"args": " ".join(a), "max_rel_err": err}
            parity["max_rel_err"] = max(parity["max_rel_err"], err)
            dist.barrier()
        parity["case"] = ",".join(parity["cases"])
        if not parity["cases"] or parity["max_rel_err"] > parity["rtol"]:
            if rank == 0:
                print(json.dumps({"error": "multi-GPU parity check failed, nothing was timed", "parity": parity}))
            sys.exit(3)

    p = kb.Problem(kargs)

    # e2e staging: every generated input table of the step lives in pinned host memory and is
    # re-uploaded inside the timed region (cross sections, material mix, quadrature, L/L+)
    step_inputs = ["sigt_zonal", "data/sigs", "ell", "ell_plus", "quadrature/xcos", "quadrature/ycos", "quadrature/zcos",
                   "quadrature/w", "volume", "dx", "dy", "dz", "zone_to_mixelem", "zone_to_num_mixelem",
                   "mixelem_to_zone", "mixelem_to_material", "mixelem_to_fraction", "moment_to_legendre"]
    staged, h2d_bytes = [], 0
    for name in step_inputs:
        for c in range(p.num_chunks(name)):
            arr = p.chunk(name, c)
            nbytes = arr.nbytes
            if nbytes == 0:
                continue
            hp = C.c_void_p()
            kb.api.check(A.kb200_alloc_host(nbytes, C.byref(hp)))
            C.memmove(hp, arr.ctypes.data, nbytes)
            staged.append((name, c, hp, nbytes))
            h2d_bytes += nbytes

    # Every table is needed by one entry point only.  In the end-to-end region the copies go to a second stream in the
    # order of need, with one event per entry point that the library's stream waits for right before that entry point: the
    # transfer of the one large table, sigt_zonal (134 MB at config 2, read by the sweep at the end of the step), rides
    # under LTimes/scattering/LPlusTimes, and the many small copies no longer sit in front of the first kernel.
    needed_by = {"ell": "LTimes", "data/sigs": "scattering", "zone_to_mixelem": "scattering", "zone_to_num_mixelem": "scattering",
                 "mixelem_to_zone": "scattering", "mixelem_to_material": "scattering", "mixelem_to_fraction": "scattering",
                 "moment_to_legendre": "scattering", "ell_plus": "LPlusTimes", "quadrature/xcos": "SweepSolver",
                 "quadrature/ycos": "SweepSolver", "quadrature/zcos": "SweepSolver", "dx": "SweepSolver", "dy": "SweepSolver",
                 "dz": "SweepSolver", "sigt_zonal": "SweepSolver", "quadrature/w": "population", "volume": "population"}
    order = ["LTimes", "scattering", "LPlusTimes", "SweepSolver", "population"]
    copy_stream = C.c_void_p()
    ev_copy = {}
    for k in order:
        e = C.c_void_p()
        A.kb200_event_create(C.byref(e))
        ev_copy[k] = e

    def upload_inputs(groups=None):
        if not copy_stream.value:
            if groups is None or "LTimes" in groups:  # single-stream mode: everything in front of the first kernel
                for name, c, hp, nbytes in staged:
                    A.kb200_upload(p.device_ptr(name, c, True), hp, nbytes, None)
            return
        for k in (order if groups is None else groups):
            for name, c, hp, nbytes in staged:
                if needed_by[name] == k:
                    A.kb200_upload(p.device_ptr(name, c, True), hp, nbytes, copy_stream)
            A.kb200_event_record(ev_copy[k], copy_stream)

    kernels = ["LTimes", "scattering", "source", "LPlusTimes", "SweepSolver", "population"]
    A.kb200_last_sweep_kernel.restype = C.c_char_p
    A.kb200_last_scattering_kernel.restype = C.c_char_p
    sweep_kernels, scattering_kernels = set(), set()
    ev = {}
    for k in kernels:
        a, b = C.c_void_p(), C.c_void_p()
        A.kb200_event_create(C.byref(a)); A.kb200_event_create(C.byref(b))
        ev[k] = (a, b)
    ev_step = (C.c_void_p(), C.c_void_p())
    A.kb200_event_create(C.byref(ev_step[0])); A.kb200_event_create(C.byref(ev_step[1]))
    ktime = {k: [] for k in kernels}
    particles = []

    def step(timed, with_h2d):
        # the tables of the first entry point go first; the other copies (134 MB of sigt_zonal among them) are queued right
        # after LTimes has been launched, so the GPU is not left idle while the host issues some fifty copy calls
        if with_h2d:
            upload_inputs(["LTimes"])
        for z, k in (("phi", "LTimes"), ("phi_out", "scattering"), (None, "source"), ("rhs", "LPlusTimes"),
                     (None, "SweepSolver"), (None, "population")):
            if z:
                p.call("zero:" + z)
            if timed:
                A.kb200_event_record(ev[k][0], None)
            if with_h2d and copy_stream.value and k in ev_copy:
                A.kb200_stream_wait_event(None, ev_copy[k])
            r = p.call(k)
            if with_h2d and k == "LTimes":
                upload_inputs(order[1:])
            if timed:
                A.kb200_event_record(ev[k][1], None)
            if k == "SweepSolver":
                sweep_kernels.add(A.kb200_last_sweep_kernel().decode())
            if k == "scattering":
                scattering_kernels.add(A.kb200_last_scattering_kernel().decode())
            if k == "population":
                particles.append(r)  # 8-byte device->host read of the step's result
        if timed:
            A.kb200_device_sync()
            for k in kernels:
                ms = C.c_float()
                A.kb200_event_elapsed_ms(ev[k][0], ev[k][1], C.byref(ms))
                ktime[k].append(ms.value)

    p.call("zero:psi")
    for _ in range(warmup):
        step(False, True)

    # ---- timed region A: device-resident ("value") ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    cnt0 = C.c_uint64()
    A.kb200_launch_count(C.byref(cnt0), 1)
    barrier()
    A.kb200_event_record(ev_step[0], None)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(True, False)
    A.kb200_event_record(ev_step[1], None)
    barrier()
    wall_dev = time.perf_counter() - t0
    ms = C.c_float()
    A.kb200_event_elapsed_ms(ev_step[0], ev_step[1], C.byref(ms))
    launches = C.c_uint64()
    A.kb200_launch_count(C.byref(launches), 0)
    dev_s = ms.value * 1e-3

    # ---- timed region B: end to end through the host API with host buffers ("e2e") ----
    A.kb200_stream_create(C.byref(copy_stream))
    step(False, True)  # one untimed step with the two-stream upload
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(False, True)
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.result()
    A.kb200_stream_destroy(copy_stream)
    copy_stream = C.c_void_p()

    if dist is not None:
        import torch
        t = torch.tensor([dev_s, e2e_s, wall_dev], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s, e2e_s, wall_dev = t.tolist()

    if rank == 0:
        Z = unknowns // (WORKLOADS[args.workload][1] * WORKLOADS[args.workload][2])
        G, D, L = WORKLOADS[args.workload][1], WORKLOADS[args.workload][2], WORKLOADS[args.workload][3]
        M = (L + 1) ** 2
        N_u, N_m = float(unknowns), float(M * G * Z)
        # algorithmic bytes per launch, per GPU (SURVEY 8d3 / DESIGN.md)
        alg = {"LTimes": 8 * N_u + 8 * N_m, "LPlusTimes": 8 * N_u + 8 * N_m, "scattering": 16 * N_m,
               "SweepSolver": 16 * N_u, "population": 8 * N_u, "source": 0.0}
        flops = {"LTimes": 2 * M * N_u, "LPlusTimes": 2 * M * N_u, "scattering": 2 * G * N_m}
        peak, peak_src = measured_peaks()
        pop_fused = any("+population" in x for x in sweep_kernels)
        per_kernel = {}
        for k in kernels:
            ms_k = statistics.mean(ktime[k]) if ktime[k] else None
            if k == "population" and pop_fused:
                # the sum was accumulated by the sweep kernel while psi was in registers: this entry point only adds up the
                # per-CTA partials and reads 8 bytes back -- it moves no algorithmic bytes of its own
                per_kernel[k] = {"ms": ms_k, "fused": True, "fused_into": "SweepSolver", "alg_GBs_per_gpu": None, "frac_of_hbm_peak": None}
                continue
            if k == "source":
                per_kernel[k] = {"ms": ms_k, "alg_GBs_per_gpu": None, "frac_of_hbm_peak": None}
                continue
            per_kernel[k] = {"ms": ms_k, "alg_GBs_per_gpu": (alg[k] / n_gpus / (ms_k * 1e-3) / 1e9) if ms_k else None,
                             "frac_of_hbm_peak": (alg[k] / n_gpus / (ms_k * 1e-3) / 1e9 / peak) if ms_k else None}
            if k in flops and ms_k:
                per_kernel[k]["fp64_TFLOPs_per_gpu"] = flops[k] / n_gpus / (ms_k * 1e-3) / 1e12
        dom = max(kernels, key=lambda k: per_kernel[k]["ms"] or 0)
        grind_ns = 1e9 * dev_s / args.steps / unknowns
        e2e_ns = 1e9 * e2e_s / args.steps / unknowns
        traffic, traffic_src = ncu_traffic(args.workload, args.layout, dom) if n_gpus == 1 else (None, None)
        line = {"metric": "grind_time", "value": grind_ns, "unit": "ns/(unknown*iter)", "n_gpus": n_gpus,
                "steps": args.steps, "warmup": warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": False,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "throughput_unknowns_per_s": unknowns / (dev_s / args.steps),
                "e2e": {"value": e2e_ns, "unit": "ns/(unknown*iter)", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 8,
                        "ms_per_step": 1e3 * e2e_s / args.steps,
                        "what": "source iteration through the C++ Kripke:: host layer; all generated input tables "
                                "re-uploaded from pinned host memory (on a second stream, each joined right before the entry point that reads it) "
                                "and the particle count read back every step"},
                "gpu_launches": int(launches.value),
                "command": "kripke " + " ".join(kargs),
                "sweep_kernel": "|".join(sorted(sweep_kernels)),
                "scattering_kernel": "|".join(sorted(scattering_kernels)),
                "clocks": clocks,
                "roofline": {"kernel": dom, "bound": "hbm", "achieved": per_kernel[dom]["alg_GBs_per_gpu"], "peak": peak,
                             "unit": "GB/s", "frac": per_kernel[dom]["frac_of_hbm_peak"],
                             "traffic": traffic, "traffic_source": traffic_src,
                             "algorithmic_bytes": alg[dom] / n_gpus,
                             "peak_source": peak_src},
                "per_kernel": per_kernel,
                "per_step_ms": [sum(ktime[k][i] for k in kernels) for i in range(len(ktime[kernels[0]]))],
                "sweep_ms_per_step": list(ktime["SweepSolver"]),
                "particles_last": particles[-1] if particles else None,
                "wall_ms_per_step_device_region": 1e3 * wall_dev / args.steps}
        if parity is not None:
            line["parity"] = parity
        if n_gpus == 1 and not args.no_cpu_baseline:
            try:
                r = run_reference_cpu(args.workload, args.layout, 2, 1, (16, 16, 16))
                if r:
                    line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "command")}
            except Exception as e:  # the baseline is reported, never required
                line["cpu_baseline"] = {"error": str(e)}
        print(json.dumps(line))
    p.close()
    if dist is not None:
        A.kb200_comm_destroy()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
