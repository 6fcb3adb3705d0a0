#!/usr/bin/env python
"""Multi-GPU parity check, one rank per GPU (launch with torchrun --nproc-per-node N, N = 2, 4 or 8):
the KBA sweep / block-Jacobi exchange over NCCL must reproduce the particle counts the unmodified
reference produced for the equivalent single-rank zone-set decomposition (--procs px,py,pz
--zset a,b,c == one rank with --zset px*a,py*b,pz*c).  Rank 0 prints one PASS/FAIL line per case."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kripke_b200 as kb  # noqa: E402

PROCS = {2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
RTOL = 1e-12


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    A, H = kb.abi(), kb.host()
    kb.init_device(local)
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (C.c_ubyte * 128)()
        kb.api.check(A.kb200_comm_unique_id(buf), "kb200_comm_unique_id")
        uid = torch.tensor(list(buf), dtype=torch.uint8)
    uid = uid.cuda()
    dist.broadcast(uid, 0)
    kb.api.check(A.kb200_comm_init(rank, world, bytes(uid.cpu().tolist())), "kb200_comm_init")
    H.kripke_b200_set_world(rank, world)

    with open(os.path.join(ROOT, "tests", "golden", "reference_goldens.json")) as f:
        gold = json.load(f)
    px, py, pz = PROCS[world]
    cases = []
    for name in ("G4_kba_proxy", "G5_block_jacobi", "G1z_zset222", "L_GZD", "L_ZGD"):
        a = gold[name]["args"].split()
        zs = [1, 1, 1]
        if "--zset" in a:
            i = a.index("--zset")
            zs = [int(x) for x in a[i + 1].split(",")]
            del a[i:i + 2]
        if zs[0] % px or zs[1] % py or zs[2] % pz:
            continue
        a += ["--zset", "%d,%d,%d" % (zs[0] // px, zs[1] // py, zs[2] // pz), "--procs", "%d,%d,%d" % (px, py, pz)]
        cases.append((name, a))
    ok = True
    for name, a in cases:
        p = kb.Problem(a)
        got = p.solve()
        ref = gold[name]["particles"]
        err = max(abs(x - y) / abs(y) for x, y in zip(got, ref))
        good = len(got) == len(ref) and err <= RTOL
        ok &= good
        if rank == 0:
            print(f"{'PASS' if good else 'FAIL'} {name} ranks={world} args={' '.join(a)} max_rel_err={err:.3e}", flush=True)
        p.close()
        dist.barrier()
    A.kb200_comm_destroy()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
