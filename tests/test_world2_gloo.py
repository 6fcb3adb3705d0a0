"""CPU, world_size 2 over gloo: the host-side multi-rank logic (KBA decomposition with --procs, stage
schedule, face-exchange message matching) of two cooperating ranks, without any GPU work.

Each rank builds its share of the same problem exactly as a GPU rank would (kripke_b200_set_world +
the reference's --procs decomposition, src/Kripke/Core/PartitionSpace.cpp:25-43), then the ranks
compare notes through torch.distributed: every face a rank plans to send in stage t must be a face
its peer plans to receive in stage t+1 (the NCCL send/recv pairs of host/sweep_solver.cpp), and the
per-rank generated fields must tile the undecomposed oracle's fields."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT

ARGS = "--zones 8,12,8 --groups 4 --quad 16 --legendre 1 --gset 2 --dset 8 --zset 2,1,2 --procs 2,1,1 --layout DGZ"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rank_main(rank, world, port, pmethod, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    import kripke_b200 as kb
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        kb.host().kripke_b200_set_world(rank, world)
        p = kb.Problem(ARGS + (" --pmethod bj" if pmethod == "bj" else ""))
        mine = dict(rank=rank, nsdom=p.num_subdomains(), sched=p.sweep_schedule(),
                    l2g=p.field("SdomId2GlobalSdomId").astype(np.int64).tolist(),
                    upwind=p.field("upwind").astype(np.int64).reshape(-1, 3).tolist(),
                    downwind=p.field("downwind").astype(np.int64).reshape(-1, 3).tolist(),
                    volume_sum=float(p.field("volume").sum()), nmix=int(p.field("mixelem_to_material").size),
                    sigt_sum=float(p.field("sigt_zonal").sum()))
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        if rank == 0:
            out.put(everyone)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _run_world(pmethod):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, pmethod, out)) for r in range(2)]
    for pr in procs:
        pr.start()
    everyone = out.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    return everyone


@pytest.mark.parametrize("pmethod", ["sweep", "bj"])
def test_two_ranks_agree_on_decomposition_and_face_exchange(native_built, pmethod):
    from oracle import ko
    ranks = _run_world(pmethod)
    assert [r["rank"] for r in ranks] == [0, 1]
    # 2 x 1 x 1 ranks, 2 x 1 x 2 zone sets per rank, 2 group sets, 8 direction sets
    assert all(r["nsdom"] == 8 * 2 * 4 for r in ranks)
    # the ranks tile the undecomposed problem (--procs px,.. == single rank --zset px*a,..)
    o = ko.Problem(zones=(8, 12, 8), groups=4, quad=16, legendre=1, gset=2, dset=8, zset=(4, 1, 2))
    assert abs(sum(r["volume_sum"] for r in ranks) - float(o.field("volume").sum())) <= 1e-9 * float(o.field("volume").sum())
    assert sum(r["nmix"] for r in ranks) == o.field("mixelem_to_material").size
    assert abs(sum(r["sigt_sum"] for r in ranks) - float(o.field("sigt_zonal").sum())) <= 1e-12 * float(o.field("sigt_zonal").sum())
    # global ids are a permutation of 0..127
    assert sorted(g for r in ranks for g in r["l2g"]) == list(range(128))

    # message matching: (sender rank, stage, receiver global id, dim) sent == received one stage later
    sends, recvs = set(), set()
    for r in ranks:
        for s in r["sched"]:
            sd = s["sdom"]
            for dim in range(3):
                if s["send_to"][dim] >= 0 and s["send_to"][dim] != r["rank"]:
                    sends.add((r["rank"], s["send_to"][dim], s["stage"], r["downwind"][sd][dim], dim))
                if s["recv_from"][dim] >= 0 and s["recv_from"][dim] != r["rank"]:
                    stage_sent = s["stage"] - 1 if pmethod == "sweep" else s["stage"]
                    recvs.add((s["recv_from"][dim], r["rank"], stage_sent, r["l2g"][sd], dim))
    assert sends == recvs
    assert len(sends) > 0
    # in a 2 x 1 x 1 decomposition only i faces (dim 0) cross ranks
    assert {m[4] for m in sends} == {0}
    if pmethod == "sweep":
        # every octant's pipeline crosses the rank boundary exactly once per (group set, direction set, y-z zone set)
        assert len(sends) == 8 * 2 * 2
