"""CPU: host-side logic of the product (no compute): generator vs oracle, C-ABI surface, loud failure
without a GPU, command-line handling, sweep schedule."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT

GEN_FIELDS = ["ell", "ell_plus", "data/sigs", "sigt_zonal", "quadrature/xcos", "quadrature/ycos", "quadrature/zcos",
              "quadrature/w", "quadrature/id", "quadrature/jd", "quadrature/kd", "dx", "dy", "dz", "volume",
              "zone_to_num_mixelem", "zone_to_mixelem", "mixelem_to_zone", "mixelem_to_material", "mixelem_to_fraction",
              "moment_to_legendre", "upwind", "downwind"]


def test_abi_library_exports_every_declared_symbol(native_built):
    header = open(os.path.join(ROOT, "include", "kripke_b200.h")).read()
    names = sorted(set(re.findall(r"\b(kb200_[a-z0-9_]+)\s*\(", header)))
    assert len(names) > 40
    lib = C.CDLL(native_built[0])
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.kb200_abi_version() == 1


def test_no_cpu_fallback(native_built):
    """Without a GPU the product must fail loudly, never compute on the host."""
    import kripke_b200 as kb
    if kb.have_gpu():
        pytest.skip("GPU present")
    with pytest.raises(kb.KB200Error):
        kb.init_device(0)
    exe = os.path.join(ROOT, "kripke_b200", "bin", "kripke.exe")
    r = subprocess.run([exe, "--niter", "1"], capture_output=True, text=True)
    assert r.returncode != 0 and "no CUDA device" in (r.stdout + r.stderr)


def test_product_never_references_the_oracle():
    for base, _, files in os.walk(os.path.join(ROOT, "kripke_b200")):
        if "/build" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cpp", ".h", ".cu", ".cuh", "Makefile")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "oracle/" not in text and "kripke_oracle" not in text and "from oracle" not in text, f


@pytest.mark.parametrize("layout", ["DGZ", "DZG", "GDZ", "GZD", "ZDG", "ZGD"])
def test_host_generator_bitwise_equals_oracle(native_built, layout):
    import kripke_b200 as kb
    from oracle import ko
    p = kb.Problem(f"--zones 8,8,12 --groups 8 --quad 16 --legendre 3 --zset 2,1,2 --gset 2 --dset 8 --layout {layout}")
    o = ko.Problem(zones=(8, 8, 12), groups=8, quad=16, legendre=3, zset=(2, 1, 2), gset=2, dset=8, layout=layout)
    for n in GEN_FIELDS:
        a, b = p.field(n), o.field(n)
        assert a.shape == b.shape and np.array_equal(a.astype(b.dtype), b), n
    assert p.num_subdomains() == o.num_subdomains() == 64
    assert p.num_unknowns() == 8 * 16 * 8 * 8 * 12


def test_host_gauss_legendre_matches_oracle(native_built):
    import kripke_b200 as kb
    from oracle import ko
    p = kb.Problem("--zones 8,8,8 --groups 4 --quad 4:4 --legendre 2")
    o = ko.Problem(zones=(8, 8, 8), groups=4, quad=(4, 4), legendre=2)
    for n in GEN_FIELDS:
        assert np.array_equal(p.field(n).astype(o.field(n).dtype), o.field(n)), n


def test_command_line_validation(native_built):
    import kripke_b200 as kb
    for bad in ["--groups 33 --gset 2", "--quad 4", "--quad 100", "--legendre -1", "--niter 0", "--zset 0,1,1",
                "--layout XYZ", "--arch OpenMP", "--arch Sequential", "--bogus 1", "--zones 16,16"]:
        with pytest.raises(ValueError):
            kb.Problem(bad)
    kb.Problem("--arch CUDA --layout zgd --pmethod BJ --name x --sigt 1,2,3 --sigs 0.1,0.2,0.3")


def test_stage_schedule_is_a_topological_order(native_built):
    """the batched SweepSolver visits subdomains stage by stage; every upwind neighbour sits in an earlier stage."""
    import kripke_b200 as kb
    from oracle import ko
    args = "--zones 8,8,8 --gset 1 --dset 8 --zset 2,2,1"
    p = kb.Problem(args)
    o = ko.Problem(zones=(8, 8, 8), gset=1, dset=8, zset=(2, 2, 1))
    sched = p.sweep_schedule()
    assert sorted(s["sdom"] for s in sched) == list(range(32))
    stage = {s["sdom"]: s["stage"] for s in sched}
    for s in range(32):
        up, down = o.adjacency(s)
        for u in up:
            if u >= 0:
                assert stage[u] == stage[s] - 1
        assert all(x == -1 for x in sched[0]["recv_from"])
    assert max(stage.values()) == 2 + 2 + 1 - 3
