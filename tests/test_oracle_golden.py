"""CPU: the plain-C oracle against the golden vectors produced by the unmodified reference
(tests/golden/reference_goldens.json, tools/make_golden.py) and, where the reference binary is
present (oracle/_ref, built from /root/reference), element-wise against its field dumps."""
import math
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import ROOT, oracle_problem

REF = os.path.join(ROOT, "oracle", "_ref", "kripke_ref")
FAST = ["G3_legendre9", "G4_kba_proxy", "G4b_undecomposed", "G6_gauss_legendre_4x4", "R1_ragged", "R2_legendre0",
        "R3_custom_xs", "L_DGZ", "L_DZG", "L_GDZ", "L_GZD", "L_ZDG", "L_ZGD"]


def l2_and_sum(chunks):
    s2 = math.fsum(float(x) * float(x) for c in chunks for x in c) if sum(len(c) for c in chunks) < 200000 else \
        float(sum(np.sum(c.astype(np.longdouble) ** 2) for c in chunks))
    s1 = float(sum(np.sum(c.astype(np.longdouble)) for c in chunks))
    return math.sqrt(s2), s1


@pytest.mark.parametrize("name", FAST)
def test_oracle_matches_reference_goldens(goldens, name):
    g = goldens[name]
    p, niter, bj = oracle_problem(g["args"])
    parts = p.solve(niter, bj)
    assert len(parts) == len(g["particles"])
    for a, b in zip(parts, g["particles"]):
        assert abs(a - b) <= 1e-13 * abs(b)  # only the population reduction order may differ
    for f in ("phi", "phi_out", "rhs", "psi"):
        chunks = [p.chunk(f, c) for c in range(p.num_chunks(f))]
        l2, s1 = l2_and_sum(chunks)
        ref = g["norms"][f]
        assert sum(len(c) for c in chunks) == ref["n"]
        assert abs(l2 - ref["l2"]) <= 1e-14 * ref["l2"]
        assert abs(s1 - ref["sum"]) <= 1e-11 * max(abs(ref["sum"]), ref["l2"])


def test_oracle_default_problem_first_iterations(goldens):
    """BASELINE config 1 (reference default), first 3 of the 10 golden iterations (the full 10 run on the GPU)."""
    g = goldens["G1_default"]
    p, _, bj = oracle_problem(g["args"])
    parts = p.solve(3, bj)
    for a, b in zip(parts, g["particles"][:3]):
        assert abs(a - b) <= 1e-13 * abs(b)


def test_oracle_decomposition_invariance(goldens):
    """SURVEY section 4 property 2: zset/gset re-decomposition leaves particle counts unchanged (1e-13)."""
    a, b = goldens["G4_kba_proxy"]["particles"], goldens["G4b_undecomposed"]["particles"]
    for x, y in zip(a, b):
        assert abs(x - y) <= 1e-12 * abs(y)


def test_sweep_visit_order_fixture():
    """SURVEY 3.3 fixture: --zones 8,8,8 --gset 1 --dset 8 --zset 2,2,1, 32 subdomains."""
    from oracle import ko
    p = ko.Problem(zones=(8, 8, 8), gset=1, dset=8, zset=(2, 2, 1))
    assert p.num_subdomains() == 32
    assert p.adjacency(0) == ([2, 1, -1], [-1, -1, -1])
    assert p.adjacency(1) == ([3, -1, -1], [-1, 0, -1])
    assert p.adjacency(28) == ([-1, -1, -1], [30, 29, -1])
    assert p.adjacency(31) == ([29, 30, -1], [-1, -1, -1])
    assert p.sweep_order() == [3, 1, 2, 0, 5, 4, 7, 6, 10, 8, 11, 9, 12, 13, 14, 15, 19, 17, 18, 16, 21, 20, 23, 22, 26,
                               24, 27, 25, 28, 29, 30, 31]


def test_generator_spot_values():
    """SURVEY 8d2 spot values of the default problem."""
    from oracle import ko
    p = ko.Problem()
    assert p.chunk("quadrature/xcos", 0)[0] == float.fromhex("0x1.6a09e667f3bcdp-1")
    assert p.chunk("quadrature/ycos", 0)[0] == 0.5
    assert p.chunk("quadrature/zcos", 0)[0] == float.fromhex("0x1.ffffffffffffep-2")
    assert p.chunk("quadrature/w", 0)[0] == float.fromhex("0x1.0c152382d7365p-3")
    assert p.chunk("dx", 0)[0] == 7.5 and p.chunk("dy", 0)[0] == 12.5 and p.chunk("volume", 0)[0] == 703.125
    assert p.chunk("ell", 0)[0] == 0.010416666666666666
    mats = p.field("mixelem_to_material")
    assert len(mats) == 4352 and [int((mats == m).sum()) for m in range(3)] == [32, 352, 3968]
    sig = p.field("sigt_zonal")
    assert sig.min() == pytest.approx(1e-4) and sig.max() == pytest.approx(0.1)
    assert len(p.field("psi")) == 12582912 and len(p.field("phi")) == 3276800 and len(p.field("i_plane")) == 786432


@pytest.mark.skipif(not os.path.exists(REF), reason="reference binary not built (needs /root/reference)")
@pytest.mark.parametrize("layout", ["DGZ", "GZD", "ZGD", "DZG", "GDZ", "ZDG"])
def test_oracle_bitwise_equals_reference_dump(layout):
    """every field of the oracle is BIT-identical to the reference's Sequential path."""
    args = f"--zones 8,6,8 --groups 8 --quad 16 --legendre 2 --niter 2 --zset 2,1,2 --gset 2 --dset 8 --layout {layout}"
    p, niter, bj = oracle_problem(args)
    with tempfile.TemporaryDirectory() as d:
        out = subprocess.check_output([REF, "--arch", "Sequential", "--dump", d] + args.split(), text=True)
        ref_parts = [float(l.split("=")[1]) for l in out.splitlines() if l.startswith("ITER")]
        mine = p.solve(niter, bj)
        for a, b in zip(mine, ref_parts):
            assert abs(a - b) <= 1e-13 * abs(b)
        for name in ["psi", "rhs", "phi", "phi_out", "i_plane", "j_plane", "k_plane", "ell", "ell_plus", "data/sigs",
                     "sigt_zonal", "quadrature/w", "volume", "mixelem_to_fraction"]:
            ref = np.fromfile(os.path.join(d, name.replace("/", "_") + ".bin"), dtype=np.float64)
            got = p.field(name)
            assert ref.shape == got.shape and np.array_equal(ref.view(np.uint64), got.view(np.uint64)), name
        for name in ["zone_to_mixelem", "mixelem_to_zone", "mixelem_to_material", "moment_to_legendre", "upwind", "downwind"]:
            ref = np.fromfile(os.path.join(d, name + ".bin"), dtype=np.int64)  # RAJA index types are 8 bytes
            assert np.array_equal(ref, p.field(name).astype(np.int64)), name
