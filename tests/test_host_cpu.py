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


class _ScatDesc(C.Structure):  # kb200_scattering_desc (include/kripke_b200.h)
    _fields_ = [("layout", C.c_int), ("M", C.c_int), ("L1", C.c_int), ("G", C.c_int), ("Gs", C.c_int), ("Zs", C.c_int),
                ("nsrc", C.c_int), ("glower_dst", C.c_int), ("accumulate", C.c_int), ("glower_src", C.c_int * 64),
                ("phi_src", C.c_void_p * 64), ("phi_out", C.c_void_p), ("sigs", C.c_void_p), ("moment_to_legendre", C.c_void_p),
                ("zone_to_mixelem", C.c_void_p), ("zone_to_num_mixelem", C.c_void_p), ("mixelem_to_material", C.c_void_p),
                ("mixelem_to_fraction", C.c_void_p)]


def _scat_descs(ngsets, nzsets, Gs, Zs, M=25, L1=5, layout=0, zone_major=False):
    """The list Kernel::scattering hands over: one descriptor per (destination group set, zone set), every one reading all
    source group sets of its zone set.  Pointers are fake but aligned and distinct per chunk."""
    n = ngsets * nzsets
    d = (_ScatDesc * n)()
    pairs = [(g, z) for z in range(nzsets) for g in range(ngsets)] if zone_major else [(g, z) for g in range(ngsets) for z in range(nzsets)]
    for i, (g, z) in enumerate(pairs):
        e = d[i]
        e.layout, e.M, e.L1, e.G, e.Gs, e.Zs, e.nsrc, e.glower_dst, e.accumulate = layout, M, L1, ngsets * Gs, Gs, Zs, ngsets, g * Gs, 0
        for s in range(ngsets):
            e.glower_src[s] = s * Gs
            e.phi_src[s] = 0x10000000 + (z * ngsets + s) * 0x100000
        e.phi_out = 0x50000000 + (z * ngsets + g) * 0x100000
        e.sigs, e.moment_to_legendre = 0x90000000, 0x90100000
        e.zone_to_mixelem, e.zone_to_num_mixelem = 0xA0000000 + z * 0x1000, 0xA1000000 + z * 0x1000
        e.mixelem_to_material, e.mixelem_to_fraction = 0xA2000000 + z * 0x1000, 0xA3000000 + z * 0x1000
    return d, n


def test_scattering_plan_groups_descriptors_that_share_their_sources(native_built):
    """Host logic of kb200_scatter_slab.cu, no device needed: which descriptor lists the one-read kernel takes, how it groups
    them (all destination group sets of a zone set form one group) and where it splits the outputs over sibling CTAs."""
    import kripke_b200 as kb
    A = kb.abi()
    A.kb200_scattering_plan.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]

    def plan(d, n):
        g, o, s = C.c_int(-1), C.c_int(-1), C.c_int(-1)
        rc = A.kb200_scattering_plan(C.byref(d), n, C.byref(g), C.byref(o), C.byref(s))
        return (rc, g.value, o.value, s.value)

    # BASELINE config 2: two group sets of 32, one zone set -> one group, 64 outputs in one CTA
    assert plan(*_scat_descs(2, 1, 32, 64 ** 3)) == (0, 1, 64, 1)
    # config 4 per GPU: four group sets of 16, 64 zone sets, in either enumeration order -> 64 groups of four descriptors
    assert plan(*_scat_descs(4, 64, 16, 16 ** 3)) == (0, 64, 64, 1)
    assert plan(*_scat_descs(4, 64, 16, 16 ** 3, zone_major=True)) == (0, 64, 64, 1)
    # config 3: 128 groups in one set, M = 100: the three 128 x 128 matrices do not fit -> four siblings of 32 outputs
    assert plan(*_scat_descs(1, 1, 128, 32 ** 3, M=100, L1=10)) == (0, 1, 32, 4)
    assert plan(*_scat_descs(3, 2, 32, 480, M=4, L1=2)) == (0, 2, 32, 3)          # 96 outputs: three siblings
    assert plan(*_scat_descs(1, 1, 32, 720, layout=2)) == (0, 1, 32, 1)           # GDZ
    # declined: outputs not a multiple of 32, zone count not a multiple of 4, a moment-fastest nesting, ...
    assert plan(*_scat_descs(1, 1, 36, 720))[0] == -1
    assert plan(*_scat_descs(1, 1, 32, 630))[0] == -1
    assert plan(*_scat_descs(2, 1, 32, 720, layout=5))[0] == -1
    # ... descriptors with different cross-section tables, an unaligned destination chunk, a group that misses a descriptor
    d, n = _scat_descs(2, 2, 32, 720)
    d[3].sigs = 0x90000100
    assert plan(d, n)[0] == -1
    d, n = _scat_descs(2, 2, 32, 720)
    d[1].phi_out += 8
    assert plan(d, n)[0] == -1
    d, n = _scat_descs(2, 2, 32, 720)
    assert plan(d, n - 1)[0] == -1
