"""GPU (-m gpu): parity of the CUDA path, called through the C++ host layer and the C ABI, against
the CPU oracle on identical seeded inputs and against the reference's golden vectors.

Tolerances (BASELINE.json north_star): per-iteration particle count and ||phi||_2 within 1e-12
relative (fp64 re-ordering / FMA contraction only).  With kb200_set_exact(1) the kernels keep the
reference's multiply-then-add arithmetic and summation order, and every field must then be
BIT-identical to the oracle (which is bit-identical to the reference's Sequential path)."""
import ctypes as C
import math

import numpy as np
import pytest

from conftest import oracle_problem

pytestmark = pytest.mark.gpu
LAYOUTS = ["DGZ", "DZG", "GDZ", "GZD", "ZDG", "ZGD"]
RTOL = 1e-12


def seeded(n, seed, lo=0.0, hi=1.0):
    return np.random.default_rng(seed).uniform(lo, hi, n)


def fill_both(p, o, field, seed, lo=0.0, hi=1.0):
    for c in range(o.num_chunks(field)):
        v = seeded(len(o.chunk(field, c)), seed + c, lo, hi)
        o.chunk(field, c)[:] = v
        p.set_chunk(field, c, v)


def assert_close(got, ref, what, exact):
    assert got.shape == ref.shape, what
    if exact:
        assert np.array_equal(got.view(np.uint64), ref.view(np.uint64)), f"{what}: not bit-identical"
    else:
        scale = max(float(np.max(np.abs(ref))), 1e-300)
        err = float(np.max(np.abs(got - ref))) / scale
        assert err <= RTOL, f"{what}: max rel err {err:.3e}"


def pair(gpu, args):
    p = gpu.Problem(args)
    o, niter, bj = oracle_problem(args)
    return p, o, niter, bj


SMALL = "--zones 12,8,10 --groups 8 --quad 32 --legendre 3 --gset 2 --dset 8 --zset 2,1,2"


@pytest.mark.parametrize("exact", [False, True])
@pytest.mark.parametrize("layout", LAYOUTS)
def test_ltimes(gpu, layout, exact):
    gpu.abi().kb200_set_exact(int(exact))
    p, o, _, _ = pair(gpu, f"{SMALL} --layout {layout}")
    fill_both(p, o, "psi", 100, -1.0, 2.0)
    o.zero("phi"); o.ltimes()
    p.call("zero:phi"); p.call("LTimes")
    assert_close(p.field("phi"), o.field("phi"), f"LTimes {layout}", exact)
    gpu.abi().kb200_set_exact(0)


@pytest.mark.parametrize("exact", [False, True])
@pytest.mark.parametrize("layout", LAYOUTS)
def test_lplustimes(gpu, layout, exact):
    gpu.abi().kb200_set_exact(int(exact))
    p, o, _, _ = pair(gpu, f"{SMALL} --layout {layout}")
    fill_both(p, o, "phi_out", 200, -1.0, 1.0)
    o.zero("rhs"); o.lplustimes()
    p.call("zero:rhs"); p.call("LPlusTimes")
    assert_close(p.field("rhs"), o.field("rhs"), f"LPlusTimes {layout}", exact)
    gpu.abi().kb200_set_exact(0)


MOMENT_SHAPES = [
    # (legendre, quad): M = 25 exercises the 3-tile + DFMA-row LTimes and the K = 4*6+1 LPlusTimes paths of
    # kb200_moments_mma.cu; M = 100 the wide-output / streamed-K paths; M = 1 the degenerate one
    (4, 96), (9, 16), (0, 8), (5, 40),
    (4, 192),  # BASELINE config 2's shape per direction set (Ds = 24, M = 25): 16-byte fragment loads + the DFMA column/row
    (5, 64),   # M = 36, 64 directions: LPlusTimes as a plain product on the slab kernel (kb200_gemm_slab_try), 64 outputs per CTA
    (9, 256),  # M = 100, 256 directions: the same with four sibling CTAs of 64 outputs (BASELINE config 3's regime); in GZD/ZGD the
               # 3-group sets are too narrow for the row-operand tensor kernel: the DFMA row kernel with a 128-row tile
]


@pytest.mark.parametrize("layout", ["DGZ", "DZG", "GDZ", "GZD", "ZGD"])  # kb200_moments_mma.cu / kb200_moments_rowmma.cu
@pytest.mark.parametrize("shape", range(len(MOMENT_SHAPES)))
def test_moments_tensor_core_shapes(gpu, shape, layout):
    L, quad = MOMENT_SHAPES[shape]
    args = f"--zones 10,6,8 --groups 6 --quad {quad} --legendre {L} --gset 2 --dset 8 --zset 1,2,1 --layout {layout}"
    p, o, _, _ = pair(gpu, args)
    A = gpu.abi()
    A.kb200_ltimes_slab_launches.restype = C.c_ulonglong
    slab0 = A.kb200_ltimes_slab_launches()
    fill_both(p, o, "psi", 1100 + shape, -1.0, 2.0)
    o.zero("phi"); o.ltimes()
    p.call("zero:phi"); p.call("LTimes")
    assert_close(p.field("phi"), o.field("phi"), f"LTimes L={L} {layout}", False)
    # M = 25 (or 100 = 4 sibling CTAs of 25 moments) with contiguous (group, zone) columns runs on kb200_moments_slab.cu
    assert (A.kb200_ltimes_slab_launches() > slab0) == (L in (4, 9) and layout in ("DGZ", "DZG", "GDZ")), (L, layout)
    A.kb200_lplustimes_slab_launches.restype = C.c_ulonglong
    slab1 = A.kb200_lplustimes_slab_launches()
    fill_both(p, o, "phi_out", 1200 + shape, -1.0, 1.0)
    o.zero("rhs"); o.lplustimes()
    p.call("zero:rhs"); p.call("LPlusTimes")
    assert_close(p.field("rhs"), o.field("rhs"), f"LPlusTimes L={L} {layout}", False)
    assert (A.kb200_lplustimes_slab_launches() > slab1) == (L >= 5 and quad % 32 == 0 and layout in ("DGZ", "DZG")), (L, quad, layout)
    # accumulate semantics (no pending zero-fill): a second call adds on top
    o.ltimes(); p.call("LTimes")
    assert_close(p.field("phi"), o.field("phi"), f"LTimes accumulate L={L} {layout}", False)
    o.lplustimes(); p.call("LPlusTimes")
    assert_close(p.field("rhs"), o.field("rhs"), f"LPlusTimes accumulate L={L} {layout}", False)


@pytest.mark.parametrize("exact", [False, True])
@pytest.mark.parametrize("layout", LAYOUTS)
def test_scattering_and_source_with_dense_asymmetric_sigs(gpu, layout, exact):
    """the shipped sigs is diagonal, which would hide a g/gp swap: use a dense random one."""
    gpu.abi().kb200_set_exact(int(exact))
    p, o, _, _ = pair(gpu, f"{SMALL} --layout {layout}")
    fill_both(p, o, "data/sigs", 300, 0.0, 0.1)
    fill_both(p, o, "phi", 400, -1.0, 1.0)
    o.zero("phi_out"); o.scattering(); o.source()
    p.call("zero:phi_out"); p.call("scattering"); p.call("source")
    assert_close(p.field("phi_out"), o.field("phi_out"), f"scattering+source {layout}", exact)
    gpu.abi().kb200_set_exact(0)


@pytest.mark.parametrize("exact", [False, True])
@pytest.mark.parametrize("layout", LAYOUTS)
def test_sweep_subdomain_with_incoming_faces(gpu, layout, exact):
    """single-subdomain entry point with non-zero incoming face fluxes on all three planes."""
    gpu.abi().kb200_set_exact(int(exact))
    p, o, _, _ = pair(gpu, f"{SMALL} --layout {layout}")
    fill_both(p, o, "rhs", 500, 0.0, 1.0)
    for f, s in (("i_plane", 600), ("j_plane", 700), ("k_plane", 800)):
        fill_both(p, o, f, s, 0.0, 0.5)
    for sdom in (0, 5, o.num_subdomains() - 1):
        o.sweep_subdomain(sdom)
        p.call(f"sweepSubdomain:{sdom}")
    for f in ("psi", "i_plane", "j_plane", "k_plane"):
        got, ref = p.field(f), o.field(f)
        if f == "psi":  # only the swept subdomains are defined
            n = len(o.chunk("psi", 0))
            for sdom in (0, 5, o.num_subdomains() - 1):
                assert_close(got[sdom * n:(sdom + 1) * n], ref[sdom * n:(sdom + 1) * n], f"sweep psi {layout}", exact)
        else:
            assert_close(got, ref, f"sweep {f} {layout}", exact)
    gpu.abi().kb200_set_exact(0)


ZLINE_CASES = [
    # zone-fastest layouts with ni % 4 == 0 take the line-streaming kernel (kb200_sweep_zline.cu)
    "--zones 16,40,20 --groups 4 --quad 16 --legendre 1 --gset 1 --dset 8 --zset 1,1,1",   # 2x2 column tiles, ragged
    "--zones 8,8,4 --groups 6 --quad 24 --legendre 1 --gset 1 --dset 8 --zset 1,1,1",      # packed sub-streams
    "--zones 32,16,48 --groups 4 --quad 16 --legendre 1 --gset 2 --dset 8 --zset 2,1,3",   # decomposed, 16^3 subdomains
    "--zones 4,33,17 --groups 3 --quad 8 --legendre 0 --gset 1 --dset 8 --zset 1,1,1",     # one-lane / one-warp tiles
    "--zones 16,12,20 --groups 8 --quad 48 --legendre 1 --gset 1 --dset 8 --zset 1,1,1",   # 48 elements: two 32-element slices
    # default arithmetic on DGZ/GDZ with ni = 4 * 2^k takes the i-row scan kernel (kb200_sweep_irow.cu):
    "--zones 16,40,44 --groups 3 --quad 16 --legendre 0 --gset 1 --dset 8 --zset 1,1,1",   # nk > 32: two k tiles of 22 planes
    "--zones 64,6,5 --groups 3 --quad 8 --legendre 0 --gset 1 --dset 8 --zset 1,1,1",      # 16 lanes per row, odd element count
    "--zones 128,4,3 --groups 2 --quad 8 --legendre 0 --gset 1 --dset 8 --zset 1,1,1",     # one row per warp (32-lane scan)
    "--zones 32,34,33 --groups 2 --quad 8 --legendre 0 --gset 1 --dset 8 --zset 1,1,1",    # 8 lanes per row, k tiles 17+16
    # rows whose lane count is not a power of two ride in the next larger segment with idle lanes at its end:
    "--zones 48,6,5 --groups 3 --quad 8 --legendre 0 --gset 1 --dset 8 --zset 1,1,1",      # 12 of 16 lanes, odd nk
    "--zones 24,8,18 --groups 2 --quad 16 --legendre 0 --gset 1 --dset 8 --zset 1,1,1",    # 6 of 8 lanes, row pairs, two k tiles
    "--zones 96,4,3 --groups 2 --quad 8 --legendre 0 --gset 1 --dset 8 --zset 1,1,1",      # 24 of 32 lanes
    "--zones 20,7,4 --groups 2 --quad 8 --legendre 0 --gset 1 --dset 8 --zset 1,1,1",      # 5 of 8 lanes, odd nj (single rows)
]


@pytest.mark.parametrize("exact", [False, True])
@pytest.mark.parametrize("layout", LAYOUTS)  # DGZ/GDZ: kb200_sweep_zline.cu; the other four: kb200_sweep_elem.cu
@pytest.mark.parametrize("case", range(len(ZLINE_CASES)))
def test_sweep_zone_fastest_line_kernel(gpu, case, layout, exact):
    gpu.abi().kb200_set_exact(int(exact))
    try:
        p, o, _, _ = pair(gpu, f"{ZLINE_CASES[case]} --layout {layout}")
        fill_both(p, o, "rhs", 1500 + case, 0.0, 1.0)
        for f, s in (("i_plane", 1600), ("j_plane", 1700), ("k_plane", 1800)):
            fill_both(p, o, f, s, 0.0, 0.5)
        n = len(o.chunk("psi", 0))
        sdoms = sorted(set([0, 3, o.num_subdomains() // 2, o.num_subdomains() - 1]))
        for sdom in sdoms:  # every entry point call sees non-zero inflow on all three faces
            o.sweep_subdomain(sdom)
            p.call(f"sweepSubdomain:{sdom}")
        got = p.field("psi"); ref = o.field("psi")
        for sdom in sdoms:
            assert_close(got[sdom * n:(sdom + 1) * n], ref[sdom * n:(sdom + 1) * n], f"zline psi {layout} sdom {sdom}", exact)
        for f in ("i_plane", "j_plane", "k_plane"):
            assert_close(p.field(f), o.field(f), f"zline {f} {layout}", exact)
        # the batched solver (all octants, on-rank face delivery) on top of a fresh rhs
        fill_both(p, o, "rhs", 1900 + case, 0.0, 1.0)
        o.sweep_solver(False)
        p.call("SweepSolver")
        for f in ("psi", "i_plane", "j_plane", "k_plane"):
            assert_close(p.field(f), o.field(f), f"zline SweepSolver {f} {layout}", exact)
    finally:
        gpu.abi().kb200_set_exact(0)


@pytest.mark.parametrize("layout", LAYOUTS)
def test_sweep_solver_and_population(gpu, layout):
    p, o, _, _ = pair(gpu, f"{SMALL} --layout {layout}")
    fill_both(p, o, "rhs", 900, 0.0, 1.0)
    o.sweep_solver(False)
    p.call("SweepSolver")
    for f in ("psi", "i_plane", "j_plane", "k_plane"):
        assert_close(p.field(f), o.field(f), f"SweepSolver {f} {layout}", False)
    ref = o.population()
    got = p.call("population")
    assert abs(got - ref) <= RTOL * abs(ref)


@pytest.mark.parametrize("name", ["G1_default", "G1z_zset222", "G3_legendre9", "G4_kba_proxy", "G4b_undecomposed",
                                  "G5_block_jacobi", "G6_gauss_legendre_4x4", "G6b_gauss_legendre_8x8", "R1_ragged",
                                  "R2_legendre0", "R3_custom_xs", "L_DGZ", "L_DZG", "L_GDZ", "L_GZD", "L_ZDG", "L_ZGD"])
def test_full_solve_matches_reference_goldens(gpu, goldens, name):
    """Kripke::SteadyStateSolver on the GPU vs numbers the unmodified reference produced."""
    g = goldens[name]
    p = gpu.Problem(g["args"])
    parts = p.solve()
    assert len(parts) == len(g["particles"])
    for it, (a, b) in enumerate(zip(parts, g["particles"])):
        assert abs(a - b) <= RTOL * abs(b), f"{name} iter {it}: {a!r} vs {b!r}"
    for f in ("phi", "psi", "rhs", "phi_out"):
        v = p.field(f).astype(np.longdouble)
        l2 = float(np.sqrt(np.sum(v * v)))
        ref = g["norms"][f]
        assert len(v) == ref["n"]
        assert abs(l2 - ref["l2"]) <= RTOL * ref["l2"], f"{name} ||{f}||"


# Full-size BASELINE shapes.  The particle count is a SERIAL sum of up to 1.6e9 terms in the reference's Sequential
# path: its own rounding noise is ~sqrt(N)*eps ~ 3e-12 relative (SURVEY section 4's OpenMP value of G7 differs from the
# Sequential golden by 1.7e-12), so the count is held to 2e-11 here; the field norms (accumulated in extended
# precision on both sides) are held to the 1e-12 of north_star.
BIG_PARTICLE_RTOL = 2e-11


@pytest.mark.parametrize("name", ["G7_config3_full", "G8_config2_half_slab"])
def test_full_size_baseline_shapes_match_reference_goldens(gpu, goldens, name):
    """BASELINE config 3 at full size (M = 100 moments, 128 groups: the fp64-bound instantiations of the moment and
    scattering kernels) and the largest slab of BASELINE config 2 the golden host could hold (config 2's 64 groups x 192
    directions on a 64 x 64 x 32 zone block), against numbers the unmodified reference produced."""
    g = goldens[name]
    p = gpu.Problem(g["args"])
    parts = p.solve()
    assert len(parts) == len(g["particles"])
    for it, (a, b) in enumerate(zip(parts, g["particles"])):
        assert abs(a - b) <= BIG_PARTICLE_RTOL * abs(b), f"{name} iter {it}: {a!r} vs {b!r}"
    for f in ("phi", "phi_out", "rhs", "psi"):
        l2, n = p.norm2(f)
        ref = g["norms"][f]
        assert n == ref["n"]
        assert abs(l2 - ref["l2"]) <= RTOL * ref["l2"], f"{name} ||{f}||: {l2!r} vs {ref['l2']!r}"
    p.close()


def test_config2_full_size_layout_and_decomposition_invariance(gpu):
    """BASELINE config 2 at full size (64^3 zones x 64 groups x 192 directions, 59.5 GB of fields) cannot be run by the
    reference on the golden host, so it is pinned by the reference's own invariants:
    every field is independent of the storage order and of the zone-set decomposition.  Two source iterations in DGZ,
    GZD and ZGD (three different sweep / scattering / moment kernels) and in DGZ with --zset 2,2,2 must agree on the
    particle counts and on ||phi||_2 to 1e-12."""
    base = "--zones 64,64,64 --groups 64 --quad 192 --legendre 4 --niter 2"
    res = {}
    for tag, extra in (("DGZ", "--layout DGZ"), ("GZD", "--layout GZD"), ("ZGD", "--layout ZGD"), ("DGZ_zset222", "--layout DGZ --zset 2,2,2")):
        p = gpu.Problem(f"{base} {extra}")
        parts = p.solve()
        res[tag] = (parts, p.norm2("phi")[0])
        p.close()
    ref_parts, ref_phi = res["DGZ"]
    assert ref_parts[1] > ref_parts[0] > 0.0
    for tag, (parts, phi) in res.items():
        for a, b in zip(parts, ref_parts):
            assert abs(a - b) <= RTOL * abs(b), f"config 2 {tag}: particles {a!r} vs DGZ {b!r}"
        assert abs(phi - ref_phi) <= RTOL * ref_phi, f"config 2 {tag}: ||phi|| {phi!r} vs DGZ {ref_phi!r}"


@pytest.mark.parametrize("layout", ["DGZ", "GZD", "ZGD"])
def test_full_solve_bit_exact_mode(gpu, layout):
    """exact mode: psi, phi, rhs, phi_out after 3 iterations are bit-identical to the oracle."""
    gpu.abi().kb200_set_exact(1)
    try:
        args = f"--zones 8,8,8 --groups 8 --quad 16 --legendre 2 --gset 2 --zset 2,2,1 --niter 3 --layout {layout}"
        p, o, niter, bj = pair(gpu, args)
        got = p.solve()
        ref = o.solve(niter, bj)
        for a, b in zip(got, ref):
            assert abs(a - b) <= 1e-13 * abs(b)
        for f in ("psi", "phi", "rhs", "phi_out", "i_plane", "j_plane", "k_plane"):
            assert_close(p.field(f), o.field(f), f"{f} {layout}", True)
    finally:
        gpu.abi().kb200_set_exact(0)


def test_reference_visit_order_mode(gpu, monkeypatch):
    """KB200_SWEEP_ORDER=reference: one subdomain at a time in the reference's queue order gives the same psi."""
    args = "--zones 8,8,8 --groups 4 --quad 16 --legendre 1 --gset 1 --dset 8 --zset 2,2,1 --niter 2"
    p1 = gpu.Problem(args)
    a = p1.solve()
    monkeypatch.setenv("KB200_SWEEP_ORDER", "reference")
    p2 = gpu.Problem(args)
    b = p2.solve()
    assert a == b
    assert np.array_equal(p1.field("psi"), p2.field("psi"))


def test_decomposition_invariance_at_scale(gpu):
    """size-independent property on a problem too big for the scalar oracle
    in test time: 32^3 x 32 groups x 96 directions, zset 1,1,1 vs 4,2,2 / gset 1 vs 4."""
    base = "--zones 32,32,32 --groups 32 --quad 96 --legendre 4 --niter 3 --layout ZGD"
    a = gpu.Problem(base + " --zset 1,1,1 --gset 1").solve()
    b = gpu.Problem(base + " --zset 4,2,2 --gset 4").solve()
    for x, y in zip(a, b):
        assert abs(x - y) <= RTOL * abs(y)


def test_layout_transform_round_trip(gpu):
    A = gpu.abi()
    na, ng, nz = 6, 5, 77
    n = na * ng * nz
    src = seeded(n, 42)
    bufs = [C.c_void_p() for _ in range(3)]
    for b in bufs:
        assert A.kb200_alloc(n * 8, C.byref(b)) == 0
    A.kb200_upload(bufs[0], src.ctypes.data_as(C.c_void_p), n * 8, None)
    for dst in range(6):
        assert A.kb200_layout_transform(0, dst, na, ng, nz, bufs[0], bufs[1], None) == 0
        assert A.kb200_layout_transform(dst, 0, na, ng, nz, bufs[1], bufs[2], None) == 0
        out = np.empty(n)
        A.kb200_download(out.ctypes.data_as(C.c_void_p), bufs[2], n * 8, None)
        A.kb200_stream_sync(None)
        assert np.array_equal(out, src)
        mid = np.empty(n)
        A.kb200_download(mid.ctypes.data_as(C.c_void_p), bufs[1], n * 8, None)
        A.kb200_stream_sync(None)
        cube = src.reshape(na, ng, nz)
        perm = {0: (0, 1, 2), 1: (0, 2, 1), 2: (1, 0, 2), 3: (1, 2, 0), 4: (2, 0, 1), 5: (2, 1, 0)}[dst]
        assert np.array_equal(mid, np.ascontiguousarray(cube.transpose(perm)).ravel())
    for b in bufs:
        A.kb200_free(b)


class _LTimesDesc(C.Structure):  # kb200_ltimes_desc (include/kripke_b200.h)
    _fields_ = [("layout", C.c_int), ("M", C.c_int), ("Ds", C.c_int), ("Gs", C.c_int), ("Zs", C.c_int), ("nsets", C.c_int),
                ("accumulate", C.c_int), ("ell", C.c_void_p * 64), ("psi", C.c_void_p * 64), ("phi", C.c_void_p)]


@pytest.mark.parametrize("shared_ell, zs", [(True, 96), (False, 96), (True, 90)])
def test_ltimes_abi_slab_kernel_and_its_fallbacks(gpu, shared_ell, zs):
    """kb200_ltimes straight through the C ABI at M = 25 (DGZ), two phi chunks x two direction sets, against numpy:
    chunks that share their ell tables run on kb200_moments_slab.cu; different tables per chunk, or a column count that is
    not a multiple of four, must be declined by it and served by the per-chunk kernel -- all with the same result."""
    A = gpu.abi()
    A.kb200_ltimes_slab_launches.restype = C.c_ulonglong
    A.kb200_ltimes.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    M, Ds, Gs, nsets, nchunk = 25, 12, 3, 2, 2
    rng = np.random.default_rng(77)
    held = []

    def dev(a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        p = C.c_void_p()
        assert A.kb200_alloc(a.nbytes, C.byref(p)) == 0
        A.kb200_upload(p, a.ctypes.data_as(C.c_void_p), a.nbytes, None)
        held.append(p)
        return p

    ells = [[rng.uniform(-1, 1, (Ds, M)) for _ in range(nsets)] for _ in range(1 if shared_ell else nchunk)]
    d_ells = [[dev(e) for e in es] for es in ells]
    descs = (_LTimesDesc * nchunk)()
    expect, d_phi = [], []
    for c in range(nchunk):
        es, des = (ells[0], d_ells[0]) if shared_ell else (ells[c], d_ells[c])
        psis = [rng.uniform(-1, 2, (Ds, Gs, zs)) for _ in range(nsets)]
        phi0 = rng.uniform(-1, 1, (M, Gs, zs))
        d = descs[c]
        d.layout, d.M, d.Ds, d.Gs, d.Zs, d.nsets, d.accumulate = 0, M, Ds, Gs, zs, nsets, 1
        for q in range(nsets):
            d.ell[q] = des[q].value
            d.psi[q] = dev(psis[q]).value
        d_phi.append(dev(phi0))
        d.phi = d_phi[-1].value
        expect.append(phi0 + sum(np.einsum("dm,dgz->mgz", es[q], psis[q]) for q in range(nsets)))
    before = A.kb200_ltimes_slab_launches()
    assert A.kb200_ltimes(C.byref(descs), nchunk, None) == 0
    assert (A.kb200_ltimes_slab_launches() > before) == (shared_ell and (Gs * zs) % 4 == 0)
    for c in range(nchunk):
        out = np.empty((M, Gs, zs))
        A.kb200_download(out.ctypes.data_as(C.c_void_p), d_phi[c], out.nbytes, None)
        A.kb200_stream_sync(None)
        assert_close(out, expect[c], f"kb200_ltimes chunk {c}", False)
    for p in held:
        A.kb200_free(p)


def test_fused_population_abi_matches_separate_kernel(gpu):
    """kb200_sweep_population + kb200_population_reduce (sum left behind by the sweep) == kb200_population on the
    same psi, and the host layer falls back to the separate kernel as soon as psi is written by anyone else."""
    args = "--zones 16,12,20 --groups 8 --quad 16 --legendre 1 --gset 2 --dset 8 --zset 1,1,1 --layout DGZ"
    p, o, _, _ = pair(gpu, args)
    fill_both(p, o, "rhs", 4100, 0.0, 1.0)
    o.sweep_solver(False)
    p.call("SweepSolver")
    fused = p.call("population")          # finishes the partial sums of the sweep kernels
    ref = o.population()
    assert abs(fused - ref) <= RTOL * abs(ref)
    # touching psi invalidates the cached partials: the value must follow the new contents
    v = seeded(len(o.chunk("psi", 0)), 4200, 0.0, 1.0)
    o.chunk("psi", 0)[:] = v
    p.set_chunk("psi", 0, v)
    ref2 = o.population()
    got2 = p.call("population")
    assert abs(got2 - ref2) <= RTOL * abs(ref2)
    assert abs(ref2 - ref) > 1e-6 * abs(ref)


def test_device_allocation_pool_reuses_blocks(gpu):
    A = gpu.abi()
    a, b = C.c_void_p(), C.c_void_p()
    size = (1 << 20) + 7 * 4096  # a size nothing else in the suite pools
    assert A.kb200_pool_trim() == 0  # the pool keeps at most 4 GB: start from an empty one
    assert A.kb200_alloc(size, C.byref(a)) == 0
    first = a.value
    assert A.kb200_free(a) == 0
    assert A.kb200_alloc(size, C.byref(b)) == 0
    assert b.value == first               # same size: the pooled block comes back, no cudaMalloc
    big = C.c_void_p()
    assert A.kb200_alloc(96 << 20, C.byref(big)) == 0 and A.kb200_free(big) == 0   # above the pool limit: plain cudaFree
    assert A.kb200_free(b) == 0


def test_irow_sweep_matches_line_kernel_at_full_zone_extent(gpu, monkeypatch):
    """size-independent property at BASELINE config 2's zone extent (64^3: four k tiles of 16 planes, 16-lane rows,
    TMA staging two rows ahead): the i-row scan kernel and the thread-per-row kernel (KB200_SWEEP_IROW=0), which is
    bit-exact against the oracle in the small cases, must agree on psi and on all outgoing faces."""
    args = "--zones 64,64,64 --groups 2 --quad 16 --legendre 0 --gset 1 --dset 8 --zset 1,1,1 --layout DGZ"
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("KB200_SWEEP_IROW", mode)
        p = gpu.Problem(args)
        for c in range(p.num_chunks("rhs")):
            p.set_chunk("rhs", c, seeded(len(p.chunk("rhs", c)), 5000 + c, 0.0, 1.0))
        p.call("SweepSolver")
        res[mode] = {f: p.field(f).copy() for f in ("psi", "i_plane", "j_plane", "k_plane")}
        res[mode]["pop"] = p.call("population")
        p.close()
    for f in ("psi", "i_plane", "j_plane", "k_plane"):
        assert_close(res["1"][f], res["0"][f], f"irow vs zline {f}", False)
    assert abs(res["1"]["pop"] - res["0"]["pop"]) <= RTOL * abs(res["0"]["pop"])


@pytest.mark.parametrize("legendre,quad", [(4, 96), (2, 16)])
def test_moments_tensor_core_packed_zdg(gpu, legendre, quad):
    """ZDG (group fastest): runs of Gs = 16 columns per zone are packed eight to a 128-column tile of kb200_moments_mma.cu"""
    args = f"--zones 10,6,8 --groups 32 --quad {quad} --legendre {legendre} --gset 2 --dset 8 --zset 1,2,1 --layout ZDG"
    p, o, _, _ = pair(gpu, args)
    fill_both(p, o, "psi", 6100, -1.0, 2.0)
    o.zero("phi"); o.ltimes()
    p.call("zero:phi"); p.call("LTimes")
    assert_close(p.field("phi"), o.field("phi"), "LTimes packed ZDG", False)
    fill_both(p, o, "phi_out", 6200, -1.0, 1.0)
    o.zero("rhs"); o.lplustimes()
    p.call("zero:rhs"); p.call("LPlusTimes")
    assert_close(p.field("rhs"), o.field("rhs"), "LPlusTimes packed ZDG", False)
    o.ltimes(); p.call("LTimes")  # accumulate on top
    assert_close(p.field("phi"), o.field("phi"), "LTimes accumulate packed ZDG", False)


@pytest.mark.parametrize("exact", [False, True])
@pytest.mark.parametrize("layout", ["DGZ", "GZD", "ZGD", "DZG"])
def test_sweep_with_non_uniform_mesh(gpu, layout, exact):
    """the generated mesh is uniform, which would hide a wrong 2*cos/delta per zone: stretch dx, dy, dz on both
    sides (SweepSubdomain.cpp:88-93 divides by the width of each zone) and sweep with non-zero inflow."""
    gpu.abi().kb200_set_exact(int(exact))
    try:
        args = f"--zones 16,10,12 --groups 4 --quad 16 --legendre 1 --gset 1 --dset 8 --zset 1,1,1 --layout {layout}"
        p, o, _, _ = pair(gpu, args)
        for name, seed in (("dx", 7100), ("dy", 7200), ("dz", 7300)):
            for c in range(o.num_chunks(name)):
                v = o.chunk(name, c) * seeded(len(o.chunk(name, c)), seed + c, 0.5, 1.5)
                o.chunk(name, c)[:] = v
                p.set_chunk(name, c, v)
        fill_both(p, o, "rhs", 7400, 0.0, 1.0)
        for f, s in (("i_plane", 7500), ("j_plane", 7600), ("k_plane", 7700)):
            fill_both(p, o, f, s, 0.0, 0.5)
        for sdom in range(o.num_subdomains()):
            o.sweep_subdomain(sdom)
            p.call(f"sweepSubdomain:{sdom}")
        for f in ("psi", "i_plane", "j_plane", "k_plane"):
            assert_close(p.field(f), o.field(f), f"non-uniform mesh {f} {layout}", exact)
    finally:
        gpu.abi().kb200_set_exact(0)


@pytest.mark.parametrize("layout", LAYOUTS)
def test_scattering_at_config2_group_structure(gpu, layout):
    """two source group sets of 32 groups, M = 25 (the 4-tile CTAs of kb200_scatter_mma.cu and, for the nestings whose
    zone index is not fastest, the transposes around it), dense asymmetric sigs, mixed-material zones, then '+='."""
    args = f"--zones 12,8,10 --groups 64 --quad 8 --legendre 4 --gset 2 --dset 8 --zset 1,2,1 --layout {layout}"
    p, o, _, _ = pair(gpu, args)
    fill_both(p, o, "data/sigs", 8100, 0.0, 0.1)
    fill_both(p, o, "phi", 8200, -1.0, 1.0)
    o.zero("phi_out"); o.scattering(); o.source()
    p.call("zero:phi_out"); p.call("scattering"); p.call("source")
    assert_close(p.field("phi_out"), o.field("phi_out"), f"scattering {layout}", False)
    o.scattering(); p.call("scattering")  # accumulate on top of the previous result
    assert_close(p.field("phi_out"), o.field("phi_out"), f"scattering accumulate {layout}", False)


SLAB_CASES = [
    # (arguments, kernel family expected in DGZ)                                                      what the shape exercises
    ("--zones 12,8,10 --groups 64 --quad 8 --legendre 4 --gset 2 --zset 1,2,1", "slab"),   # config 2: 64 outputs in one CTA, two zone sets, ragged last tile
    ("--zones 16,8,6 --groups 64 --quad 8 --legendre 2 --gset 4 --zset 2,1,1", "slab"),    # config 4: four destination sets of 16 groups per CTA
    ("--zones 20,6,6 --groups 128 --quad 8 --legendre 2 --gset 1 --zset 1,1,1", "slab"),   # config 3: four sibling CTAs of 32 outputs, 4 stages per tile
    ("--zones 12,6,6 --groups 96 --quad 8 --legendre 1 --gset 3 --zset 1,1,2", "slab"),    # three siblings, K = 96 (3 stages of 32)
    ("--zones 12,10,6 --groups 32 --quad 8 --legendre 3 --gset 1 --zset 1,1,1", "slab"),   # config 1 / 5: one 32-output CTA, K = 32
    ("--zones 12,10,6 --groups 36 --quad 8 --legendre 1 --gset 1 --zset 1,1,1", "mma"),    # outputs not a multiple of 32: per-descriptor kernel
    ("--zones 10,9,7 --groups 32 --quad 8 --legendre 1 --gset 1 --zset 1,1,1", "mma"),     # zone count not a multiple of 4
]


@pytest.mark.parametrize("layout", ["DGZ", "GDZ", "ZGD"])
@pytest.mark.parametrize("case", range(len(SLAB_CASES)))
def test_scattering_one_read_kernel_shapes(gpu, case, layout):
    """kb200_scatter_slab.cu (one staged read of the source moments for all destination group sets; sibling CTAs where the
    matrices of all outputs do not fit) against the oracle: dense asymmetric sigs, mixed-material zones, source folded in
    through the solver's entry point, then '+=' on top; shapes it declines must land on the per-descriptor kernel."""
    args, family = SLAB_CASES[case]
    p, o, _, _ = pair(gpu, f"{args} --dset 8 --layout {layout}")
    A = gpu.abi()
    A.kb200_last_scattering_kernel.restype = C.c_char_p
    fill_both(p, o, "data/sigs", 9100 + case, 0.0, 0.1)
    fill_both(p, o, "phi", 9200 + case, -1.0, 1.0)
    o.zero("phi_out"); o.scattering(); o.source()
    p.call("zero:phi_out"); p.call("scattering"); p.call("source")
    got = A.kb200_last_scattering_kernel().decode()
    assert got == (family if layout != "ZGD" else "transposed+" + family), got
    assert_close(p.field("phi_out"), o.field("phi_out"), f"scattering {layout} {args}", False)
    o.scattering(); p.call("scattering")  # accumulate on top of the previous result
    assert_close(p.field("phi_out"), o.field("phi_out"), f"scattering accumulate {layout} {args}", False)
    p.close()


def test_kripke_exe_command_line_and_output(gpu, goldens):
    """the drop-in executable: the reference's command line in, the reference's iteration lines, TIMER_NAMES/TIMER_DATA
    and figures of merit out (src/kripke.cpp:480-516, SteadyStateSolver.cpp:86-90, Timing.cpp:78-95)."""
    import os
    import re
    import subprocess
    from conftest import ROOT
    g = goldens["G1_default"]
    exe = os.path.join(ROOT, "kripke_b200", "bin", "kripke.exe")
    out = subprocess.run([exe] + g["args"].split(), capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    parts = [float(m) for m in re.findall(r"iter \d+: particle count=([0-9.e+-]+)", out.stdout)]
    assert len(parts) == len(g["particles"])
    for a, b in zip(parts, g["particles"]):
        assert abs(a - b) <= 1e-6 * abs(b)  # printed with %e like the reference
    names = re.search(r"TIMER_NAMES:(.*)", out.stdout).group(1).split(",")
    data = re.search(r"TIMER_DATA:(.*)", out.stdout).group(1).split(",")
    assert len(names) == len(data) and {"Solve", "LTimes", "LPlusTimes", "Scattering", "Source", "SweepSolver", "SweepSubdomain",
                                         "Population", "Generate"} <= {n.strip() for n in names}
    assert all(float(x) >= 0.0 for x in data)
    assert "Figures of Merit" in out.stdout and "Grind time" in out.stdout and "Number of unknowns: 12582912" in out.stdout
    assert out.stdout.rstrip().endswith("END")
    assert "Min time/rank" not in out.stdout
    # Caliper's runtime-report of the reference's nested regions (Timing.h:93-109, README "CALI_CONFIG_PROFILE")
    env = dict(os.environ, CALI_CONFIG_PROFILE="runtime-report")
    out = subprocess.run([exe] + g["args"].split(), capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0
    rows = [l for l in out.stdout.splitlines() if re.match(r"^\s*(Generate|Solve|LTimes|LPlusTimes|Scattering|Source|SweepSolver|SweepSubdomain|Population)\s", l)]
    tree = {l.split()[0]: len(l) - len(l.lstrip()) for l in rows if "Timer" not in l}
    assert "Path" in out.stdout and tree["Solve"] == 0 and tree["SweepSolver"] == 2 and tree["SweepSubdomain"] == 4 and tree["LTimes"] == 2


GEN_ZONE_FIELDS = ["sigt_zonal", "volume", "zone_to_num_mixelem", "zone_to_mixelem", "mixelem_to_zone", "mixelem_to_material",
                   "mixelem_to_fraction"]


@pytest.mark.parametrize("layout", LAYOUTS)
def test_device_generator_bitwise_equals_oracle(gpu, layout):
    """with a GPU bound, the zone loops of generateProblem (material sub-sampling, mixed-zone tables,
    sigt_zonal; Generate/Space.cpp:189-363) run in csrc/kb200_generate.cu.  Every table must be bit-identical to the oracle's
    (the device twin of tests/test_host_cpu.py::test_host_generator_bitwise_equals_oracle), in all six nestings, on a
    decomposed problem that cuts through the material interfaces, with non-default cross sections."""
    args = f"--zones 24,40,12 --groups 6 --quad 16 --legendre 1 --zset 2,2,3 --gset 2 --dset 8 --sigt 0.2,0.001,0.3 --layout {layout}"
    p, o, _, _ = pair(gpu, args)
    for n in GEN_ZONE_FIELDS:
        a, b = p.field(n), o.field(n)
        assert a.shape == b.shape and np.array_equal(a.astype(b.dtype).view(np.uint8), b.view(np.uint8)), n


def test_device_generator_matches_host_generator_at_scale(gpu, monkeypatch):
    """the same tables from the device kernels and from the host loops (KB200_HOST_GENERATOR=1) at 64^3 zones."""
    args = "--zones 64,64,64 --groups 4 --quad 8 --legendre 0 --zset 1,2,1 --gset 1 --dset 8 --layout GZD"
    dev = gpu.Problem(args)
    monkeypatch.setenv("KB200_HOST_GENERATOR", "1")
    host = gpu.Problem(args)
    for n in GEN_ZONE_FIELDS:
        assert np.array_equal(dev.field(n), host.field(n)), n


@pytest.mark.parametrize("groups", [8, 64])  # 64 groups in two sets: the one-read kernel (kb200_scatter_slab.cu)
@pytest.mark.parametrize("layout", ["DGZ", "GDZ", "GZD", "ZGD"])
def test_source_folded_into_scattering_matches_separate_kernels(gpu, monkeypatch, layout, groups):
    """SteadyStateSolver's scattering call adds Kernel::source's term in the epilogue of the tensor-core kernel
    (Kernel/Source.cpp:59-75 touches only the moment-0 slab that kernel has just written).  One iteration through the
    solver with and without the fold must leave bit-identical phi_out and rhs, and the same particle count."""
    args = f"--zones 12,20,10 --groups {groups} --quad 16 --legendre 2 --gset 2 --dset 8 --zset 1,2,1 --niter 2 --layout {layout}"
    res = {}
    for fold in ("1", "0"):
        monkeypatch.setenv("KB200_FOLD_SOURCE", fold)
        p = gpu.Problem(args)
        parts = p.solve()
        res[fold] = (parts, p.field("phi_out").copy(), p.field("rhs").copy())
        p.close()
    assert res["1"][0] == res["0"][0]
    assert float(np.max(np.abs(res["1"][1]))) > 0.0
    assert np.array_equal(res["1"][1], res["0"][1]) and np.array_equal(res["1"][2], res["0"][2])
