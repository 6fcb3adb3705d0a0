// SweepSubdomain for the element-fastest storage orders (DZG, GZD, ZDG, ZGD) on sm_100a: "pencil" kernel.
//
// Reference arithmetic: src/Kripke/Kernel/SweepSubdomain.cpp:86-108.  In these layouts the (direction, group)
// elements of one zone are adjacent in memory, so
//   * lanes of a warp  <->  32 consecutive elements in storage order (a "slice"): every rhs / psi / plane access of
//     a warp is one coalesced 256-byte row, every lane runs its own independent recurrence (no shuffles, no scan);
//   * a warp owns a PENCIL of PJ x PK = 4 x 2 zone lines and marches it along i.  Inside a step the eight zones are
//     visited in dependency order, so the j- and k-face fluxes between the lines of a pencil never leave registers and
//     the i-face flux of every line is a register carried from step to step: eight independent i recurrences per lane
//     give the fp64 pipe instruction-level parallelism without more warps;
//   * the 16 warps of a CTA form a 4 x 4 systolic array of pencils (a 16 x 8 column tile): pencil (J,K) is one step
//     behind (J-1,K) and (J,K-1).  Faces cross pencils through small shared-memory rings; there is NO CTA-wide
//     barrier in the loop: every warp publishes a monotone step counter (st.release.cta) and its two consumers
//     poll it (ld.acquire.cta), so a late load stalls one warp, not sixteen;
//   * rhs and sigt of the NEXT step are requested right after the current step's values have been consumed (one
//     register per line and array, no staging buffer): with 16 warps per SM a step of a warp recurs about every
//     microsecond, so the loads have landed when they are needed; tile-boundary faces of the next step come in by
//     cp.async into a per-warp shared-memory slot (no registers at all);
//   * work items are (subdomain, column tile, slice); the items of a launch are cut into one contiguous run per SM, and
//     a CTA streams through its run without draining the systolic array between items.  Column tiles of a subdomain
//     form a 2-D wavefront of launches (tile diagonals), all subdomains of a stage share them; tile-boundary faces use
//     the i/j/k_plane arrays in place like the reference;
//   * optionally Kernel::population's sum (src/Kripke/Kernel/Population.cpp:49-63) is accumulated while psi is in
//     registers: one partial per CTA, summed later in a fixed order by kb200_population_reduce.
// EXACT mode keeps the reference's multiply-then-add order and IEEE division (bit-identical psi); the default
// contracts to DFMA and divides by reciprocal + cubic Newton step.
#include "kb200_common.cuh"
#include <stddef.h>
#include <type_traits>
#include <vector>

namespace kb200 {

constexpr int PW = 16;                   // warps per CTA
constexpr int PWJ = 4, PWK = 4;          // pencils per tile
constexpr int PJ = 4, PK = 2;            // lines per pencil
constexpr int PL = PJ * PK;
constexpr int PTJ = PWJ * PJ, PTK = PWK * PK;  // column tile in lines
constexpr int PR = 4;                    // ring slots between neighbouring pencils
constexpr int PT = PW * 32;              // threads per CTA
#ifndef PENCIL_SLEEP_NS
#define PENCIL_SLEEP_NS 0
#endif

struct PGeom {  // kernel parameter: lives in the constant bank, costs no registers
  int layout, Ds, Gs, ni, nj, nk;
  int diag, tj_lo, ntiles;        // this launch: tile diagonal, first tile and number of tiles on it
  int nslices, nitems, nctas;     // items = (subdomain, tile of the diagonal, slice), cut into nctas contiguous runs
  int j_tiles, k_tiles;
  int d_fast;                     // d is the faster of the two element indices
  int eps;                        // elements per slice (<= 32): lanes >= eps idle.  Where the elements of a zone are NOT one
                                  // contiguous run (GZD: runs of Ds directions, DZG: runs of Gs groups) a slice holds whole runs,
                                  // so that a warp's access is one contiguous piece instead of two pieces from different pages
  int interleave;                 // items of a CTA: 0 = one contiguous run, 1 = every nctas-th item (neighbouring CTAs then work on
                                  // neighbouring slices of the same zones at the same time: contiguous DRAM bursts)
  int exp;                        // timing experiments (KB200_PENCIL_EXP bitmask; results are wrong when non-zero)
  unsigned fa, fg, fz;            // psi/rhs strides: direction, group, zone
  unsigned sg, sz;                // sigt strides: group, zone
  unsigned ipd, ipg, ipa, ipb, jpd, jpg, jpa, jpb, kpd, kpg, kpa, kpb;  // plane strides (strides_plane)
};

// mbarrier primitives (the hand-shakes between neighbouring pencils): arrive = release, wait = acquire at CTA scope, the
// waiting warp sleeps in hardware instead of spinning
__device__ __forceinline__ void p_mb_init(unsigned addr, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void p_mb_arrive(unsigned addr) {
  asm volatile("{ .reg .b64 t; mbarrier.arrive.shared::cta.b64 t, [%0]; }" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void p_mb_wait(unsigned addr, unsigned parity) {
  asm volatile("{\n .reg .pred p;\n PMW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @!p bra PMW;\n}" ::"r"(addr), "r"(parity) : "memory");
}
__device__ __forceinline__ double p_ldg(const double *p) {  // streamed once: do not keep in L1
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
// (measured dead end: an L2 evict-first policy on the rhs loads and psi stores, meant to keep sigt and the face planes in
// L2, made the kernel slower -- 15.8 -> 16.7 ms in ZGD)
__device__ __forceinline__ double p_ld_cg(const double *p) {  // faces written by other CTAs / earlier launches: L2
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void p_cp_async8(unsigned smem_dst, const void *gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void p_cp_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ double p_lds(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void p_sts(unsigned addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }

// per-warp, warp-uniform description of the current (subdomain, tile): kept in shared memory, read by broadcast
struct PTile {
  kb200_sweep_desc ds;
  int jz[PJ], kz[PK];      // memory j / k index of the pencil's lines (clamped to the tile)
  const double *vol;       // zone volumes of the subdomain (population sum)
  double vol_c;            // > 0: every zone has this volume
};
// Everything a warp touches in shared memory sits in ONE block per warp, so that a single register (block base + 8 * lane)
// addresses all of it with immediate offsets, and the blocks of the upwind pencils are a constant distance away.
constexpr unsigned OJ = 0;                                  // [PR][PK][32] doubles: j faces leaving the pencil
constexpr unsigned OK = OJ + PR * PK * 256;                 // [PR][PJ][32] doubles: k faces leaving the pencil
constexpr unsigned OB = OK + PR * PJ * 256;                 // [2][PK + PJ][32] doubles: tile-boundary faces entering it
constexpr unsigned OC = OB + 2 * (PK + PJ) * 256;           // [2 + PJ + PK][32] doubles: 2*cos/delta of every lane's direction
constexpr unsigned OP = OC + (2 + PJ + PK) * 256;           // [4][32] 8-byte slots: i/j/k plane element offset, direction
constexpr unsigned OS = OP + 4 * 256;                       // [32] doubles: running population sum of every lane
constexpr unsigned OT = OS + 256;                           // PTile
constexpr unsigned OM = OT + ((sizeof(PTile) + 255) / 256) * 256;  // mbarriers: full_j[PR], full_k[PR], empty_j[PR], empty_k[PR]
constexpr unsigned OMFJ = OM, OMFK = OM + 8 * PR, OMEJ = OM + 16 * PR, OMEK = OM + 24 * PR;
constexpr unsigned WBSZ = OM + 256;

enum : unsigned {
  F_JLAST = 1u, F_KLAST = 2u,      // the tile's last j / k line lies in this pencil: outgoing faces go to the plane arrays
  F_JB = 4u, F_KB = 8u,            // incoming j / k faces come from the plane arrays (tile boundary, not vacuum)
  F_JOUT = 16u, F_KOUT = 32u,      // ... and also to the downwind subdomain's plane chunk
  F_IZERO = 64u, F_UNIX = 128u, F_FWD = 256u, F_VOLU = 512u, F_VALID = 1024u,
  F_JRING = 2048u, F_KRING = 4096u, F_JNEXT = 8192u, F_KNEXT = 16384u  // neighbours inside the CTA
};

__device__ __forceinline__ unsigned p_lds32(unsigned addr) {
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void p_sts32(unsigned addr, unsigned v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

template <bool EXACT, bool POP>
__global__ void __launch_bounds__(PT, 1) sweep_pencil_kernel(const kb200_sweep_desc *__restrict__ descs, const __grid_constant__ PGeom gm,
                                                           const double *const *__restrict__ pop_w,
                                                           const double *const *__restrict__ pop_vol,
                                                           const double *__restrict__ pop_vol_const,
                                                           double *__restrict__ pop_partial) {
  extern __shared__ __align__(16) unsigned char psm[];   // [PW] warp blocks, then [PW] doubles for the final reduction
#define TL (*reinterpret_cast<PTile *>(psm + (threadIdx.x >> 5) * WBSZ + OT))
  const int lane = threadIdx.x & 31;
  const int ni = gm.ni;
  const unsigned mb = (unsigned)__cvta_generic_to_shared(psm) + (threadIdx.x >> 5) * WBSZ + 8u * lane;  // block base + lane slot

  // ring hand-shakes: the 32 lanes of the producer arrive on "full" after their stores, the 32 lanes of the consumer on
  // "empty" after their loads.  Every pair of neighbouring warps shakes hands on EVERY local step (also when one of the
  // pencils lies outside a ragged tile), so slot and phase follow from the local step number alone.
  if (lane < 4 * PR) p_mb_init(mb - 8u * lane + OM + 8u * lane, 32u);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  // ---- this CTA's run of items ----
  // `item` counts this CTA's items; ITEM(k) is the k-th one's index in the launch's (subdomain, tile, slice) list
  const int il = gm.interleave;
  const int item_hi = il ? (gm.nitems - (int)blockIdx.x + gm.nctas - 1) / gm.nctas
                         : (int)(((long long)(blockIdx.x + 1) * gm.nitems) / gm.nctas);
  int item = il ? 0 : (int)(((long long)blockIdx.x * gm.nitems) / gm.nctas);
#define ITEM(k) (il ? (int)blockIdx.x + (k) * gm.nctas : (k))
  const int T = (item_hi - item) * ni;               // local steps of every warp
  const unsigned fz = gm.fz, sz = gm.sz;

  // ---- state of the current item ----
  int pair = -1;            // (subdomain, tile) of the current item
  unsigned fl = 0;          // F_* flags
  unsigned lmask = 0;       // bit l = jl * PK + kl: the line exists in this (possibly ragged) tile
  unsigned zrow[PL];        // first zone of the line: (kz * nj + jz) * ni
  const double *__restrict__ rhs_b = nullptr;
  const double *__restrict__ sigt_b = nullptr;
  double *__restrict__ psi_b = nullptr;
  unsigned b = 0, sb = 0;   // psi/rhs and sigt element offset of this lane at the current zone column (without the line part)
  unsigned iz = 0;          // memory i index of the current step
  int i = 0;
  double acc = 0.0;         // population: sum of psi (x volume) over the current slice; the total lives in shared memory
  if (POP) p_sts(mb + OS, 0.0);
  double R[PL], S[PL], FI[PL];
#pragma unroll
  for (int l = 0; l < PL; ++l) { R[l] = 0.0; S[l] = 1.0; FI[l] = 0.0; zrow[l] = 0u; }

  // element offsets of slice `sl` for this lane
  auto slice_elem = [&](int sl, unsigned &eoff, unsigned &soff, unsigned &ipo, unsigned &jpo, unsigned &kpo, int &d) {
    int e = sl * gm.eps + min(lane, gm.eps - 1);  // idle lanes shadow the slice's last element: no extra sectors
    const bool v = lane < gm.eps && e < gm.Ds * gm.Gs;
    e = min(e, gm.Ds * gm.Gs - 1);
    int g;
    if (gm.d_fast) { d = e % gm.Ds; g = e / gm.Ds; }
    else { g = e % gm.Gs; d = e / gm.Gs; }
    eoff = (unsigned)d * gm.fa + (unsigned)g * gm.fg;
    soff = (unsigned)g * gm.sg;
    ipo = (unsigned)d * gm.ipd + (unsigned)g * gm.ipg;
    jpo = (unsigned)d * gm.jpd + (unsigned)g * gm.jpg;
    kpo = (unsigned)d * gm.kpd + (unsigned)g * gm.kpg;
    return v;
  };
  // (subdomain, tile) changed: descriptor, lines of the pencil, flags.  Warp-uniform; rare.
  auto tile_setup = [&](int pr) {
    pair = pr;
    const int J = (threadIdx.x >> 5) & (PWJ - 1), K = (threadIdx.x >> 5) / PWJ;
    const int sd = pr / gm.ntiles, tj = gm.tj_lo + (pr - sd * gm.ntiles), tk = gm.diag - tj;
    PTile &tl = TL;
    __syncwarp();
    for (int w = lane; w < (int)(sizeof(kb200_sweep_desc) / sizeof(int)); w += 32)
      reinterpret_cast<int *>(&tl.ds)[w] = reinterpret_cast<const int *>(&descs[sd])[w];
    __syncwarp();
    const int tjn = min(PTJ, gm.nj - tj * PTJ), tkn = min(PTK, gm.nk - tk * PTK);
    const int jd = tl.ds.jd, kd = tl.ds.kd;
    int jz[PJ], kz[PK];
#pragma unroll
    for (int jl = 0; jl < PJ; ++jl) {
      const int jlz = tj * PTJ + min(J * PJ + jl, tjn - 1);
      jz[jl] = (jd > 0) ? jlz : gm.nj - 1 - jlz;
    }
#pragma unroll
    for (int kl = 0; kl < PK; ++kl) {
      const int klz = tk * PTK + min(K * PK + kl, tkn - 1);
      kz[kl] = (kd > 0) ? klz : gm.nk - 1 - klz;
    }
    if (lane == 0) {
#pragma unroll
      for (int jl = 0; jl < PJ; ++jl) tl.jz[jl] = jz[jl];
#pragma unroll
      for (int kl = 0; kl < PK; ++kl) tl.kz[kl] = kz[kl];
      if (POP) {
        tl.vol = pop_vol[sd];
        tl.vol_c = pop_vol_const ? pop_vol_const[sd] : 0.0;
      }
    }
    lmask = 0;
#pragma unroll
    for (int jl = 0; jl < PJ; ++jl)
#pragma unroll
      for (int kl = 0; kl < PK; ++kl) {
        const int l = jl * PK + kl;
        if (J * PJ + jl < tjn && K * PK + kl < tkn) lmask |= 1u << l;
        zrow[l] = (unsigned)((kz[kl] * gm.nj + jz[jl]) * ni);
      }
    fl = 0;
    if ((tjn - 1) / PJ == J && !(gm.exp & 2)) fl |= F_JLAST;
    if ((tkn - 1) / PK == K && !(gm.exp & 2)) fl |= F_KLAST;
    if (J > 0) fl |= F_JRING;
    if (K > 0) fl |= F_KRING;
    if (J < PWJ - 1) fl |= F_JNEXT;
    if (K < PWK - 1) fl |= F_KNEXT;
    if (J == 0 && !(tl.ds.inflow_zero[1] != 0 && tj == 0) && !(gm.exp & 1)) fl |= F_JB;
    if (K == 0 && !(tl.ds.inflow_zero[2] != 0 && tk == 0) && !(gm.exp & 1)) fl |= F_KB;
    if (tj == gm.j_tiles - 1 && tl.ds.out_plane[1] != nullptr) fl |= F_JOUT;
    if (tk == gm.k_tiles - 1 && tl.ds.out_plane[2] != nullptr) fl |= F_KOUT;
    if (tl.ds.inflow_zero[0] != 0) fl |= F_IZERO;
    if (tl.ds.id > 0) fl |= F_FWD;
    rhs_b = tl.ds.rhs; sigt_b = tl.ds.sigt; psi_b = tl.ds.psi;
    bool uni = true;
    for (int ii = 1; ii < ni; ++ii) uni = uni && (tl.ds.dx[ii] == tl.ds.dx[0]);
    if (uni) fl |= F_UNIX;
    if (POP && pop_vol_const && pop_vol_const[sd] > 0.0) fl |= F_VOLU;
    __syncwarp();
  };
  // a new slice of the current tile: element offsets, coefficients (SweepSubdomain.cpp:88-93), incoming i faces
  auto slice_setup = [&](int sl) {
    unsigned eoff, soff, ipo, jpo, kpo;
    int d;
    PTile &tl = TL;
    fl = (slice_elem(sl, eoff, soff, ipo, jpo, kpo, d) && !(gm.exp & 4)) ? (fl | F_VALID) : (fl & ~F_VALID);
    iz = (fl & F_FWD) ? 0u : (unsigned)(ni - 1);
    i = 0;
    b = eoff + iz * fz; sb = soff + iz * sz;
    p_sts32(mb + OP, ipo); p_sts32(mb + OP + 256, jpo); p_sts32(mb + OP + 512, kpo); p_sts32(mb + OP + 768, (unsigned)d);
    const double tx = 2.0 * tl.ds.xcos[d];
    p_sts(mb + OC, tx / tl.ds.dx[0]);
    p_sts(mb + OC + 256, tx);
    const double ty = 2.0 * tl.ds.ycos[d], tz = 2.0 * tl.ds.zcos[d];
#pragma unroll
    for (int jl = 0; jl < PJ; ++jl) p_sts(mb + OC + (2 + jl) * 256, ty / tl.ds.dy[tl.jz[jl]]);
#pragma unroll
    for (int kl = 0; kl < PK; ++kl) p_sts(mb + OC + (2 + PJ + kl) * 256, tz / tl.ds.dz[tl.kz[kl]]);
#pragma unroll
    for (int jl = 0; jl < PJ; ++jl)
#pragma unroll
      for (int kl = 0; kl < PK; ++kl) {
        const int l = jl * PK + kl;
        FI[l] = 0.0;
        if (!(fl & F_IZERO) && (lmask >> l & 1u))
          FI[l] = p_ld_cg(tl.ds.i_plane + (ipo + (unsigned)tl.jz[jl] * gm.ipa + (unsigned)tl.kz[kl] * gm.ipb));
      }
  };
  // rhs and sigt of the zone column at offsets (bb, sbb)
  auto request_operands = [&](unsigned bb, unsigned sbb) {
#pragma unroll
    for (int l = 0; l < PL; ++l)
      if (lmask >> l & 1u) {
        R[l] = p_ldg(rhs_b + (zrow[l] * fz + bb));
        S[l] = __ldg(sigt_b + (zrow[l] * sz + sbb));
      }
  };
  // tile-boundary faces of zone column izz into boundary slot `slot`
  auto request_boundary = [&](unsigned jpo, unsigned kpo, unsigned izz, unsigned slot) {
    PTile &tl = TL;
    const unsigned dst = mb + OB + slot * ((PK + PJ) * 256);
    if (fl & F_JB) {
#pragma unroll
      for (int kl = 0; kl < PK; ++kl)
        if (lmask >> kl & 1u) p_cp_async8(dst + kl * 256, tl.ds.j_plane + (jpo + izz * gm.jpa + (unsigned)tl.kz[kl] * gm.jpb));
    }
    if (fl & F_KB) {
#pragma unroll
      for (int jl = 0; jl < PJ; ++jl)
        if (lmask >> (jl * PK) & 1u)
          p_cp_async8(dst + (PK + jl) * 256, tl.ds.k_plane + (kpo + izz * gm.kpa + (unsigned)tl.jz[jl] * gm.kpb));
    }
  };
  // start of an item that does not continue the previous one (first item, or the tile changed): everything from scratch
  auto item_start = [&](int k) {
    const int it = ITEM(k);
    const int pr = it / gm.nslices, sl = it - pr * gm.nslices;
    if (pr != pair) tile_setup(pr);
    slice_setup(sl);
    if (lmask) {
      request_operands(b, sb);
      request_boundary(p_lds32(mb + OP + 256), p_lds32(mb + OP + 512), iz, 0u);
    }
  };

  if (T > 0) item_start(item);

  // one local step at global step H (even H read boundary slot 0 ... the slot alternates with the LOCAL step, see bsl)
  // FULL: every line of the pencil exists (no per-line tests, long basic blocks for the scheduler)
  auto step = [&](auto full_tag, auto volu_tag, const int t, const unsigned bslot) {
    constexpr bool FULL = decltype(full_tag)::value;
    constexpr bool VOLU = decltype(volu_tag)::value;  // population: every zone has the same volume (no volume loads)
    const unsigned ws = (unsigned)t & (PR - 1), rs = ws;  // ring slot of this local step (same number on both sides)
    const unsigned ph = ((unsigned)t / PR) & 1u;           // phase parity of the slot's current use
    const bool last_i = (i == ni - 1);
    const bool fwd = (fl & F_FWD) != 0, valid = (fl & F_VALID) != 0;
    // where the next local step reads: next zone column of this slice, or the first one of the next slice of this tile
    bool more = true;
    unsigned nb, nsb, njp = 0, nkp = 0, izn;
    if (!last_i) {
      izn = fwd ? iz + 1u : iz - 1u;
      nb = fwd ? b + fz : b - fz; nsb = fwd ? sb + sz : sb - sz;
      if (fl & (F_JB | F_KB)) { njp = p_lds32(mb + OP + 256); nkp = p_lds32(mb + OP + 512); }
    } else {
      izn = fwd ? 0u : (unsigned)(ni - 1);
      const int nit = ITEM(item + 1);
      const int npr = nit / gm.nslices;
      more = item + 1 < item_hi && npr == pair;
      unsigned eo = 0, so = 0, ipo;
      int d;
      if (more) slice_elem(nit - npr * gm.nslices, eo, so, ipo, njp, nkp, d);
      nb = eo + izn * fz; nsb = so + izn * sz;
      // the operand requests below are unconditional (a predicated load would have to merge with the old register value,
      // which costs a copy that waits for the load): with nothing to fetch they re-read the current zone column
      if (!more) { nb = b; nsb = sb; }
    }
    const unsigned wb = mb - 8u * lane;  // this warp's block
    if (fl & (F_JB | F_KB)) { p_cp_async_wait(); __syncwarp(); }

    double cxu = p_lds(mb + OC);
    if (!(fl & F_UNIX)) cxu = __ddiv_rn(p_lds(mb + OC + 256), TL.ds.dx[iz]);
    const unsigned bsl = mb + OB + bslot * ((PK + PJ) * 256);

    // The pencil is worked through one k line (PJ zones) at a time, each in two phases.  Phase A touches nothing that
    // depends on another zone of this step -- reciprocal of the denominator, the rhs and i-face part of psi, the k-face
    // part (the k faces come from the ring or from the previous k line) -- so its PJ zones interleave freely in the
    // fp64 pipe; phase B is the j recurrence, two dependent DFMAs per zone:
    //   psi = (rhs + fi*cx)*rc + fk*(cz*rc) + fj*(cy*rc),   face' = 2*psi - face.
    // EXACT mode keeps the reference's expression and order instead (one zone after the other).
    double fko[PJ];  // k faces: entering the current k line, then leaving it
    if (fl & F_KRING) {  // pencil (J,K-1) has filled the slot; hand it back as soon as it is read
      p_mb_wait(wb - PWJ * WBSZ + OMFK + 8u * ws, ph);
#pragma unroll
      for (int jl = 0; jl < PJ; ++jl) fko[jl] = p_lds(mb - PWJ * WBSZ + OK + (rs * PJ + jl) * 256);
      p_mb_arrive(wb - PWJ * WBSZ + OMEK + 8u * ws);
    } else {
#pragma unroll
      for (int jl = 0; jl < PJ; ++jl) fko[jl] = (fl & F_KB) ? p_lds(bsl + (PK + jl) * 256) : 0.0;
    }
    double fji[PK];  // j faces entering the pencil
    if (fl & F_JRING) {
      p_mb_wait(wb - WBSZ + OMFJ + 8u * ws, ph);
#pragma unroll
      for (int kl = 0; kl < PK; ++kl) fji[kl] = p_lds(mb - WBSZ + OJ + (rs * PK + kl) * 256);
      p_mb_arrive(wb - WBSZ + OMEJ + 8u * ws);
    } else {
#pragma unroll
      for (int kl = 0; kl < PK; ++kl) fji[kl] = (fl & F_JB) ? p_lds(bsl + kl * 256) : 0.0;
    }
    double fjo[PK];  // j faces leaving the pencil
#pragma unroll
    for (int kl = 0; kl < PK; ++kl) {
      if (!FULL && !(lmask >> kl & 1u)) { fjo[kl] = 0.0; continue; }
      const double czk = p_lds(mb + OC + (2 + PJ + kl) * 256);
      double fj = fji[kl];
      if (EXACT) {
#pragma unroll
        for (int jl = 0; jl < PJ; ++jl) {
          const int l = jl * PK + kl;
          if (!FULL && !(lmask >> l & 1u)) continue;
          const double cyj = p_lds(mb + OC + (2 + jl) * 256);
          const double r = R[l], st = S[l], fi = FI[l], fk = fko[jl];
          R[l] = p_ldg(rhs_b + (zrow[l] * fz + nb));  // operands of the next step of this line
          S[l] = __ldg(sigt_b + (zrow[l] * sz + nsb));
          const double den = __dadd_rn(__dadd_rn(__dadd_rn(cxu, cyj), czk), st);
          double num = __dadd_rn(r, __dmul_rn(fi, cxu));
          num = __dadd_rn(num, __dmul_rn(fj, cyj));
          num = __dadd_rn(num, __dmul_rn(fk, czk));
          const double p = __ddiv_rn(num, den);
          const double p2 = 2.0 * p;
          FI[l] = p2 - fi; fj = p2 - fj; fko[jl] = p2 - fk;
          if (valid) {
            psi_b[zrow[l] * fz + b] = p;
            if (POP) {
              if (VOLU) acc += p;
              else acc = fma(p, __ldg(TL.vol + (zrow[l] + iz)), acc);
            }
          }
        }
      } else {
        double q[PJ], cyr[PJ];
#pragma unroll
        for (int jl = 0; jl < PJ; ++jl) {  // phase A
          const int l = jl * PK + kl;
          q[jl] = 0.0; cyr[jl] = 0.0;
          if (!FULL && !(lmask >> l & 1u)) continue;
          const double cyj = p_lds(mb + OC + (2 + jl) * 256);
          const double den = __dadd_rn(__dadd_rn(__dadd_rn(cxu, cyj), czk), S[l]);
          double y;
          asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(den));
          const double e1 = fma(-den, y, 1.0);  // cubic step: y*(1 + e + e^2), e ~ 2^-23 -> error ~ 2^-69
          const double e2 = fma(e1, e1, e1);
          const double rc = fma(y, e2, y);
          const double a = fma(FI[l], cxu, R[l]) * rc;
          R[l] = p_ldg(rhs_b + (zrow[l] * fz + nb));  // operands of the next step of this line
          S[l] = __ldg(sigt_b + (zrow[l] * sz + nsb));
          q[jl] = fma(fko[jl], czk * rc, a);
          cyr[jl] = cyj * rc;
        }
#pragma unroll
        for (int jl = 0; jl < PJ; ++jl) {  // phase B
          const int l = jl * PK + kl;
          if (!FULL && !(lmask >> l & 1u)) continue;
          const double p = fma(fj, cyr[jl], q[jl]);
          fj = fma(2.0, p, -fj);
          FI[l] = fma(2.0, p, -FI[l]);
          fko[jl] = fma(2.0, p, -fko[jl]);
          if (valid) {
            psi_b[zrow[l] * fz + b] = p;
            if (POP) {
              if (VOLU) acc += p;
              else acc = fma(p, __ldg(TL.vol + (zrow[l] + iz)), acc);
            }
          }
        }
      }
      fjo[kl] = fj;  // belongs to the last existing j line
    }
    // k faces leaving the pencil (fko belongs to the last existing k line of every j line)
    if (fl & F_KLAST) {
      if (valid) {
        PTile &tl = TL;
        const unsigned kp = p_lds32(mb + OP + 512) + iz * gm.kpa;
#pragma unroll
        for (int jl = 0; jl < PJ; ++jl)
          if (lmask >> (jl * PK) & 1u) {
            const unsigned ix = kp + (unsigned)tl.jz[jl] * gm.kpb;
            tl.ds.k_plane[ix] = fko[jl];
            if (fl & F_KOUT) tl.ds.out_plane[2][ix] = fko[jl];
          }
      }
    }
    if (fl & F_KNEXT) {  // pencil (J,K+1) has emptied the slot's previous contents (PR steps ago)
      if (t >= PR) p_mb_wait(wb + OMEK + 8u * ws, ph ^ 1u);
#pragma unroll
      for (int jl = 0; jl < PJ; ++jl) p_sts(mb + OK + (ws * PJ + jl) * 256, fko[jl]);
      p_mb_arrive(wb + OMFK + 8u * ws);
    }
    // j faces leaving the pencil
    if (fl & F_JLAST) {
      if (valid) {
        PTile &tl = TL;
        const unsigned jp = p_lds32(mb + OP + 256) + iz * gm.jpa;
#pragma unroll
        for (int kl = 0; kl < PK; ++kl)
          if (lmask >> kl & 1u) {
            const unsigned ix = jp + (unsigned)tl.kz[kl] * gm.jpb;
            tl.ds.j_plane[ix] = fjo[kl];
            if (fl & F_JOUT) tl.ds.out_plane[1][ix] = fjo[kl];
          }
      }
    }
    if (fl & F_JNEXT) {
      if (t >= PR) p_mb_wait(wb + OMEJ + 8u * ws, ph ^ 1u);
#pragma unroll
      for (int kl = 0; kl < PK; ++kl) p_sts(mb + OJ + (ws * PK + kl) * 256, fjo[kl]);
      p_mb_arrive(wb + OMFJ + 8u * ws);
    }
    if (last_i && valid) {  // i faces leaving the subdomain
      PTile &tl = TL;
      const unsigned ip = p_lds32(mb + OP);
#pragma unroll
      for (int jl = 0; jl < PJ; ++jl)
#pragma unroll
        for (int kl = 0; kl < PK; ++kl) {
          const int l = jl * PK + kl;
          if (lmask >> l & 1u) {
            const unsigned ix = ip + (unsigned)tl.jz[jl] * gm.ipa + (unsigned)tl.kz[kl] * gm.ipb;
            tl.ds.i_plane[ix] = FI[l];
            if (tl.ds.out_plane[0]) tl.ds.out_plane[0][ix] = FI[l];
          }
        }
    }
    // tile-boundary faces of the next step (the slot read in this step is the other one)
    if (more && (fl & (F_JB | F_KB))) request_boundary(njp, nkp, izn, bslot ^ 1u);
    // ---- advance the local schedule ----
    if (last_i) {
      if (POP) {
        if (valid) {
          const double vc = TL.vol_c;
          p_sts(mb + OS, fma(pop_w[pair / gm.ntiles][p_lds32(mb + OP + 768)] * (vc > 0.0 ? vc : 1.0), acc, p_lds(mb + OS)));
        }
        acc = 0.0;
      }
      ++item;
      if (item < item_hi) {
        if (more) slice_setup(ITEM(item) - pair * gm.nslices);
        else item_start(item);  // another tile or subdomain
      }
      return more ? (bslot ^ 1u) : 0u;
    }
    ++i; iz = izn; b = nb; sb = nsb;
    return bslot ^ 1u;
  };

  // Every warp runs its T local steps on its own; the hand-shakes alone keep pencil (J,K) behind (J-1,K) and (J,K-1).
  unsigned bslot = 0;  // boundary slot holding the current step's faces
#pragma unroll 1
  for (int t = 0; t < T; ++t) {
    if (lmask == (1u << PL) - 1u) {
      if (!POP || (fl & F_VOLU)) bslot = step(std::true_type{}, std::true_type{}, t, bslot);
      else bslot = step(std::true_type{}, std::false_type{}, t, bslot);
    } else if (lmask != 0u) {
      if (!POP || (fl & F_VOLU)) bslot = step(std::false_type{}, std::true_type{}, t, bslot);
      else bslot = step(std::false_type{}, std::false_type{}, t, bslot);
    }
    else {  // a pencil outside a ragged tile: keep the hand-shakes and the item bookkeeping going
      const unsigned ws = (unsigned)t & (PR - 1), ph = ((unsigned)t / PR) & 1u, wb = mb - 8u * lane;
      if (fl & F_KRING) { p_mb_wait(wb - PWJ * WBSZ + OMFK + 8u * ws, ph); p_mb_arrive(wb - PWJ * WBSZ + OMEK + 8u * ws); }
      if (fl & F_JRING) { p_mb_wait(wb - WBSZ + OMFJ + 8u * ws, ph); p_mb_arrive(wb - WBSZ + OMEJ + 8u * ws); }
      if (fl & F_KNEXT) { if (t >= PR) p_mb_wait(wb + OMEK + 8u * ws, ph ^ 1u); p_mb_arrive(wb + OMFK + 8u * ws); }
      if (fl & F_JNEXT) { if (t >= PR) p_mb_wait(wb + OMEJ + 8u * ws, ph ^ 1u); p_mb_arrive(wb + OMFJ + 8u * ws); }
      if (i == ni - 1) {
        ++item;
        i = 0;
        if (item < item_hi) {
          const int pr = ITEM(item) / gm.nslices;
          if (pr != pair) { item_start(item); bslot = 0; }
        }
      } else {
        ++i;
      }
    }
  }
  p_cp_async_wait();

  if (POP) {  // fixed-order block reduction: lanes, then warps
    double *red = reinterpret_cast<double *>(psm + PW * WBSZ);
    double pop = warp_sum(p_lds(mb + OS));
    if (lane == 0) red[threadIdx.x >> 5] = pop;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int w = 0; w < PW; ++w) s += red[w];
      pop_partial[blockIdx.x] = s;
    }
  }
#undef TL
#undef ITEM
}

}  // namespace kb200

using namespace kb200;

// Returns 0 if the batch was handled, -1 if this path does not apply (caller falls back), >0 on error.
// pop_* (optional): device pointer tables of n entries (+ per-subdomain uniform volume, 0 = read the volume array) and a
// scratch of pop_capacity doubles for the fused population partials; *pop_count receives the number written.
int kb200_sweep_pencil_try(const kb200_sweep_desc *h, int n, const void *d_descs, const double *const *d_pop_w,
                           const double *const *d_pop_vol, const double *d_pop_vol_const, double *d_pop_partial, int pop_capacity,
                           int *pop_count, cudaStream_t st) {
  const int layout = h[0].layout;
  if (pop_count) *pop_count = 0;
  if (layout != 1 && layout != 3 && layout != 4 && layout != 5) return -1;
  const char *env = getenv("KB200_SWEEP_PENCIL");
  if (env && env[0] == '0') return -1;
  PGeom gm;
  memset(&gm, 0, sizeof(gm));
  gm.layout = layout; gm.Ds = h[0].Ds; gm.Gs = h[0].Gs; gm.ni = h[0].ni; gm.nj = h[0].nj; gm.nk = h[0].nk;
  if ((double)gm.Ds * gm.Gs * gm.ni * gm.nj * gm.nk >= 2147483648.0) return -1;  // 32-bit element offsets
  { const char *x = getenv("KB200_PENCIL_EXP"); gm.exp = x ? atoi(x) : 0; }
  // measured at BASELINE config 2: ZGD 16.0 -> 14.8 ms with interleaved items (a zone's 24 slices are one 6 KB run that
  // 24 neighbouring CTAs then stream together), GZD 18.9 -> 19.1 ms (its slices are 12 KB apart anyway)
  gm.interleave = (layout == 4 || layout == 5) ? 1 : 0;
  { const char *x = getenv("KB200_PENCIL_INTERLEAVE"); if (x) gm.interleave = (x[0] == '1') ? 1 : 0; }
  gm.j_tiles = (gm.nj + PTJ - 1) / PTJ;
  gm.k_tiles = (gm.nk + PTK - 1) / PTK;
  const int E = gm.Ds * gm.Gs;
  gm.eps = 32;
  if (layout == 1 || layout == 3) {  // DZG: runs of Gs groups, GZD: runs of Ds directions
    const int run = (layout == 3) ? gm.Ds : gm.Gs;
    const char *ae = getenv("KB200_PENCIL_ALIGN");
    if (run < 32 && !(ae && ae[0] == '0')) gm.eps = (32 / run) * run;
  }
  gm.nslices = (E + gm.eps - 1) / gm.eps;
  {
    const long long Zs = (long long)gm.ni * gm.nj * gm.nk;
    const Strides3 fs = strides_dgz(layout, gm.Ds, gm.Gs, Zs);
    const StridesP ips = strides_plane(layout, gm.Ds, gm.Gs, gm.nj, gm.nk), jps = strides_plane(layout, gm.Ds, gm.Gs, gm.ni, gm.nk),
                   kps = strides_plane(layout, gm.Ds, gm.Gs, gm.ni, gm.nj);
    long long sg_l, sz_l;
    strides_gz(layout, gm.Gs, Zs, sg_l, sz_l);
    gm.d_fast = fs.a < fs.g ? 1 : 0;
    gm.fa = (unsigned)fs.a; gm.fg = (unsigned)fs.g; gm.fz = (unsigned)fs.z;
    gm.sg = (unsigned)sg_l; gm.sz = (unsigned)sz_l;
    gm.ipd = (unsigned)ips.d; gm.ipg = (unsigned)ips.g; gm.ipa = (unsigned)ips.a; gm.ipb = (unsigned)ips.b;
    gm.jpd = (unsigned)jps.d; gm.jpg = (unsigned)jps.g; gm.jpa = (unsigned)jps.a; gm.jpb = (unsigned)jps.b;
    gm.kpd = (unsigned)kps.d; gm.kpg = (unsigned)kps.g; gm.kpa = (unsigned)kps.a; gm.kpb = (unsigned)kps.b;
  }
  const size_t smem = (size_t)PW * WBSZ + PW * sizeof(double);
  const int sms = sm_count();
  const int ndiag = gm.j_tiles + gm.k_tiles - 1;
  long long total_ctas = 0;
  std::vector<int> nct(ndiag);
  for (int diag = 0; diag < ndiag; ++diag) {
    const int tj_lo = diag - (gm.k_tiles - 1) > 0 ? diag - (gm.k_tiles - 1) : 0;
    const int tj_hi = diag < gm.j_tiles - 1 ? diag : gm.j_tiles - 1;
    const long long items = (long long)(tj_hi - tj_lo + 1) * n * gm.nslices;
    nct[diag] = (int)(items < sms ? items : sms);
    total_ctas += nct[diag];
  }
  double *pp = d_pop_partial;
  if (pp && (!d_pop_w || !d_pop_vol || total_ctas > pop_capacity)) pp = nullptr;
  const bool exact = exact_mode();
  auto kern = pp ? (exact ? sweep_pencil_kernel<true, true> : sweep_pencil_kernel<false, true>)
                 : (exact ? sweep_pencil_kernel<true, false> : sweep_pencil_kernel<false, false>);
  KB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long used = 0;
  for (int diag = 0; diag < ndiag; ++diag) {
    gm.diag = diag;
    gm.tj_lo = diag - (gm.k_tiles - 1) > 0 ? diag - (gm.k_tiles - 1) : 0;
    const int tj_hi = diag < gm.j_tiles - 1 ? diag : gm.j_tiles - 1;
    gm.ntiles = tj_hi - gm.tj_lo + 1;
    gm.nitems = gm.ntiles * n * gm.nslices;
    gm.nctas = nct[diag];
    kern<<<gm.nctas, PT, smem, st>>>((const kb200_sweep_desc *)d_descs, gm, d_pop_w, d_pop_vol, d_pop_vol_const, pp ? pp + used : nullptr);
    int rc = post_launch("sweep_pencil");
    if (rc) return rc;
    used += gm.nctas;
  }
  if (pop_count) *pop_count = pp ? (int)used : 0;
  return 0;
}
