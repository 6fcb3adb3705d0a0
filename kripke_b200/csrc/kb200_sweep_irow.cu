// SweepSubdomain for the zone-fastest storage orders (DGZ, GDZ) on sm_100a, default arithmetic:
// i-rows across lanes, the i recurrence as a warp-level affine scan, k-planes as a warp pipeline.
//
// Reference arithmetic: src/Kripke/Kernel/SweepSubdomain.cpp:86-108.  With A = cx/den and
// B = (rhs + fj*cy + fk*cz)/den the update of one zone is  psi = fi*A + B,  fi' = 2*psi - fi
// = fi*(2A-1) + 2B : along i the face flux obeys an AFFINE recurrence whose coefficients do not
// depend on fi.  Affine maps compose associatively, so a whole i-row can be advanced by a
// log-depth scan instead of ni dependent steps.  That turns the memory-unfriendly mapping of
// kb200_sweep_zline.cu (one thread per i-row, 32 different rows per request) into the natural one:
//   * LR = ni/4 consecutive lanes hold one i-row of one phase-space element (d,g), four zones per
//     lane: rhs, sigt and psi move as fully coalesced 256-bit accesses (ni*8 contiguous bytes per
//     row); a warp carries ER = 32/LR rows of ER different elements;
//   * a warp walks j = 0..nj-1 of ONE k-plane, so the j-face flux of a row is simply the
//     register copy left by the previous step of the same lane (no shuffle, no memory);
//   * the NW warps of a CTA hold NW consecutive k-planes.  Warp w hands the k-face flux of a row
//     to warp w+1 through a small shared-memory ring guarded by mbarriers (full/empty per slot).
//     There is NO CTA-wide barrier in the loop: the planes settle into a skewed pipeline on their
//     own, and a late load stalls one warp instead of all of them.  k tiles (nk > NW) and
//     elements follow each other in the same CTA without draining the pipeline; tile-boundary k
//     faces go through the k_plane array (the last warp publishes a row counter, warp 0 waits on it);
//   * inside a row: the reciprocals and the "A" half of every zone's map are computed BEFORE the
//     warp waits for its upwind k faces; then each lane composes its four zone maps, a
//     Kogge-Stone scan over the LR lanes (log2(LR) steps of 64-bit shuffles) gives every lane its
//     incoming i-face flux, and the four zones are finished locally;
//   * everything a row needs from global memory (rhs rows, sigt rows, tile-boundary k faces, zone
//     volumes for the fused population sum) is brought in ahead of time by TMA bulk copies
//     (cp.async.bulk, completion counted in bytes on a per-slot mbarrier) into a per-warp staging
//     ring: no registers held by loads, and none of the per-lane cp.async (LDGSTS) traffic that
//     saturated first.  Rows j and j+1 of a k-plane are adjacent in memory, so with an even nj one
//     copy per array brings in a row PAIR and the k-face ring is handed over once per pair.
// The scan re-associates the i recurrence and 2*cos/delta is formed as (2*cos)*(1/delta), so
// results differ from the reference in the last bits (|difference| ~ 1e-16 relative; contractions
// |2A-1| < 1 keep the recurrence stable).  EXACT mode and shapes this kernel does not cover use
// kb200_sweep_zline.cu / kb200_sweep.cu.
// When pop_partial is non-null the kernel also accumulates Kernel::population's sum
// (w(d)*psi)*volume(z) (src/Kripke/Kernel/Population.cpp:49-63) while psi is still in registers.
#include "kb200_common.cuh"
#include <cuda.h>
#include <type_traits>
#include <vector>

namespace kb200 {

constexpr int IROW_RING = 2;   // slots of the k-face ring between neighbouring warps
constexpr int IROW_PD = 2;     // prefetch distance in row steps
constexpr int IROW_NS = IROW_PD + 1;  // staging slots per warp
constexpr int IROW_MAXW = 16, IROW_MAXT = IROW_MAXW * 32;  // warps per CTA: up to 128 registers per thread (22 warps x 80 registers spill)
// per warp and staging slot: the rhs rows and the sigt rows of the warp's ER elements in memory order (1 KB each per
// row of the slot) + one row of zone volumes per row of the slot when the population sum is fused in: IGeom::sb

struct IGeom {  // kernel parameter: lives in the constant bank, costs no registers
  int layout, Ds, Gs, ni, nj, nk;
  int NW, nkt;  // warps (= k-planes) per CTA, k tiles
  int lra;      // lanes of a row segment that really hold zones (ni / 4); the template's LR is the next power of two and the
                // lanes beyond lra idle at the end of their segment (their maps only reach other idle lanes in the scan)
  int E, ngroups;
  unsigned sb, vol_off;  // bytes of one staging slot and offset of the zone-volume rows in it
  int sig1;              // the ER rows of a warp always share their group (Ds % ER == 0): one sigt row (pair) per step
  unsigned sa, sg, Zs, ipd, ipg, jpd, jpg, kpd, kpg;  // element strides: psi/rhs (direction, group), zones, planes
};

__device__ __forceinline__ void ir_ldg256_cg(const double *p, double (&v)[4]) {
  asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p) : "memory");
}
__device__ __forceinline__ void ir_stg256(double *p, const double (&v)[4]) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}
__device__ __forceinline__ unsigned ir_smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double2 ir_lds128(unsigned addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void ir_sts128(unsigned addr, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(addr), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ void ir_st_release(unsigned addr, unsigned v) {
  asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// spin until the counter at `addr` has reached `target` (acquire at CTA scope; the counter only grows)
__device__ __forceinline__ void ir_wait_ge(unsigned addr, unsigned target) {
  unsigned v;
  do {
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  } while ((int)(v - target) < 0);
}
// TMA bulk copy global -> shared (contiguous bytes, 16-byte aligned, multiple of 16), completion on an mbarrier
__device__ __forceinline__ void ir_bulk_g2s(unsigned smem_dst, const void *gsrc, unsigned bytes, unsigned mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(mbar) : "memory");
}
// TMA tensor copy global -> shared of one box of a 2-D tensor map (rows of 16 doubles, 128-byte swizzle): coordinates are
// (element inside the 128-byte row = 0, row index)
__device__ __forceinline__ void ir_tensor_g2s(unsigned smem_dst, const void *tmap, unsigned row, unsigned mbar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_dst), "l"(tmap), "r"(0u), "r"(row), "r"(mbar) : "memory");
}
// 128-byte swizzle of the staging slots filled through a tensor map: the 16-byte chunk index (address bits 4-6) is XORed
// with the 128-byte row index (bits 7-9).  A lane's two 16-byte pieces sit 32 bytes apart from its neighbours', which
// without the swizzle is a 2-way bank conflict on every ld.shared.v2.f64 of the row.
__device__ __forceinline__ unsigned ir_swz(unsigned a) { return a ^ ((a >> 3) & 0x70u); }
__device__ __forceinline__ void ir_mb_expect_tx(unsigned addr, unsigned bytes) {
  asm volatile("{ .reg .b64 t; mbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1; }" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ double ir_ld_cg(const double *p) {
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void ir_mb_init(unsigned addr, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void ir_mb_arrive(unsigned addr) {  // release at CTA scope, no fence instruction
  asm volatile("{ .reg .b64 t; mbarrier.arrive.shared::cta.b64 t, [%0]; }" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void ir_mb_wait(unsigned addr, unsigned parity) {  // acquire at CTA scope, hardware-assisted wait
  asm volatile("{\n .reg .pred p;\n IRW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @!p bra IRW;\n}" ::"r"(addr), "r"(parity) : "memory");
}

struct IShared {  // shared-window addresses of the pieces of dynamic shared memory
  unsigned fkx;          // [IROW_RING][NW+1][2][32] double2: k-face exchange ring (entry 0 of a slot row is unused)
  unsigned kin;          // [IROW_NS][2][32] double2: warp 0's tile-boundary k faces
  unsigned stage;        // [NW][2 or IROW_NS] staging slots of IGeom::sb bytes
  unsigned full, empty;  // [NW+1][IROW_RING] mbarriers of the ring slots (32 arrivals each)
  unsigned sbar;         // [NW][IROW_NS] mbarriers of the staging slots (1 arrival + the bytes of the bulk copies)
  unsigned prod;         // [1] rows whose tile-boundary k faces the last warp has put into k_plane
  const double *cxt, *txc, *tyc, *tzc, *rdy, *rdz;  // [Ds] 2*xcos/dx[0], 2*xcos, 2*ycos, 2*zcos ; [nj] 1/dy ; [nk] 1/dz
};

// What a warp needs to know about one stream item = (element group, k tile): ER elements in one k-plane.
struct IItem {
  int d, g;                    // phase-space element of this lane's segment
  bool ev, kv;                 // segment holds a real element; this warp's k-plane exists in the tile
  int kz;                      // memory k index of the plane
  unsigned off0, soff0, ipx0;  // element offsets of the item's first row (its i = 0 zone): rhs/psi, sigt, i_plane entry
  unsigned kpx0;               // k_plane row of the first row (warp 0 only)
};

template <int LR>
__device__ __forceinline__ IItem irow_item(const IGeom &gm, int gi, int t, int w, int seg, unsigned i0, int jd, int kd) {
  constexpr int ER = 32 / LR;
  IItem it;
  const int e = gi * ER + seg;
  it.ev = e < gm.E;
  const int ee = it.ev ? e : gm.E - 1;  // idle segments shadow a valid element (loads stay in bounds, stores are masked)
  it.g = ee / gm.Ds;
  it.d = ee - it.g * gm.Ds;
  const int kl = t * gm.NW + w;
  it.kv = kl < gm.nk;
  const int klc = it.kv ? kl : gm.nk - 1;
  it.kz = (kd > 0) ? klc : gm.nk - 1 - klc;
  const int jz0 = (jd > 0) ? 0 : gm.nj - 1;
  const unsigned zoff = (unsigned)((it.kz * gm.nj + jz0) * gm.ni);
  (void)i0;
  it.off0 = (unsigned)it.d * gm.sa + (unsigned)it.g * gm.sg + zoff;
  it.soff0 = (unsigned)it.g * gm.Zs + zoff;
  it.ipx0 = (unsigned)it.d * gm.ipd + (unsigned)it.g * gm.ipg + (unsigned)(it.kz * gm.nj + jz0);
  it.kpx0 = (unsigned)it.d * gm.kpd + (unsigned)it.g * gm.kpg + (unsigned)(jz0 * gm.ni);
  return it;
}

// PAIR: rows j and j+1 of a k-plane are adjacent in memory, so with an even nj one bulk copy per array brings in
// TWO rows: half as many copy/mbarrier sequences per zone, two staging slots of a row pair each.
// POPM: 0 = no population sum, 1 = (w*psi)*volume(z) with the zone volumes staged per row, 2 = every zone has the volume
// `volc` (what Kripke's generator produces): the rows are summed first and scaled by w(d)*volc once per item
// SWZ (PAIR only): rhs and sigt row pairs come in through tensor maps with the 128-byte swizzle (tm_rhs / tm_sigt)
template <int LR, bool FWD, bool UNI, int POPM, bool PAIR, bool SWZ>
__device__ __forceinline__ double irow_run(const kb200_sweep_desc &ds, const IGeom &gm, const IShared &sh,
                                           const double *__restrict__ wq, const double *__restrict__ vol, const double volc,
                                           const void *tm_rhs, const void *tm_sigt) {
  constexpr bool POP = POPM == 1;  // the staged-volume machinery
  constexpr int R = IROW_RING, PD = IROW_PD, NS = PAIR ? 2 : IROW_NS;
  constexpr unsigned RB = PAIR ? 2048u : 1024u;  // bytes of the row block(s) of a warp's ER segments in one slot
  const unsigned SB = gm.sb;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int seg = lane / LR, ls = lane % LR;
  const int NW = gm.NW, ni = gm.ni, nj = gm.nj, nk = gm.nk, nkt = gm.nkt;
  const bool act = ls < gm.lra;  // this lane holds zones (ni need not be 4 * 2^k)
  const unsigned i0 = !act ? 0u : (FWD ? 4u * ls : (unsigned)(ni - 4 - 4 * ls));  // memory position of this lane's four zones
  const int jd = ds.jd, kd = ds.kd;
  const bool i_zero = ds.inflow_zero[0] != 0, j_zero = ds.inflow_zero[1] != 0, k_zero = ds.inflow_zero[2] != 0;
  const int jstep = (jd > 0) ? ni : -ni;
  const int gstride = (int)gridDim.x;

  const unsigned rowb = 8u * (unsigned)ni;                                  // bytes of one i-row
  const unsigned stage0 = sh.stage + (unsigned)(w * NS) * SB;               // + slot * SB : [rhs rows of the ER segments][sigt rows]
  const unsigned kin0 = sh.kin;                                             // + slot * RB (warp 0): k_plane rows of the ER segments
  const unsigned sbar0 = sh.sbar + 8u * (unsigned)(w * NS);                 // + 8 * slot
  const unsigned mine = (unsigned)seg * (PAIR ? 2u : 1u) * rowb + 8u * i0;   // this lane's four zones inside a staged row block
  // k-face ring: one row per slot, or (PAIR) the two rows of a pair per slot and hand-shake
  constexpr unsigned RINGB = PAIR ? 2048u : 1024u;
  const unsigned fk_in0 = sh.fkx + (unsigned)w * RINGB + 16u * lane;        // + ring slot * (NW+1)*RINGB (+ 1024 for the pair's second row)
  const unsigned fkx_slot = (unsigned)(NW + 1) * RINGB;
  const unsigned full_in = sh.full + 8u * (unsigned)(w * R), full_out = full_in + 8u * R;
  const unsigned empty_in = sh.empty + 8u * (unsigned)(w * R), empty_out = empty_in + 8u * R;

  double pop = 0.0;
  double fj[4] = {0, 0, 0, 0};
  int gi = (int)blockIdx.x, t = 0;
  const bool any = gi < gm.ngroups;
  IItem nx = irow_item<LR>(gm, any ? gi : 0, 0, w, seg, i0, jd, kd);
  bool nx_ok = any;   // the item `nx` exists
  unsigned q = 0;     // row steps done by this warp
  unsigned slot = 0;  // q % R
  unsigned ph = 0;    // (q / R) & 1: phase parity of the ring slot's current use
  unsigned sq = 0;    // staging slot of the current row (advances with the rows this warp actually computes)
  unsigned sph = 0;   // phase parity of that slot's current use
  double fi0n = 0.0;  // incoming i-face flux of the next row

  // Loads of row `r` (0-based in item `it`, r < nj) whose step is qn, into staging slot sqn: one TMA bulk copy per
  // element row for rhs, sigt and (warp 0) the tile-boundary k faces, completion counted on the slot's mbarrier.
  auto prefetch = [&](const IItem &it, int r, bool ktile0, unsigned qn, unsigned sqn) {
    const bool kload = (w == 0) && !(k_zero && ktile0);
    // rows of a later k tile were written by the tile's predecessor nj row steps earlier
    if (w == 0 && !ktile0) {
      ir_wait_ge(sh.prod, qn + 1u - (unsigned)nj);
      asm volatile("fence.proxy.async.global;" ::: "memory");  // acquired generic-proxy stores -> the bulk copies below
    }
    const unsigned bar = sbar0 + 8u * sqn;
    if (lane == 0) ir_mb_expect_tx(bar, (kload ? 3u : 2u) * rowb * (unsigned)(32 / LR) + (POP ? rowb : 0u));
    __syncwarp();  // also: every lane is done reading the slot's previous contents
    if (ls == 0) {
      const unsigned roff = (unsigned)(r * jstep);
      const unsigned dst = stage0 + sqn * SB + (unsigned)seg * rowb;
      ir_bulk_g2s(dst, ds.rhs + (it.off0 + roff), rowb, bar);
      ir_bulk_g2s(dst + RB, ds.sigt + (it.soff0 + roff), rowb, bar);
      if (kload) ir_bulk_g2s(kin0 + sqn * RB + (unsigned)seg * rowb, ds.k_plane + (it.kpx0 + roff), rowb, bar);
      if (POP && lane == 0)  // the zone volumes of the row: the same for every element
        ir_bulk_g2s(stage0 + sqn * SB + gm.vol_off, vol + (it.soff0 - (unsigned)it.g * gm.Zs + roff), rowb, bar);
    }
  };
  // PAIR: rows 2*pp and 2*pp+1 (sweep order) of item `it`, whose first step is qn, as one copy per array
  auto prefetch2 = [&](const IItem &it, int pp, bool ktile0, unsigned qn, unsigned sqn) {
    const bool kload = (w == 0) && !(k_zero && ktile0);
    if (w == 0 && !ktile0) {
      ir_wait_ge(sh.prod, qn + 2u - (unsigned)nj);
      asm volatile("fence.proxy.async.global;" ::: "memory");
    }
    const unsigned bar = sbar0 + 8u * sqn;
    if (lane == 0)
      ir_mb_expect_tx(bar, (kload ? 2u : 1u) * 2u * rowb * (unsigned)(32 / LR) + 2u * rowb * (gm.sig1 ? 1u : (unsigned)(32 / LR)) + (POP ? 2u * rowb : 0u));
    __syncwarp();  // also: every lane is done reading the slot's previous contents
    if (ls == 0) {
      const unsigned roff = (unsigned)(((jd > 0) ? 2 * pp : 2 * pp + 1) * jstep);  // the lower of the two rows in memory
      const unsigned dst = stage0 + sqn * SB + (unsigned)seg * 2u * rowb;
      if (SWZ) {
        ir_tensor_g2s(dst, tm_rhs, (it.off0 + roff) >> 4, bar);
        if (!gm.sig1 || lane == 0) ir_tensor_g2s(dst + RB, tm_sigt, (it.soff0 + roff) >> 4, bar);
      } else {
        ir_bulk_g2s(dst, ds.rhs + (it.off0 + roff), 2u * rowb, bar);
        if (!gm.sig1 || lane == 0) ir_bulk_g2s(dst + RB, ds.sigt + (it.soff0 + roff), 2u * rowb, bar);
      }
      if (kload) ir_bulk_g2s(kin0 + sqn * RB + (unsigned)seg * 2u * rowb, ds.k_plane + (it.kpx0 + roff), 2u * rowb, bar);
      if (POP && lane == 0) ir_bulk_g2s(stage0 + sqn * SB + gm.vol_off, vol + (it.soff0 - (unsigned)it.g * gm.Zs + roff), 2u * rowb, bar);
    }
  };
  auto next_sq = [](unsigned s, int by) { unsigned v = s + (unsigned)by; return v >= (unsigned)NS ? v - NS : v; };

  // prologue: rows 0..PD-1 (the first row pair) of the first item
  if (PAIR) {
    if (nx_ok && nx.kv) prefetch2(nx, 0, true, 0u, 0u);
  } else {
#pragma unroll
    for (int r = 0; r < PD; ++r)
      if (nx_ok && nx.kv && r < nj) prefetch(nx, r, true, (unsigned)r, (unsigned)r);
  }
  if (nx_ok && nx.kv && !i_zero) fi0n = ir_ld_cg(ds.i_plane + nx.ipx0);

  while (nx_ok) {
    IItem it = nx;
    it.ev = it.ev && act;  // idle lanes compute on shadow data and store nothing
    const bool ktile0 = (t == 0);
    // the item after this one
    int ngi = gi, nt = t + 1;
    if (nt == nkt) { nt = 0; ngi += gstride; }
    nx_ok = ngi < gm.ngroups;
    nx = irow_item<LR>(gm, nx_ok ? ngi : gi, nt, w, seg, i0, jd, kd);
    const bool nx_pre = nx_ok && nx.kv;

    if (!it.kv) {  // this warp's plane does not exist in the (short) last k tile: only keep the ring handshakes going
      for (int j = 0; j < nj; j += (PAIR ? 2 : 1)) {
        if (w > 0) { ir_mb_wait(full_in + 8u * slot, ph); ir_mb_arrive(empty_in + 8u * slot); }
        if (w < NW - 1) {
          if (q >= (unsigned)(PAIR ? 2 * R : R)) ir_mb_wait(empty_out + 8u * slot, ph ^ 1u);
          ir_mb_arrive(full_out + 8u * slot);
        }
        q += PAIR ? 2u : 1u;
        if (++slot == R) { slot = 0; ph ^= 1u; }
      }
      if (PAIR) {  // the next item (tile 0 of the next group) starts from scratch
        if (nx_pre) prefetch2(nx, 0, nt == 0, q, sq);
      } else {
#pragma unroll
        for (int r = 0; r < PD; ++r)
          if (nx_pre && r < nj) prefetch(nx, r, nt == 0, q + (unsigned)r, next_sq(sq, r));
      }
      if (nx_pre && !i_zero) fi0n = ir_ld_cg(ds.i_plane + nx.ipx0);
      gi = ngi; t = nt;
      continue;
    }

    const int kl = t * NW + w;
    const bool k_out_global = (kl == nk - 1) || (w == NW - 1);
    const double cx = sh.cxt[it.d], cz = sh.tzc[it.d] * sh.rdz[it.kz], tyc = sh.tyc[it.d];
    const double wd = POPM ? wq[it.d] : 0.0;  // quadrature weight of this segment's direction
    double psum = 0.0;                        // POPM == 2: sum of this item's psi
    unsigned off = it.off0, ipx = it.ipx0;
    int jz = (jd > 0) ? 0 : nj - 1;

    // one row step; PAR = position of the row in its pair (PAIR mode, compile time) or -1
    auto row_step = [&](auto par_tag, const int j) {
      constexpr int PAR = decltype(par_tag)::value;
      // ---- loads of the row(s) PD steps ahead ----
      if (PAIR) {
        if (PAR == 0) {
          const unsigned sqn = sq ^ 1u;
          if (j + 2 < nj) prefetch2(it, j / 2 + 1, ktile0, q + 2u, sqn);
          else if (nx_pre) prefetch2(nx, 0, nt == 0, q + 2u, sqn);
        }
      } else {
        const int r = j + PD;
        const unsigned sqn = next_sq(sq, PD);
        if (r < nj) prefetch(it, r, ktile0, q + (unsigned)PD, sqn);
        else if (nx_pre && r - nj < nj) prefetch(nx, r - nj, nt == 0, q + (unsigned)PD, sqn);
      }
      const double fi0 = fi0n;
      if (!i_zero) {  // incoming i-face flux of the next row: one step ahead, in a register
        if (j + 1 < nj) fi0n = ir_ld_cg(ds.i_plane + (ipx + (unsigned)jd));
        else if (nx_pre) fi0n = ir_ld_cg(ds.i_plane + nx.ipx0);
      }
      if (!PAIR || PAR == 0) ir_mb_wait(sbar0 + 8u * sq, sph);  // the bulk copies issued for this row (pair) have landed

      // ---- everything that does not need the upwind k and j faces ----
      // PAIR: the first row of the pair is the lower one in memory when sweeping up in j, the upper one otherwise
      const unsigned prow = PAIR ? (((jd > 0) ? (unsigned)PAR : (unsigned)(PAR ^ 1)) * rowb) : 0u;
      const unsigned st = stage0 + sq * SB + mine + prow;
      double A[4], rc[4], r4[4];
      const double cy = tyc * sh.rdy[jz];
      {
        const unsigned sgo = (PAIR && gm.sig1) ? st - (unsigned)seg * 2u * rowb : st;  // shared rows: segment 0's block
        const double2 e = ir_lds128(SWZ ? ir_swz(sgo + RB) : sgo + RB), f = ir_lds128(SWZ ? ir_swz(sgo + RB + 16) : sgo + RB + 16);
        const double s4[4] = {e.x, e.y, f.x, f.y};
        const double2 a = ir_lds128(SWZ ? ir_swz(st) : st), b = ir_lds128(SWZ ? ir_swz(st + 16) : st + 16);
        r4[0] = a.x; r4[1] = a.y; r4[2] = b.x; r4[3] = b.y;
        const double csum = __dadd_rn(__dadd_rn(cx, cy), cz);
#pragma unroll
        for (int u = 0; u < 4; ++u) {  // u = position in sweep order, m = memory slot
          const int m = FWD ? u : 3 - u;
          double cxu = cx, cs = csum;
          if (!UNI) {
            cxu = __ddiv_rn(sh.txc[it.d], ds.dx[i0 + m]);
            cs = __dadd_rn(__dadd_rn(cxu, cy), cz);
          }
          const double den = __dadd_rn(cs, s4[m]);
          double y;
          asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(den));
          const double e1 = fma(-den, y, 1.0);  // cubic step: y*(1 + e + e^2), e ~ 2^-23 -> error ~ 2^-69
          const double e2 = fma(e1, e1, e1);
          rc[u] = fma(y, e2, y);
          A[u] = cxu * rc[u];
        }
      }
      if (j == 0) {
        if (j_zero) { fj[0] = fj[1] = fj[2] = fj[3] = 0.0; }
        else ir_ldg256_cg(ds.j_plane + ((unsigned)it.d * gm.jpd + (unsigned)it.g * gm.jpg + (unsigned)(it.kz * ni) + i0), fj);
      }
#pragma unroll
      for (int m = 0; m < 4; ++m) r4[m] = fma(fj[m], cy, r4[m]);  // rhs + fj*cy

      // ---- upwind k faces: from warp w-1 through the ring, or (warp 0) from the staged k_plane row ----
      double fk[4];
      {
        if (w > 0) {
          if (!PAIR || PAR == 0) ir_mb_wait(full_in + 8u * slot, ph);
          const unsigned src = fk_in0 + slot * fkx_slot + ((PAIR && PAR == 1) ? 1024u : 0u);
          const double2 k0 = ir_lds128(src), k1 = ir_lds128(src + 512);
          fk[0] = k0.x; fk[1] = k0.y; fk[2] = k1.x; fk[3] = k1.y;
          if (!PAIR || PAR == 1) ir_mb_arrive(empty_in + 8u * slot);  // the slot may be refilled
        } else if (k_zero && ktile0) {
          fk[0] = fk[1] = fk[2] = fk[3] = 0.0;
        } else {
          const unsigned src = kin0 + sq * RB + mine + prow;
          const double2 k0 = ir_lds128(src), k1 = ir_lds128(src + 16);
          fk[0] = k0.x; fk[1] = k0.y; fk[2] = k1.x; fk[3] = k1.y;
        }
      }
      // B = (rhs + fj*cy + fk*cz) / den ; this lane's composite map  fi_out = al * fi_in + be
      double B[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int m = FWD ? u : 3 - u;
        B[u] = fma(fk[m], cz, r4[m]) * rc[u];
      }
      double al = fma(2.0, A[0], -1.0), be = B[0] + B[0];
#pragma unroll
      for (int u = 1; u < 4; ++u) {
        const double a = fma(2.0, A[u], -1.0);
        be = fma(a, be, B[u] + B[u]);
        al *= a;
      }
      // inclusive scan over the LR lanes of the row
#pragma unroll
      for (int dlt = 1; dlt < LR; dlt <<= 1) {
        const double alp = __shfl_up_sync(0xffffffffu, al, dlt, LR);
        const double bep = __shfl_up_sync(0xffffffffu, be, dlt, LR);
        if (ls >= dlt) {
          be = fma(al, bep, be);
          al *= alp;
        }
      }
      const double fo = fma(al, fi0, be);  // outgoing i face of this lane's last zone
      double fi = fi0;
      if (LR > 1) {
        const double fprev = __shfl_up_sync(0xffffffffu, fo, 1, LR);
        if (ls > 0) fi = fprev;
      }
      double p4[4], ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int m = FWD ? u : 3 - u;
        const double p = fma(fi, A[u], B[u]);
        p4[m] = p;
        fi = fma(2.0, p, -fi);
        fj[m] = fma(2.0, p, -fj[m]);
        ok[m] = fma(2.0, p, -fk[m]);
      }
      if (w < NW - 1) {  // (the last plane of a short tile publishes too: idle warps keep shaking hands)
        if ((!PAIR || PAR == 0) && q >= (unsigned)(PAIR ? 2 * R : R)) ir_mb_wait(empty_out + 8u * slot, ph ^ 1u);  // warp w+1 has emptied this ring slot
        const unsigned dst = fk_in0 + slot * fkx_slot + RINGB + ((PAIR && PAR == 1) ? 1024u : 0u);
        ir_sts128(dst, ok[0], ok[1]);
        ir_sts128(dst + 512, ok[2], ok[3]);
        if (!PAIR || PAR == 1) ir_mb_arrive(full_out + 8u * slot);
      }
      if (it.ev) {
        ir_stg256(ds.psi + (off + i0), p4);
        if (POP) {
          const unsigned vs = stage0 + sq * SB + gm.vol_off + prow + 8u * i0;
          const double2 va = ir_lds128(vs), vb = ir_lds128(vs + 16);
          const double v4[4] = {va.x, va.y, vb.x, vb.y};
#pragma unroll
          for (int u = 0; u < 4; ++u) pop = fma(wd * p4[u], v4[u], pop);
        }
        if (POPM == 2) psum += (p4[0] + p4[1]) + (p4[2] + p4[3]);
        if (ls == gm.lra - 1) {
          ds.i_plane[ipx] = fo;
          if (ds.out_plane[0]) ds.out_plane[0][ipx] = fo;
        }
        if (j == nj - 1) {
          const unsigned jpx = (unsigned)it.d * gm.jpd + (unsigned)it.g * gm.jpg + (unsigned)(it.kz * ni) + i0;
          ir_stg256(ds.j_plane + jpx, fj);
          if (ds.out_plane[1]) ir_stg256(ds.out_plane[1] + jpx, fj);
        }
        if (k_out_global) {
          const unsigned kpx = it.kpx0 + (unsigned)(j * jstep) + i0;
          ir_stg256(ds.k_plane + kpx, ok);
          if (kl == nk - 1 && ds.out_plane[2]) ir_stg256(ds.out_plane[2] + kpx, ok);
        }
      }
      if (w == NW - 1 && nkt > 1) {  // tile-boundary k faces are in k_plane: tell warp 0
        // warp 0 fetches them with bulk copies (async proxy): order this lane's generic-proxy stores before them
        asm volatile("fence.proxy.async.global;" ::: "memory");
        __syncwarp();
        if (lane == 0) ir_st_release(sh.prod, q + 1u);
      }
      ++q;
      if (PAIR) {
        if (PAR == 1) { sq ^= 1u; if (sq == 0u) sph ^= 1u; }
      } else {
        if (++sq == (unsigned)NS) { sq = 0; sph ^= 1u; }
      }
      if (!PAIR || PAR == 1) {
        if (++slot == R) { slot = 0; ph ^= 1u; }
      }
      off += (unsigned)jstep; ipx += (unsigned)jd; jz += jd;
    };
    if (PAIR) {
#pragma unroll 1
      for (int j = 0; j < nj; j += 2) {
        row_step(std::integral_constant<int, 0>{}, j);
        row_step(std::integral_constant<int, 1>{}, j + 1);
      }
    } else {
#pragma unroll 2  // measured: 1 -> 17.0 ms, 2 -> 16.4 ms, 4 -> 21.8 ms (instruction cache) with the fused population sum
      for (int j = 0; j < nj; ++j) row_step(std::integral_constant<int, -1>{}, j);
    }
    if (POPM == 2) pop = fma(wd * volc, psum, pop);
    gi = ngi; t = nt;
  }
  return pop;
}

template <int LR, int POPM, bool PAIR, bool SWZ>
__global__ void __launch_bounds__(IROW_MAXT, 1) sweep_irow_kernel(const kb200_sweep_desc *__restrict__ descs, const __grid_constant__ IGeom gm,
                                                             const double *const *__restrict__ pop_w,
                                                             const double *const *__restrict__ pop_vol,
                                                             const double *__restrict__ pop_vol_const,
                                                             double *__restrict__ pop_partial,
                                                             const CUtensorMap *__restrict__ tmaps) {
  constexpr bool POP = POPM != 0;
  extern __shared__ __align__(1024) unsigned char ism[];
  __shared__ kb200_sweep_desc ds;  // the descriptor is read all the time: keep it one LDS away (measured: faster than
  // passing the descriptors by value and reading them through the constant bank)
  if (threadIdx.x < sizeof(kb200_sweep_desc) / sizeof(int))
    reinterpret_cast<int *>(&ds)[threadIdx.x] = reinterpret_cast<const int *>(&descs[blockIdx.y])[threadIdx.x];
  __syncthreads();
  const int Ds = gm.Ds, nj = gm.nj, nk = gm.nk, NW = gm.NW;
  IShared sh;
  unsigned char *p = ism;
  if (SWZ) p += (1024u - (ir_smem_addr(ism) & 1023u)) & 1023u;  // the swizzle pattern repeats every 1 KB of shared memory
  sh.fkx = ir_smem_addr(p); p += (size_t)IROW_RING * (NW + 1) * (PAIR ? 2048 : 1024);
  sh.kin = ir_smem_addr(p); p += (size_t)4 * 1024;
  sh.stage = ir_smem_addr(p); p += (size_t)NW * (PAIR ? 2 : IROW_NS) * gm.sb;
  double *tab = reinterpret_cast<double *>(p);
  double *cxt = tab, *txc = cxt + Ds, *tyc = txc + Ds, *tzc = tyc + Ds, *rdy = tzc + Ds, *rdz = rdy + nj;
  double *red = rdz + nk;  // [32] block reduction of the population partials
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(red + 32);
  sh.full = ir_smem_addr(bars);
  sh.empty = sh.full + 8u * (unsigned)((NW + 1) * IROW_RING);
  sh.sbar = sh.empty + 8u * (unsigned)((NW + 1) * IROW_RING);
  unsigned *prod = reinterpret_cast<unsigned *>(bars + 2 * (NW + 1) * IROW_RING + NW * IROW_NS);
  sh.prod = ir_smem_addr(prod);
  sh.cxt = cxt; sh.txc = txc; sh.tyc = tyc; sh.tzc = tzc; sh.rdy = rdy; sh.rdz = rdz;
  if (threadIdx.x == 0) *prod = 0u;
  if ((int)threadIdx.x < 2 * (NW + 1) * IROW_RING) ir_mb_init(sh.full + 8u * threadIdx.x, 32u);
  if ((int)threadIdx.x < NW * IROW_NS) ir_mb_init(sh.sbar + 8u * threadIdx.x, 1u);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // the TMA unit sees initialised barriers

  // 2*cos tables and reciprocal zone widths (SweepSubdomain.cpp:88-93), indexed by the MEMORY zone index
  for (int d = threadIdx.x; d < Ds; d += blockDim.x) {
    const double t2 = 2.0 * ds.xcos[d];
    txc[d] = t2;
    cxt[d] = t2 / ds.dx[0];
    tyc[d] = 2.0 * ds.ycos[d];
    tzc[d] = 2.0 * ds.zcos[d];
  }
  for (int j = threadIdx.x; j < nj; j += blockDim.x) rdy[j] = 1.0 / ds.dy[j];
  for (int k = threadIdx.x; k < nk; k += blockDim.x) rdz[k] = 1.0 / ds.dz[k];
  int uni = 1;
  for (int i = threadIdx.x; i < gm.ni; i += blockDim.x) uni &= (ds.dx[i] == ds.dx[0]);
  const bool uniform_x = __syncthreads_and(uni) != 0;  // also orders the table writes and the mbarrier inits

  const double *wq = POP ? pop_w[blockIdx.y] : nullptr;
  const double *vol = POPM == 1 ? pop_vol[blockIdx.y] : nullptr;
  const double volc = POPM == 2 ? pop_vol_const[blockIdx.y] : 0.0;
  const void *tm_rhs = SWZ ? (const void *)(tmaps + 2 * blockIdx.y) : nullptr;
  const void *tm_sigt = SWZ ? (const void *)(tmaps + 2 * blockIdx.y + 1) : nullptr;
  double pop;
  if (uniform_x) {
    if (ds.id > 0) pop = irow_run<LR, true, true, POPM, PAIR, SWZ>(ds, gm, sh, wq, vol, volc, tm_rhs, tm_sigt);
    else pop = irow_run<LR, false, true, POPM, PAIR, SWZ>(ds, gm, sh, wq, vol, volc, tm_rhs, tm_sigt);
  } else {
    if (ds.id > 0) pop = irow_run<LR, true, false, POPM, PAIR, SWZ>(ds, gm, sh, wq, vol, volc, tm_rhs, tm_sigt);
    else pop = irow_run<LR, false, false, POPM, PAIR, SWZ>(ds, gm, sh, wq, vol, volc, tm_rhs, tm_sigt);
  }

  if (POP) {  // fixed-order block reduction: lanes, then warps
    pop = warp_sum(pop);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = pop;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int i = 0; i < NW; ++i) t += red[i];
      pop_partial[blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
  }
}

}  // namespace kb200

using namespace kb200;

template <int LR, bool PAIR, bool SWZ>
static int launch_irow(const kb200_sweep_desc *d_descs, int n, const IGeom &gm, int cps, size_t smem, const double *const *pw,
                       const double *const *pv, const double *pvc, double *pp, const CUtensorMap *tm, cudaStream_t st) {
  dim3 grid(cps, n, 1);
  if (pp && pvc) {
    auto k = sweep_irow_kernel<LR, 2, PAIR, SWZ>;
    KB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, gm.NW * 32, smem, st>>>(d_descs, gm, pw, pv, pvc, pp, tm);
  } else if (pp) {
    auto k = sweep_irow_kernel<LR, 1, PAIR, SWZ>;
    KB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, gm.NW * 32, smem, st>>>(d_descs, gm, pw, pv, nullptr, pp, tm);
  } else {
    auto k = sweep_irow_kernel<LR, 0, PAIR, SWZ>;
    KB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, gm.NW * 32, smem, st>>>(d_descs, gm, nullptr, nullptr, nullptr, nullptr, tm);
  }
  return post_launch("sweep_irow");
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    cudaGetLastError();
  }
  return fn;
}
// a chunk of `count` doubles seen as rows of 16 doubles (128 bytes); one box = `box_rows` rows, 128-byte swizzle
static bool irow_tensor_map(CUtensorMap *tm, const double *base, unsigned long long count, unsigned box_rows) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc || count % 16 != 0 || ((uintptr_t)base & 127) != 0) return false;
  const cuuint64_t dims[2] = {16, count / 16};
  const cuuint64_t strides[1] = {128};
  const cuuint32_t box[2] = {16, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Returns 0 if the batch was handled, -1 if this path does not apply (caller falls back), >0 on error.
// pop_* (optional, device pointer tables of n entries + a scratch of pop_capacity doubles): fused population partials.
// h_pop_vol_const / d_pop_vol_const (optional, host and device copies of n doubles): > 0 = every zone of that subdomain has
// this volume; the uniform path is taken when that holds for all n subdomains.
int kb200_sweep_irow_try(const kb200_sweep_desc *h, int n, const void *d_descs, const double *const *d_pop_w,
                         const double *const *d_pop_vol, const double *h_pop_vol_const, const double *d_pop_vol_const,
                         double *d_pop_partial, int pop_capacity, int *pop_count, cudaStream_t st) {
  const int layout = h[0].layout;
  if (pop_count) *pop_count = 0;
  if (layout != 0 && layout != 2) return -1;
  if (exact_mode()) return -1;
  const char *env = getenv("KB200_SWEEP_IROW");
  if (env && env[0] == '0') return -1;
  const int ni = h[0].ni, nj = h[0].nj, nk = h[0].nk;
  if (ni % 4 != 0) return -1;
  const int lra = ni / 4;  // lanes per row; the kernel is instantiated for the next power of two
  if (lra > 32) return -1;
  int LR = 1;
  while (LR < lra) LR <<= 1;
  for (int i = 0; i < n; ++i) {
    const void *ptrs[] = {h[i].rhs, h[i].psi, h[i].sigt, h[i].j_plane, h[i].k_plane, h[i].out_plane[1], h[i].out_plane[2]};
    for (const void *p : ptrs)
      if (((uintptr_t)p & 31) != 0) return -1;
    if (((uintptr_t)h[i].i_plane & 7) != 0) return -1;
  }
  if ((double)h[0].Ds * h[0].Gs * ni * nj * nk >= 2147483648.0) return -1;  // 32-bit element offsets
  IGeom gm;
  gm.layout = layout; gm.Ds = h[0].Ds; gm.Gs = h[0].Gs; gm.ni = ni; gm.nj = nj; gm.nk = nk;
  {
    const long long Zs = (long long)ni * nj * nk;
    const Strides3 fs = strides_dgz(layout, gm.Ds, gm.Gs, Zs);
    const StridesP ips = strides_plane(layout, gm.Ds, gm.Gs, nj, nk), jps = strides_plane(layout, gm.Ds, gm.Gs, ni, nk),
                   kps = strides_plane(layout, gm.Ds, gm.Gs, ni, nj);
    gm.sa = (unsigned)fs.a; gm.sg = (unsigned)fs.g; gm.Zs = (unsigned)Zs;
    gm.ipd = (unsigned)ips.d; gm.ipg = (unsigned)ips.g; gm.jpd = (unsigned)jps.d; gm.jpg = (unsigned)jps.g;
    gm.kpd = (unsigned)kps.d; gm.kpg = (unsigned)kps.g;
    gm.E = gm.Ds * gm.Gs;
  }
  gm.lra = lra;
  gm.nkt = (nk + IROW_MAXW - 1) / IROW_MAXW;
  gm.NW = (nk + gm.nkt - 1) / gm.nkt;
  // rows are fetched IROW_PD steps ahead, at most into the next item; warp 0 must not wait for tile-boundary
  // k faces of a row the last warp can only produce after warp 0 has moved on
  if (nj <= IROW_PD) return -1;
  // staging layout: row pairs (two slots) when nj is even and it fits, single rows (three slots) otherwise
  gm.sig1 = (gm.Ds % (32 / LR) == 0) ? 1 : 0;
  const unsigned rowb = 8u * (unsigned)ni;
  const char *pe = getenv("KB200_IROW_PAIR");
  bool pair = (nj % 2 == 0) && nj >= 4 && !(pe && pe[0] == '0');
  bool volu = d_pop_partial && h_pop_vol_const && d_pop_vol_const;
  for (int i = 0; volu && i < n; ++i) volu = h_pop_vol_const[i] > 0.0;
  const char *se = getenv("KB200_IROW_SWIZZLE");
  bool swz = false;
  size_t smem = 0;
  for (;;) {
    const unsigned rb = pair ? 2048u : 1024u;
    gm.vol_off = rb + ((pair && gm.sig1) ? 2u * rowb : rb);  // a shared sigt row pair takes 2*rowb instead of a full block
    gm.sb = gm.vol_off + ((d_pop_partial && !volu) ? (pair ? 2u : 1u) * rowb : 0u);
    // swizzled staging (tensor maps): 1 KB row pairs per segment, every slot a multiple of 1 KB from a 1 KB aligned base
    swz = pair && LR >= 16 && lra == LR && !(se && se[0] == '0') && encode_tiled() != nullptr;
    if (swz) gm.sb = (gm.sb + 1023u) & ~1023u;
    const size_t align_slack = swz ? 1024 : 0;
    smem = (size_t)IROW_RING * (gm.NW + 1) * rb + (size_t)4 * 1024 + (size_t)gm.NW * (pair ? 2 : IROW_NS) * gm.sb +
           ((size_t)4 * gm.Ds + nj + nk + 32 + 2 * (gm.NW + 1) * IROW_RING + gm.NW * IROW_NS + 2) * sizeof(double) + align_slack;
    if (smem <= 226 * 1024 || !pair) break;
    pair = false;
  }
  const int ER = 32 / LR;
  const int ngroups = (gm.Ds * gm.Gs + ER - 1) / ER;
  gm.ngroups = ngroups;
  int per_sm = (65536 / 128) / (gm.NW * 32);
  const int by_smem = (int)((227 * 1024) / (smem + 1024));
  if (per_sm > by_smem) per_sm = by_smem;
  if (per_sm < 1) per_sm = 1;
  // CTAs per subdomain: minimise waves x (element groups per CTA + the fixed cost of a CTA, about one and a half groups'
  // worth of prologue and pipeline fill).  With many small subdomains in a batch (BASELINE config 4: ~90 subdomains of
  // 16^3 zones per stage) "one CTA per subdomain" would leave a third of the SMs idle.
  int cps = 1;
  {
    const long long slots = (long long)sm_count() * per_sm;
    double best = -1.0;
    for (int c = 1; c <= ngroups; ++c) {
      const long long waves = ((long long)c * n + slots - 1) / slots;
      const double cost = (double)waves * ((double)((ngroups + c - 1) / c) + 1.5);
      if (best < 0.0 || cost < best - 1e-9) { best = cost; cps = c; }
    }
  }
  double *pp = d_pop_partial;
  if (pp && (long long)cps * n > pop_capacity) pp = nullptr;
  if (pop_count) *pop_count = pp ? cps * n : 0;
  const kb200_sweep_desc *dd = (const kb200_sweep_desc *)d_descs;
  const double *pvc = (pp && volu) ? d_pop_vol_const : nullptr;
  // rows of at least 512 bytes, fetched as pairs: stage rhs and sigt through tensor maps with the 128-byte swizzle
  const CUtensorMap *d_tm = nullptr;
  if (swz) {
    std::vector<CUtensorMap> tm(2 * (size_t)n);
    const unsigned long long Zs = (unsigned long long)ni * nj * nk;
    bool ok = true;
    for (int i = 0; i < n && ok; ++i)
      ok = irow_tensor_map(&tm[2 * i], h[i].rhs, (unsigned long long)gm.Ds * gm.Gs * Zs, 2u * rowb / 128u) &&
           irow_tensor_map(&tm[2 * i + 1], h[i].sigt, (unsigned long long)gm.Gs * Zs, 2u * rowb / 128u);
    const void *d = nullptr;
    if (ok && device_descs(tm.data(), tm.size() * sizeof(CUtensorMap), &d, st) == 0) d_tm = (const CUtensorMap *)d;
    KB_REQUIRE(d_tm, "sweep_irow: cannot build the tensor maps of the staged rows");
  }
#define IROW_LAUNCH(LR_) return pair ? launch_irow<LR_, true, false>(dd, n, gm, cps, smem, d_pop_w, d_pop_vol, pvc, pp, nullptr, st) \
                                    : launch_irow<LR_, false, false>(dd, n, gm, cps, smem, d_pop_w, d_pop_vol, pvc, pp, nullptr, st)
#define IROW_LAUNCH_SWZ(LR_) if (swz) return launch_irow<LR_, true, true>(dd, n, gm, cps, smem, d_pop_w, d_pop_vol, pvc, pp, d_tm, st)
  switch (LR) {
    case 1: IROW_LAUNCH(1);
    case 2: IROW_LAUNCH(2);
    case 4: IROW_LAUNCH(4);
    case 8: IROW_LAUNCH(8);
    case 16: IROW_LAUNCH_SWZ(16); IROW_LAUNCH(16);
    default: IROW_LAUNCH_SWZ(32); IROW_LAUNCH(32);
  }
#undef IROW_LAUNCH
#undef IROW_LAUNCH_SWZ
}
