// Scattering, zone-fastest storage orders (DGZ, GDZ), default arithmetic: ONE read of the source moments per call.
// Reference: src/Kripke/Kernel/Scattering.cpp:73-99 (+ Kernel/Source.cpp:59-75 folded into the epilogue).
//
//   phi_out(nm,g,z) (+)= sum_m f_m(z) * sum_k S_{m,n(nm)}[g,k] * phi(nm,k,z)       k = all source groups
//
// kb200_scatter_mma.cu gives every destination group set its own CTAs, so each of them streams the whole source
// slab again (ncu, config 2: 1.44x the algorithmic DRAM bytes; config 3 needs four 32-group output chunks -> 4 reads),
// and it works in slabs of 16 source groups: 380 instructions per warp and slab around 32 DMMAs.  This kernel
// * groups the descriptors that read the same source chunks (all destination group sets of one zone set) and lets one
//   CTA produce ALL their outputs from one staged tile [K source groups][NT zones];
// * stages the tiles through bulk-TMA row copies: sixteen consumer warps in NG groups (four groups of 2 x 2 warps by default),
//   every group with its own producer warp and its own 3-stage ring ([KC source groups][NT zones] plus the tile's three
//   material-fraction rows per stage); the groups take the CTA's tiles in turn, so one group's epilogue overlaps the
//   other groups' DMMAs.  A consumer warp owns an (output block, 16-zone column block) pair and runs mma.sync.m8n8k4.f64
//   with both operands fetched by 128-bit shared loads (zones 2j/2j+1 of a column block are the two B fragments of lane
//   group j, so one load feeds two DMMAs and an accumulator lane ends up with four consecutive zones = one 32-byte store);
// * keeps the three material matrices of the current Legendre order resident (fragment-major, prebuilt once per call
//   in global scratch and copied when the order changes; tiles are walked moment-major so that is L+1 times per CTA);
// * where the matrices of all outputs do not fit (config 3: 3 x 128 x 128 doubles), CS sibling CTAs take 32-output
//   chunks of the same tile sequence and keep within a tile of each other (advisory progress counters in global
//   memory, never a correctness dependency), so the siblings' re-reads of a tile hit L2.
// Anything this kernel does not cover (odd shapes, unaligned chunks, irregular descriptor lists) returns -1 and the
// per-descriptor kernel of kb200_scatter_mma.cu runs instead.  The same kernel serves LPlusTimes with a long reduction as a
// plain product (kb200_gemm_slab_try: one matrix, unit fractions).
#include <type_traits>
#include <vector>
#include "kb200_common.cuh"

namespace kb200 {

struct SlabGeom {
  int M, Zs, K, nkc4, Otot, CS, ngroups;
  int nst, ntn, nslots, accumulate;
  int nmat, unit;       // matrices per Legendre order (3 materials); unit = 1: a single matrix, every fraction 1 (kb200_gemm_slab_try)
  int exp;              // timing experiments (KB200_SLAB_EXP): 1 no tensor-core work, 2 no copies, 4 no stores -- wrong results
  long long in_b;       // moment stride of phi / phi_out (elements)
  int ntiles;           // M * ngroups * ntn
  unsigned ws_doubles;  // fragment-major matrices of one (order, sibling): 3 * nkc4 * 32 * QP * WO
  double source;
};
// device-side tables of one call (uploaded through the content-addressed descriptor cache)
struct SlabTables {
  const double *const *inrow;  // [ngroups][4*nkc4] row bases of the source groups (nullptr beyond K)
  double *const *orow;         // [ngroups][Otot]   row bases of the destination groups
  const double *frac;          // [ngroups][3][Zs]  volume fraction of each material
  const double *wg;            // [L1][CS][ws_doubles]
  const int *m2l;              // [M] moment -> Legendre order
  unsigned *progress;          // [nslots][CS] tiles issued by each sibling's producer
};

__device__ __forceinline__ unsigned sl_smem(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sl_mb_init(unsigned a, unsigned n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(n) : "memory"); }
__device__ __forceinline__ void sl_mb_arrive(unsigned a) { asm volatile("{ .reg .b64 t; mbarrier.arrive.shared::cta.b64 t, [%0]; }" ::"r"(a) : "memory"); }
__device__ __forceinline__ void sl_mb_expect_tx(unsigned a, unsigned bytes) {
  asm volatile("{ .reg .b64 t; mbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1; }" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sl_mb_wait(unsigned a, unsigned parity) {
  asm volatile("{\n .reg .pred p;\n SLW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @!p bra SLW;\n}" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void sl_bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ double2 sl_lds128(unsigned a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void sl_dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int THREADS>
__device__ __forceinline__ void sl_bar_consumers() { asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory"); }

// volume fraction of each of the three materials per zone of a descriptor group: frac[(group*3 + m)*Zs + z]
__global__ void slab_fractions_kernel(const kb200_scattering_desc *__restrict__ descs, const int *__restrict__ first_desc, int Zs,
                                      double *__restrict__ frac) {
  const kb200_scattering_desc &ds = descs[first_desc[blockIdx.y]];
  const int z = blockIdx.x * blockDim.x + threadIdx.x;
  if (z >= Zs) return;
  double f[3] = {0.0, 0.0, 0.0};
  const int m0 = ds.zone_to_mixelem[z], nmix = ds.zone_to_num_mixelem[z];
  for (int k = 0; k < nmix; ++k) {
    const int mat = ds.mixelem_to_material[m0 + k];
    const double fr = ds.mixelem_to_fraction[m0 + k];
    if (mat == 0) f[0] += fr; else if (mat == 1) f[1] += fr; else f[2] += fr;
  }
  double *o = frac + (size_t)blockIdx.y * 3 * Zs + z;
  o[0] = f[0]; o[(size_t)Zs] = f[1]; o[2 * (size_t)Zs] = f[2];
}

// fragment-major material matrices: wg[((n*CS + c)*WO + wo)*3 + m][kc][a/2][lane][a%2] = S_{m,n}[oglob[o]][kglob[k]] with
// o = (c*WO + wo)*8*QP + 8*a + lane/4, k = 4*kc + lane%4 (the A fragment of mma.m8n8k4: row = lane/4, column = lane%4)
__global__ void slab_matrices_kernel(const double *__restrict__ sigs, Strides4 ss, const int *__restrict__ oglob, const int *__restrict__ kglob,
                                     int L1, int CS, int WO, int QP, int nkc4, int K, double *__restrict__ wg) {
  const long long total = (long long)L1 * CS * WO * 3 * nkc4 * 32 * QP;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    long long r = idx;
    const int a0 = (int)(r % 2); r /= 2;
    const int lane = (int)(r % 32); r /= 32;
    const int a = 2 * (int)(r % (QP / 2)) + a0; r /= QP / 2;
    const int kc = (int)(r % nkc4); r /= nkc4;
    const int m = (int)(r % 3); r /= 3;
    const int wo = (int)(r % WO); r /= WO;
    const int c = (int)(r % CS); r /= CS;
    const int n = (int)r;
    const int o = (c * WO + wo) * 8 * QP + 8 * a + (lane >> 2), k = 4 * kc + (lane & 3);
    double v = 0.0;
    if (k < K) v = sigs[(long long)m * ss.mat + (long long)n * ss.n + (long long)oglob[o] * ss.g + (long long)kglob[k] * ss.gp];
    wg[idx] = v;
  }
}

// A stage holds KC = 4096/NT source-group rows of NT zones followed by the three material-fraction rows of the tile.  Rows
// sit SL_PITCH_PAD bytes further apart than their length: a quarter warp of a B-fragment ld.shared.v2.f64 touches four
// consecutive rows at two adjacent 16-byte columns, and with a pitch that is a multiple of 128 bytes those four rows would
// share their banks (4-way conflict; ncu of the first version: 302 M of 486 M shared wavefronts were conflicts).
// pitch = 32 (mod 128) spreads them over the 32 banks.
constexpr int SL_STAGES = 3, SL_PITCH_PAD = 32;
__host__ __device__ constexpr int sl_stage_rows(int NT, int NG) { return 4096 / NG / NT; }
__host__ __device__ constexpr int sl_stage_bytes(int NT, int NG) { return (sl_stage_rows(NT, NG) + 3) * (8 * NT + SL_PITCH_PAD); }

// NG groups of WO x WN consumer warps (WO output blocks of 8*QP outputs x WN column blocks of 16 zones), each with
// its own producer warp and its own ring of SL_STAGES stages.  A group works on a whole tile; the groups take the tiles of
// the CTA in turn and share nothing but the material matrices, so while one is busy with its epilogue and the bookkeeping
// of its next tile the other keeps the fp64 pipe fed (with a single set of warps all of them reach their epilogues
// together -- the stage ring keeps them in step -- and the pipe idles for a third of the time).  The rings are separate
// because an mbarrier wait names a phase only by its parity: a group waiting for a stage that another group has yet to
// consume twice over would see the wait succeed on the older phase.
template <int QP, int WO, int WN, int NG>
__global__ void __launch_bounds__(32 * (WO * WN + 1) * NG, 1) scatter_slab_kernel(SlabGeom gm, SlabTables tb) {
  constexpr int GW = WO * WN, NW = GW * NG, NT = 16 * WN, KC = sl_stage_rows(NT, NG);  // zones per tile, source groups per stage
  constexpr int SL_THREADS = 32 * (GW + 1) * NG;
  constexpr unsigned PITCH = 8 * NT + SL_PITCH_PAD, SL_STAGE_BYTES = sl_stage_bytes(NT, NG), FRAC_OFF = KC * PITCH;
  extern __shared__ __align__(128) unsigned char slm[];
  const unsigned ws_b = gm.ws_doubles * 8u;
  double *Ws = reinterpret_cast<double *>(slm);
  const unsigned stage0 = sl_smem(slm) + ws_b;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grpw = warp < NW ? warp / GW : warp - NW;  // consumer group of this warp / the group this producer warp feeds
  const unsigned ring0 = stage0 + (unsigned)grpw * (SL_STAGES * SL_STAGE_BYTES);
  const unsigned full0 = stage0 + NG * SL_STAGES * SL_STAGE_BYTES + (unsigned)grpw * (16u * SL_STAGES), empty0 = full0 + 8u * SL_STAGES;
  const int slot = blockIdx.x / gm.CS, sib = blockIdx.x - slot * gm.CS;
  const int per_b = gm.ngroups * gm.ntn;

  if (lane < SL_STAGES && warp >= NW) {  // each producer sets up the barriers of its group
    sl_mb_init(full0 + 8u * lane, 1u);
    sl_mb_init(empty0 + 8u * lane, (unsigned)GW);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  // rows of a stage beyond K and columns beyond the zone count are never written by a copy: they must hold finite values
  for (unsigned i = threadIdx.x; i < NG * SL_STAGES * SL_STAGE_BYTES / 16; i += SL_THREADS)
    reinterpret_cast<double2 *>(slm + ws_b)[i] = make_double2(0.0, 0.0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();

  if (warp >= NW) {
    // ---- producer of group grpw: one bulk copy per source-group row of a stage, plus the tile's three fraction rows ----
    unsigned it = 0;
    bool lockstep = gm.CS > 1;
    unsigned done = 0;
    for (int t = slot + grpw * gm.nslots; t < gm.ntiles; t += NG * gm.nslots, ++done) {
      const int b = t / per_b, r = t - b * per_b, grp = r / gm.ntn, tn = r - grp * gm.ntn;
      const int z0 = tn * NT;
      const unsigned rb = 8u * (unsigned)min(NT, gm.Zs - z0);
      if (lockstep && done > 0) {  // advisory: stay within a tile of the siblings so that their reads of this tile meet in L2
        bool ok = true;
        if (lane < gm.CS && lane != sib) {
          const volatile unsigned *p = tb.progress + ((size_t)slot * gm.CS + lane) * NG + grpw;
          int polls = 0;
          while (*p + 1u < done && ++polls < 512) {}
          ok = polls < 512;
        }
        if (!__all_sync(0xffffffffu, ok)) lockstep = false;  // a sibling is not keeping up (not resident?): stop waiting for it
      }
      const double *const *rows = tb.inrow + (size_t)grp * 4 * gm.nkc4;
      const double *frow = gm.unit ? nullptr : tb.frac + (size_t)grp * 3 * gm.Zs + z0;
      const int nfr = gm.unit ? 0 : 3;  // fraction rows riding in the stage
      const long long off = (long long)b * gm.in_b + z0;
      for (int st = 0; st < gm.nst; ++st, ++it) {
        const unsigned s = it % SL_STAGES, ph = (it / SL_STAGES) & 1u;
        sl_mb_wait(empty0 + 8u * s, ph ^ 1u);
        const int nrows = min(KC, gm.K - st * KC);
        if (gm.exp & 2) { if (lane == 0) sl_mb_arrive(full0 + 8u * s); continue; }
        if (lane == 0) sl_mb_expect_tx(full0 + 8u * s, (unsigned)(nrows + nfr) * rb);
        __syncwarp();
        const unsigned sb = ring0 + s * SL_STAGE_BYTES;
        for (int rr = lane; rr < nrows + nfr; rr += 32) {
          if (rr < nrows) sl_bulk_g2s(sb + (unsigned)rr * PITCH, rows[st * KC + rr] + off, rb, full0 + 8u * s);
          else sl_bulk_g2s(sb + FRAC_OFF + (unsigned)(rr - nrows) * PITCH, frow + (size_t)(rr - nrows) * gm.Zs, rb, full0 + 8u * s);
        }
      }
      if (gm.CS > 1 && lane == 0) {
        __threadfence();
        *(volatile unsigned *)(tb.progress + ((size_t)slot * gm.CS + sib) * NG + grpw) = done + 1u;
      }
    }
    return;
  }

  // ---- consumers ----
  const int wg = warp - grpw * GW;
  const int wo = wg / WN, wn = wg - wo * WN;
  const int col0 = wn * 16, jq = lane >> 2, kq = lane & 3;
  const unsigned a_lane = sl_smem(Ws) + (unsigned)(wo * gm.nmat * gm.nkc4 * 32) * (8u * QP) + 16u * lane;
  const unsigned b_lane = (unsigned)kq * PITCH + (unsigned)(col0 + 2 * jq) * 8u;
  const int ncta = (gm.ntiles - slot + gm.nslots - 1) / gm.nslots;  // tiles of this CTA: t = slot + i * nslots
  bool first = true;
  double *op[QP];  // output rows of this lane in the zone set op_grp
  int op_grp = -1;
  for (int b_lo = 0; b_lo < gm.M;) {
    // the moments [b_lo, b_hi) share their Legendre order, i.e. the material matrices
    const int n_leg = gm.unit ? 0 : __ldg(tb.m2l + b_lo);
    int b_hi = b_lo + 1;
    if (gm.unit) b_hi = gm.M;
    else while (b_hi < gm.M && __ldg(tb.m2l + b_hi) == n_leg) ++b_hi;
    const int i_lo = (max(b_lo * per_b - slot, 0) + gm.nslots - 1) / gm.nslots;
    const int i_hi = min(ncta, (max(b_hi * per_b - slot, 0) + gm.nslots - 1) / gm.nslots);
    b_lo = b_hi;
    if (i_lo >= i_hi) continue;  // uniform over the consumers of the CTA
    if (!first) sl_bar_consumers<32 * NW>();  // nobody still reads the previous matrices
    first = false;
    {
      const double2 *src = reinterpret_cast<const double2 *>(tb.wg + ((size_t)n_leg * gm.CS + sib) * gm.ws_doubles);
      for (unsigned i = threadIdx.x; i < gm.ws_doubles / 2; i += 32 * NW) reinterpret_cast<double2 *>(Ws)[i] = __ldg(src + i);
    }
    sl_bar_consumers<32 * NW>();

    for (int i = i_lo + ((grpw - i_lo) % NG + NG) % NG; i < i_hi; i += NG) {
      const int t = slot + i * gm.nslots;
      const int b = t / per_b, r = t - b * per_b, grp = r / gm.ntn, tn = r - grp * gm.ntn;
      double acc[QP][2][2];
#pragma unroll
      for (int a = 0; a < QP; ++a) acc[a][0][0] = acc[a][0][1] = acc[a][1][0] = acc[a][1][1] = 0.0;
      double sx[4] = {0.0, 0.0, 0.0, 0.0};
      unsigned present = 0;
      if (grp != op_grp) {
        op_grp = grp;
        double *const *orow = tb.orow + (size_t)grp * gm.Otot + (sib * WO + wo) * 8 * QP + jq;
#pragma unroll
        for (int a = 0; a < QP; ++a) op[a] = orow[8 * a];
      }

      for (int st = 0; st < gm.nst; ++st) {
        const unsigned it = (unsigned)((i / NG) * gm.nst + st), s = it % SL_STAGES, ph = (it / SL_STAGES) & 1u;
        const unsigned sb = ring0 + s * SL_STAGE_BYTES;
        sl_mb_wait(full0 + 8u * s, ph);
        const int kc_lo = st * (KC / 4), nkc = min(KC / 4, gm.nkc4 - kc_lo);
        const unsigned fa0 = sb + FRAC_OFF + (unsigned)(col0 + 2 * jq) * 8u;
        if (gm.unit) present = 1u;
        else if (st == 0) {  // which materials occur in this warp's 16 zones (every stage of the tile carries the same fraction rows)
          const double2 f0 = sl_lds128(fa0), f1 = sl_lds128(fa0 + PITCH), f2 = sl_lds128(fa0 + 2 * PITCH);
          const unsigned mine = ((f0.x != 0.0 || f0.y != 0.0) ? 1u : 0u) | ((f1.x != 0.0 || f1.y != 0.0) ? 2u : 0u) |
                                ((f2.x != 0.0 || f2.y != 0.0) ? 4u : 0u);
          present = __reduce_or_sync(0xffffffffu, mine);
          if (gm.exp & 1) present = 0;
        }
        // one pass per material present (nearly always one); absent zones carry a zero fraction
#pragma unroll 1
        for (int m = 0; m < 3; ++m) {
          if (!((present >> m) & 1u)) continue;
          const double2 fm = gm.unit ? make_double2(1.0, 1.0) : sl_lds128(fa0 + (unsigned)m * PITCH);
          unsigned ap = a_lane + (unsigned)((m * gm.nkc4 + kc_lo) * 32) * (8u * QP);
          unsigned bp = sb + b_lane;
          // all 16 zones pure in this material (the usual case): the B fragments need no scaling
          const bool pure = gm.unit || __all_sync(0xffffffffu, fm.x == 1.0 && fm.y == 1.0);
          auto pass = [&](auto scaled) {
#pragma unroll 4
            for (int kc = 0; kc < nkc; ++kc, ap += 32u * 8u * QP, bp += 4u * PITCH) {
              const double2 bv = sl_lds128(bp);
              double af[QP];
#pragma unroll
              for (int a = 0; a < QP; a += 2) {
                const double2 av = sl_lds128(ap + 256u * a);  // [a / 2][lane][a % 2]: 16 bytes per lane, lanes contiguous
                af[a] = av.x; af[a + 1] = av.y;
              }
              const double b0 = decltype(scaled)::value ? bv.x * fm.x : bv.x, b1 = decltype(scaled)::value ? bv.y * fm.y : bv.y;
#pragma unroll
              for (int a = 0; a < QP; ++a) {
                sl_dmma(acc[a][0][0], acc[a][0][1], af[a], b0);
                sl_dmma(acc[a][1][0], acc[a][1][1], af[a], b1);
              }
            }
          };
          if (pure) pass(std::false_type{});
          else pass(std::true_type{});
        }
        if (st == gm.nst - 1 && gm.source != 0.0 && b == 0) {
          // Kernel/Source.cpp:59-75: phi_out(0,g,z) += strength * fraction of material 0, for this lane's four output zones
          const unsigned fa = sb + FRAC_OFF + (unsigned)(col0 + 4 * kq) * 8u;
          const double2 u = sl_lds128(fa), v = sl_lds128(fa + 16u);
          sx[0] = gm.source * u.x; sx[1] = gm.source * u.y; sx[2] = gm.source * v.x; sx[3] = gm.source * v.y;
        }
        // hand the stage back: every lane's shared loads above are ordered before lane 0's arrive (release) by the warp
        // barrier, and the producer's refill is ordered after its wait (acquire) on the same mbarrier.  compute-sanitizer's
        // racecheck still reports a "potential WAR hazard" between a lane != 0 and the refill, because that lane does not
        // arrive itself (profiles/r02bl_racecheck.log); one arrival per lane would cost 32 serialised shared atomics per warp
        // and stage.  memcheck is clean (profiles/r02bl_memcheck.log).
        __syncwarp();
        if (lane == 0) sl_mb_arrive(empty0 + 8u * s);
      }

      // epilogue: lane holds outputs 8a + jq at the four consecutive zones col0 + 4*kq + {0,1,2,3}
      const int z = tn * NT + col0 + 4 * kq;
      if (z < gm.Zs && !(gm.exp & 4)) {
        const long long off = (long long)b * gm.in_b + z;
#pragma unroll
        for (int a = 0; a < QP; ++a) {
          double *p = op[a] + off;
          double v0 = acc[a][0][0] + sx[0], v1 = acc[a][1][0] + sx[1], v2 = acc[a][0][1] + sx[2], v3 = acc[a][1][1] + sx[3];
          if (gm.accumulate) {
            double o0, o1, o2, o3;
            asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(o0), "=d"(o1), "=d"(o2), "=d"(o3) : "l"(p) : "memory");
            v0 += o0; v1 += o1; v2 += o2; v3 += o3;
          }
          asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v0), "d"(v1), "d"(v2), "d"(v3) : "memory");
        }
      }
    }
  }
}

struct GemmDescK {  // = MomentsDescK of kb200_moments*.cu: pointer tables inside the device copy of the ABI descriptor
  const double *const *w;
  const double *const *in;
  double *const *out;
};
// fragment-major weights of the plain product out[o][n] = sum_k W[o][k] in[k][n], W[o][k] = w[o / Ds][(o % Ds) * M + k]
// (LPlusTimes: o = direction, k = moment): wg[(c*WO + wo)][kc][a/2][lane][a%2], one matrix, no Legendre orders
__global__ void gemm_matrices_kernel(const GemmDescK *__restrict__ descs, int Ds, int M, int O, int K, int CS, int WO, int QP, int nkc4,
                                     double *__restrict__ wg) {
  const GemmDescK d0 = descs[0];
  const long long total = (long long)CS * WO * nkc4 * 32 * QP;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    long long r = idx;
    const int a0 = (int)(r % 2); r /= 2;
    const int lane = (int)(r % 32); r /= 32;
    const int a = 2 * (int)(r % (QP / 2)) + a0; r /= QP / 2;
    const int kc = (int)(r % nkc4); r /= nkc4;
    const int wo = (int)(r % WO); r /= WO;
    const int c = (int)r;
    const int o = (c * WO + wo) * 8 * QP + 8 * a + (lane >> 2), k = 4 * kc + (lane & 3);
    double v = 0.0;
    if (o < O && k < K) v = d0.w[o / Ds][(size_t)(o % Ds) * M + k];
    wg[idx] = v;
  }
}

static void *g_slab_scratch = nullptr;
static size_t g_slab_scratch_bytes = 0;

}  // namespace kb200

using namespace kb200;

// Returns 0 if handled, -1 if this path does not apply (the caller goes on to the per-descriptor kernel), >0 on error.
// What the slab kernel would do with a list of descriptors: pure host logic (no CUDA call), also exported for the CPU tests.
struct SlabPlan {
  int ngroups = 0, nd = 0, Otot = 0, K = 0, Octa = 0, CS = 0;
  std::vector<int> first;                  // first descriptor of each group
  std::vector<std::vector<int>> members;   // descriptors of each group, in call order
};
static bool slab_plan(const kb200_scattering_desc *h, int n, int sms, SlabPlan &pl) {
  const int layout = h[0].layout;
  if (layout != 0 && layout != 2) return false;
  const int Zs = h[0].Zs, Gs = h[0].Gs, nsrc = h[0].nsrc;
  if (Zs % 4 != 0) return false;
  // descriptor groups: same source chunks and zone tables (= the destination group sets of one zone set)
  std::vector<int> &first = pl.first;
  std::vector<std::vector<int>> &members = pl.members;
  for (int i = 0; i < n; ++i) {
    if (h[i].nsrc != nsrc || h[i].accumulate != h[0].accumulate || h[i].G != h[0].G || h[i].L1 != h[0].L1 || h[i].sigs != h[0].sigs ||
        h[i].moment_to_legendre != h[0].moment_to_legendre)
      return false;
    if (((uintptr_t)h[i].phi_out & 31) != 0) return false;
    for (int s = 0; s < nsrc; ++s)
      if (((uintptr_t)h[i].phi_src[s] & 15) != 0) return false;
    int g = -1;
    for (size_t q = 0; q < first.size() && g < 0; ++q) {
      const kb200_scattering_desc &f = h[first[q]];
      bool same = f.zone_to_mixelem == h[i].zone_to_mixelem && f.zone_to_num_mixelem == h[i].zone_to_num_mixelem &&
                  f.mixelem_to_material == h[i].mixelem_to_material && f.mixelem_to_fraction == h[i].mixelem_to_fraction;
      for (int s = 0; s < nsrc && same; ++s) same = f.phi_src[s] == h[i].phi_src[s] && f.glower_src[s] == h[i].glower_src[s];
      if (same) g = (int)q;
    }
    if (g < 0) { g = (int)first.size(); first.push_back(i); members.emplace_back(); }
    members[g].push_back(i);
  }
  const int ngroups = (int)first.size(), nd = (int)members[0].size();
  for (int g = 0; g < ngroups; ++g) {
    if ((int)members[g].size() != nd) return false;
    for (int q = 0; q < nd; ++q) {
      if (h[members[g][q]].glower_dst != h[members[0][q]].glower_dst) return false;
      for (int s = 0; s < nsrc; ++s)
        if (h[members[g][q]].glower_src[s] != h[members[0][0]].glower_src[s]) return false;
    }
  }
  const int Otot = nd * Gs, K = nsrc * Gs, Kp = 4 * ((K + 3) / 4);
  if (Otot % 32 != 0) return false;
  // outputs per CTA: everything in one CTA if the three matrices fit next to the stages, else 32-output sibling chunks
  const size_t smem_cap = 227 * 1024 - 4 * SL_STAGES * (size_t)sl_stage_bytes(64, 4) - 256;  // the largest rings of the variants
  int Octa = 0;
  if (Otot % 64 == 0 && (size_t)3 * Kp * 64 * 8 <= smem_cap) Octa = 64;
  else if ((size_t)3 * Kp * 32 * 8 <= smem_cap) Octa = 32;
  else return false;
  const int CS = Otot / Octa;
  if (CS > 32 || CS > sms) return false;
  pl.ngroups = ngroups; pl.nd = nd; pl.Otot = Otot; pl.K = K; pl.Octa = Octa; pl.CS = CS;
  return true;
}
// 0: the list runs on the slab kernel, with *ngroups descriptor groups, *outputs_per_cta destination groups per CTA and
// *siblings CTAs per tile sequence; -1: it is declined (per-descriptor kernel).  Needs no device.
extern "C" int kb200_scattering_plan(const kb200_scattering_desc *h, int n, int *ngroups, int *outputs_per_cta, int *siblings) {
  if (!h || n <= 0) return -1;
  SlabPlan pl;
  if (!slab_plan(h, n, sm_count() > 0 ? sm_count() : 148, pl)) return -1;
  if (ngroups) *ngroups = pl.ngroups;
  if (outputs_per_cta) *outputs_per_cta = pl.Octa;
  if (siblings) *siblings = pl.CS;
  return 0;
}

int kb200_scatter_slab_try(const kb200_scattering_desc *h, int n, const void *d_descs, int sigs_layout, double source, cudaStream_t st) {
  const char *env = getenv("KB200_SCATTER_SLAB");
  if (env && env[0] == '0') return -1;
  SlabPlan pl;
  if (!slab_plan(h, n, sm_count(), pl)) return -1;
  const int layout = h[0].layout, Zs = h[0].Zs, Gs = h[0].Gs, M = h[0].M;
  const std::vector<int> &first = pl.first;
  const std::vector<std::vector<int>> &members = pl.members;
  const int ngroups = pl.ngroups, Otot = pl.Otot, K = pl.K, nkc4 = (K + 3) / 4, Kp = 4 * nkc4, Octa = pl.Octa, CS = pl.CS;
  const int WO = Octa / 32, QP = 4;
  int NG = 4;  // consumer groups taking the tiles of a CTA in turn (16 consumer warps in all; 8 with a single group)
  { const char *we = getenv("KB200_SLAB_GROUPS"); if (we && (atoi(we) == 1 || atoi(we) == 2)) NG = atoi(we); }
  const int WN = (NG == 1 ? 8 : 16 / NG) / WO, NT = 16 * WN, KC = sl_stage_rows(NT, NG);

  SlabGeom gm;
  memset(&gm, 0, sizeof(gm));
  gm.M = M; gm.Zs = Zs; gm.K = K; gm.nkc4 = nkc4; gm.Otot = Otot; gm.CS = CS; gm.ngroups = ngroups;
  gm.nst = (K + KC - 1) / KC; gm.ntn = (Zs + NT - 1) / NT; gm.accumulate = h[0].accumulate; gm.source = source;
  const Strides3 ms = strides_dgz(layout, M, Gs, Zs);
  gm.in_b = ms.a;
  if ((long long)M * ngroups * gm.ntn > 0x3fffffffLL) return -1;
  gm.ntiles = M * ngroups * gm.ntn;
  gm.ws_doubles = 3u * nkc4 * 32 * QP * WO;
  gm.nmat = 3; gm.unit = 0;
  long long nslots = sm_count() / CS;
  if (nslots > gm.ntiles) nslots = gm.ntiles;
  gm.nslots = (int)nslots;
  { const char *xe = getenv("KB200_SLAB_EXP"); gm.exp = xe ? atoi(xe) : 0; }

  // host-built tables, uploaded through the descriptor cache: [first desc of each group][oglob][kglob][inrow][orow]
  const size_t n_int = (size_t)ngroups + Otot + Kp;
  const size_t int_bytes = (n_int * sizeof(int) + 15) & ~(size_t)15;
  const size_t ptr_count = (size_t)ngroups * Kp + (size_t)ngroups * Otot;
  std::vector<unsigned char> tab(int_bytes + ptr_count * sizeof(void *));
  int *ti = reinterpret_cast<int *>(tab.data());
  for (int g = 0; g < ngroups; ++g) ti[g] = first[g];
  for (int o = 0; o < Otot; ++o) ti[ngroups + o] = h[members[0][o / Gs]].glower_dst + o % Gs;
  for (int k = 0; k < Kp; ++k) ti[ngroups + Otot + k] = k < K ? h[0].glower_src[k / Gs] + k % Gs : 0;
  const void **tp = reinterpret_cast<const void **>(tab.data() + int_bytes);
  for (int g = 0; g < ngroups; ++g) {
    const kb200_scattering_desc &f = h[first[g]];
    for (int k = 0; k < Kp; ++k) tp[(size_t)g * Kp + k] = k < K ? (const void *)(f.phi_src[k / Gs] + (long long)(k % Gs) * ms.g) : nullptr;
  }
  const void **to = tp + (size_t)ngroups * Kp;
  for (int g = 0; g < ngroups; ++g)
    for (int o = 0; o < Otot; ++o) to[(size_t)g * Otot + o] = (const void *)(h[members[g][o / Gs]].phi_out + (long long)(o % Gs) * ms.g);
  const void *d_tab = nullptr;
  int rc = device_descs(tab.data(), tab.size(), &d_tab, st);
  if (rc) return rc;
  const int *d_int = reinterpret_cast<const int *>(d_tab);
  const unsigned char *d_ptr = reinterpret_cast<const unsigned char *>(d_tab) + int_bytes;

  // scratch: fractions, matrices, progress counters
  const size_t frac_b = (size_t)ngroups * 3 * Zs * sizeof(double);
  const size_t wg_b = (size_t)h[0].L1 * CS * gm.ws_doubles * sizeof(double);
  const size_t prog_b = ((size_t)gm.nslots * CS * 4 * sizeof(unsigned) + 255) & ~(size_t)255;
  const size_t need = frac_b + wg_b + prog_b;
  if (g_slab_scratch_bytes < need) {
    if (g_slab_scratch) { KB_CUDA(cudaDeviceSynchronize()); cudaFree(g_slab_scratch); g_slab_scratch = nullptr; g_slab_scratch_bytes = 0; }
    KB_CUDA(cudaMalloc(&g_slab_scratch, need));
    g_slab_scratch_bytes = need;
  }
  unsigned char *sc = reinterpret_cast<unsigned char *>(g_slab_scratch);
  SlabTables tb;
  tb.frac = reinterpret_cast<const double *>(sc);
  tb.wg = reinterpret_cast<const double *>(sc + frac_b);
  tb.progress = reinterpret_cast<unsigned *>(sc + frac_b + wg_b);
  tb.inrow = reinterpret_cast<const double *const *>(d_ptr);
  tb.orow = reinterpret_cast<double *const *>(d_ptr + (size_t)ngroups * Kp * sizeof(void *));
  tb.m2l = h[0].moment_to_legendre;

  slab_fractions_kernel<<<dim3((Zs + 255) / 256, ngroups, 1), 256, 0, st>>>((const kb200_scattering_desc *)d_descs, d_int, Zs,
                                                                             reinterpret_cast<double *>(sc));
  rc = post_launch("slab_fractions");
  if (rc) return rc;
  const Strides4 ss = strides_sigs(sigs_layout >= 0 ? sigs_layout : layout, h[0].L1, h[0].G);
  const long long wtotal = (long long)h[0].L1 * CS * gm.ws_doubles;
  slab_matrices_kernel<<<(unsigned)((wtotal + 255) / 256 > 1184 ? 1184 : (wtotal + 255) / 256), 256, 0, st>>>(
      h[0].sigs, ss, d_int + ngroups, d_int + ngroups + Otot, h[0].L1, CS, WO, QP, nkc4, K, reinterpret_cast<double *>(sc + frac_b));
  rc = post_launch("slab_matrices");
  if (rc) return rc;
  if (CS > 1) KB_CUDA(cudaMemsetAsync(tb.progress, 0, prog_b, st));

  const size_t smem = (size_t)gm.ws_doubles * 8 + NG * (SL_STAGES * (size_t)sl_stage_bytes(NT, NG) + 16 * SL_STAGES);
  const unsigned grid = (unsigned)(gm.nslots * CS);
#define SL_LAUNCH(WO_, WN_, NG_)                                                                                                      \
  do {                                                                                                                              \
    KB_CUDA(cudaFuncSetAttribute(scatter_slab_kernel<4, WO_, WN_, NG_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    scatter_slab_kernel<4, WO_, WN_, NG_><<<grid, 32 * (WO_ * WN_ + 1) * NG_, smem, st>>>(gm, tb);                                  \
  } while (0)
  if (WO == 2) {
    if (NG == 4) SL_LAUNCH(2, 2, 4);
    else if (NG == 2) SL_LAUNCH(2, 4, 2);
    else SL_LAUNCH(2, 4, 1);
  } else {
    if (NG == 4) SL_LAUNCH(1, 4, 4);
    else if (NG == 2) SL_LAUNCH(1, 8, 2);
    else SL_LAUNCH(1, 8, 1);
  }
#undef SL_LAUNCH
  return post_launch("scatter_slab");
}

static unsigned long long g_gemm_slab_launches = 0;
extern "C" unsigned long long kb200_lplustimes_slab_launches(void) { return g_gemm_slab_launches; }

// Plain product on the slab kernel (one matrix, unit fractions, one "moment" batch): LPlusTimes where the reduction is long
// enough for it (M >= 36: the fp64-bound regime of BASELINE config 3) and the columns of a chunk are one contiguous run.
// out rows: o = direction over all sets; in rows: k = moment.  h_ptrs: per chunk the phi_out pointer, then the nsets rhs
// pointers.  Returns 0 if handled, -1 if this path does not apply, >0 on error.
int kb200_gemm_slab_try(int M, int Ds, int nsets, int accumulate, long long N, long long in_r, long long out_r, const void *d_views, int n,
                        const void *const *h_ptrs, int same_w, cudaStream_t st) {
  const char *env = getenv("KB200_GEMM_SLAB");
  if (env && env[0] == '0') return -1;
  const int O = nsets * Ds, K = M, nkc4 = (K + 3) / 4, Kp = 4 * nkc4;
  if (!same_w || n <= 0 || N % 4 != 0 || O % 32 != 0 || N > 0x3fffffffLL) return -1;
  if ((in_r * 8) % 16 != 0 || (out_r * 8) % 32 != 0) return -1;
  for (int i = 0; i < n; ++i)
    for (int s = 0; s < nsets; ++s)
      if (((uintptr_t)h_ptrs[(size_t)i * (nsets + 1) + 1 + s] & 31) != 0) return -1;
  const size_t smem_cap = 227 * 1024 - 4 * SL_STAGES * (size_t)sl_stage_bytes(64, 4) - 256;
  int Octa = 0;
  if (O % 64 == 0 && (size_t)Kp * 64 * 8 <= smem_cap) Octa = 64;
  else if ((size_t)Kp * 32 * 8 <= smem_cap) Octa = 32;
  else return -1;
  const int CS = O / Octa;
  if (CS > 32 || CS > sm_count()) return -1;
  const int WO = Octa / 32, QP = 4, NG = 4;
  const int WN = 16 / NG / WO, NT = 16 * WN, KC = sl_stage_rows(NT, NG);
  SlabGeom gm;
  memset(&gm, 0, sizeof(gm));
  gm.M = 1; gm.Zs = (int)N; gm.K = K; gm.nkc4 = nkc4; gm.Otot = O; gm.CS = CS; gm.ngroups = n;
  gm.nst = (K + KC - 1) / KC; gm.ntn = (int)((N + NT - 1) / NT); gm.accumulate = accumulate; gm.source = 0.0;
  gm.in_b = 0;
  if ((long long)n * gm.ntn > 0x3fffffffLL) return -1;
  gm.ntiles = n * gm.ntn;
  gm.nmat = 1; gm.unit = 1;
  gm.ws_doubles = (unsigned)nkc4 * 32 * QP * WO;
  long long nslots = sm_count() / CS;
  if (nslots > gm.ntiles) nslots = gm.ntiles;
  gm.nslots = (int)nslots;

  std::vector<const void *> tab((size_t)n * Kp + (size_t)n * O);
  for (int i = 0; i < n; ++i) {
    const void *const *p = h_ptrs + (size_t)i * (nsets + 1);
    for (int k = 0; k < Kp; ++k) tab[(size_t)i * Kp + k] = k < K ? (const void *)((const double *)p[0] + (long long)k * in_r) : nullptr;
    for (int o = 0; o < O; ++o)
      tab[(size_t)n * Kp + (size_t)i * O + o] = (const void *)((const double *)p[1 + o / Ds] + (long long)(o % Ds) * out_r);
  }
  const void *d_tab = nullptr;
  int rc = device_descs(tab.data(), tab.size() * sizeof(void *), &d_tab, st);
  if (rc) return rc;
  const size_t wg_b = (size_t)CS * gm.ws_doubles * sizeof(double);
  const size_t prog_b = ((size_t)gm.nslots * CS * 4 * sizeof(unsigned) + 255) & ~(size_t)255;
  const size_t need = wg_b + prog_b;
  if (g_slab_scratch_bytes < need) {
    if (g_slab_scratch) { KB_CUDA(cudaDeviceSynchronize()); cudaFree(g_slab_scratch); g_slab_scratch = nullptr; g_slab_scratch_bytes = 0; }
    KB_CUDA(cudaMalloc(&g_slab_scratch, need));
    g_slab_scratch_bytes = need;
  }
  unsigned char *sc = reinterpret_cast<unsigned char *>(g_slab_scratch);
  SlabTables tb;
  memset(&tb, 0, sizeof(tb));
  tb.wg = reinterpret_cast<const double *>(sc);
  tb.progress = reinterpret_cast<unsigned *>(sc + wg_b);
  tb.inrow = reinterpret_cast<const double *const *>(d_tab);
  tb.orow = (double *const *)((const void *const *)d_tab + (size_t)n * Kp);
  const long long wtotal = (long long)CS * gm.ws_doubles;
  gemm_matrices_kernel<<<(unsigned)((wtotal + 255) / 256 > 1184 ? 1184 : (wtotal + 255) / 256), 256, 0, st>>>(
      (const GemmDescK *)d_views, Ds, M, O, K, CS, WO, QP, nkc4, reinterpret_cast<double *>(sc));
  rc = post_launch("gemm_matrices");
  if (rc) return rc;
  if (CS > 1) KB_CUDA(cudaMemsetAsync(tb.progress, 0, prog_b, st));
  const size_t smem = (size_t)gm.ws_doubles * 8 + NG * (SL_STAGES * (size_t)sl_stage_bytes(NT, NG) + 16 * SL_STAGES);
  const unsigned grid = (unsigned)(gm.nslots * CS);
  if (WO == 2) {
    KB_CUDA(cudaFuncSetAttribute(scatter_slab_kernel<4, 2, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    scatter_slab_kernel<4, 2, 2, 4><<<grid, 32 * (2 * 2 + 1) * 4, smem, st>>>(gm, tb);
  } else {
    KB_CUDA(cudaFuncSetAttribute(scatter_slab_kernel<4, 1, 4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    scatter_slab_kernel<4, 1, 4, 4><<<grid, 32 * (1 * 4 + 1) * 4, smem, st>>>(gm, tb);
  }
  ++g_gemm_slab_launches;
  return post_launch("gemm_slab");
}
