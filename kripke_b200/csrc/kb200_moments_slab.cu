// LTimes at M = 25 * c moments (Legendre order 4: c = 1; order 9: c = 4 sibling CTAs of 25 moments each that walk the same tiles
// in advisory lockstep, so the siblings' reads of a tile meet in L2) for the nestings whose (group, zone) plane is contiguous
// per direction (DGZ, DZG, GDZ), default arithmetic.  Reference: src/Kripke/Kernel/LTimes.cpp:54-65.
//
//   phi(nm, n) (+)= sum_k ell(nm, k) * psi(k, n)        k = all directions of all sets, n = a run of (group, zone) columns
//
// The same contraction as moments_mma_kernel<3, 2, 1, ...> (24 moments on the fp64 tensor pipe, the 25th by DFMA on the
// B-fragment layout), restructured like kb200_scatter_slab.cu: at M = 25 LTimes needs 4.35 ms of the fp64 pipe and 4.6 ms
// of HBM time at config 2, so it can only approach either roofline if the two overlap completely.  The per-warp cp.async
// pipelines of the older kernel kept the pipe 72 % busy; here
// * four groups of four consumer warps, each group with its own producer warp and its own ring of bulk-TMA stages
//   ([KC directions][64 columns], rows 32 (mod 128) bytes apart: conflict-free B fragments), take the tiles of a CTA in
//   turn, so a group's epilogue overlaps the other groups' DMMAs and nobody issues a copy between two DMMAs;
// * both operands come in by 128-bit shared loads: columns 2j/2j+1 are the two B fragments of lane group j, the A
//   fragments of the three moment blocks and the weight of the 25th moment sit in two lane-contiguous pairs;
// * an accumulator lane ends up with four consecutive columns of a moment row = one 32-byte store.
// Tiles are flattened over the descriptors of a call (phi chunks), so many small chunks still fill every SM evenly.
// Anything it does not cover returns -1 and moments_mma_kernel runs instead.
// (LPlusTimes was tried in the same structure -- B fragments of a warp's 16 columns in registers, direction blocks in pairs,
// 32-byte stores -- and measured 7.1 ms against the 6.2 ms of moments_mma_kernel at config 2: that kernel is bound by its
// 25.7 GB store stream, ncu showed the warps waiting on store-data registers, and 512-byte row pieces written by four
// warps at their own pace drain more slowly than the older kernel's 1 KB pieces; profiles/r02bf_*.  It was removed again.)
#include <vector>
#include "kb200_common.cuh"

namespace kb200 {

struct MomentsDescK {  // same as in kb200_moments.cu
  const double *const *w;
  const double *const *in;
  double *const *out;
};

struct MsGeom {
  int M, Ds, K, nkc4, nst, accumulate, adj;
  int CS, nslots;             // sibling CTAs per tile sequence (25 moments each: M = 25 * CS), tile sequences of the launch
  int ntn, per_desc, ntiles;  // column tiles per batch, tiles per descriptor, tiles of the call
  long long N, in_b, out_b;   // columns per batch; batch strides of psi and phi
};
struct MsTables {
  const double *const *inrow;  // [ndesc][nst*KC] direction rows of psi (batch 0, column 0), nullptr beyond K
  double *const *orow;         // [ndesc][M]      moment rows of phi
  unsigned *progress;          // [nslots][CS][groups] tiles issued by each sibling's producers (CS > 1 only)
};

__device__ __forceinline__ unsigned ms_smem(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ms_mb_init(unsigned a, unsigned n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(n) : "memory"); }
__device__ __forceinline__ void ms_mb_arrive(unsigned a) { asm volatile("{ .reg .b64 t; mbarrier.arrive.shared::cta.b64 t, [%0]; }" ::"r"(a) : "memory"); }
__device__ __forceinline__ void ms_mb_expect_tx(unsigned a, unsigned bytes) {
  asm volatile("{ .reg .b64 t; mbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1; }" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ms_mb_wait(unsigned a, unsigned parity) {
  asm volatile("{\n .reg .pred p;\n MSW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @!p bra MSW;\n}" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void ms_bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ double2 ms_lds128(unsigned a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void ms_dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int MS_NG = 4, MS_WN = 4, MS_NT = 16 * MS_WN, MS_KC = 24, MS_STAGES = 3;
constexpr unsigned MS_PITCH = 8 * MS_NT + 32, MS_STAGE_BYTES = MS_KC * MS_PITCH;
constexpr int MS_THREADS = 32 * MS_NG * (MS_WN + 1);

__global__ void __launch_bounds__(MS_THREADS, 1) ltimes_slab_kernel(const MomentsDescK *__restrict__ descs, MsGeom gm, MsTables tb) {
  constexpr int NW = MS_NG * MS_WN;
  extern __shared__ __align__(128) unsigned char msl[];
  double *Ws = reinterpret_cast<double *>(msl);  // [nkc4][2 pairs][32 lanes][2]: (A0, A1), (A2, weight of moment 24)
  const unsigned ws_b = (unsigned)gm.nkc4 * 1024u;
  const unsigned stage0 = ms_smem(msl) + ws_b;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = warp < NW ? warp / MS_WN : warp - NW;  // consumer group of this warp / the group this producer feeds
  // M = 25 * CS moments: CS sibling CTAs walk the same tiles, each with its own 25 moments (24 + 1) of the weights
  const int slot = (int)blockIdx.x / gm.CS, sib = (int)blockIdx.x - slot * gm.CS;
  const unsigned ring0 = stage0 + (unsigned)grp * (MS_STAGES * MS_STAGE_BYTES);
  const unsigned full0 = stage0 + MS_NG * MS_STAGES * MS_STAGE_BYTES + (unsigned)grp * (16u * MS_STAGES), empty0 = full0 + 8u * MS_STAGES;

  if (lane < MS_STAGES && warp >= NW) {
    ms_mb_init(full0 + 8u * lane, 1u);
    ms_mb_init(empty0 + 8u * lane, (unsigned)MS_WN);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  // fragment-major weights (every descriptor of the call has the same ell tables: checked by the host)
  {
    const MomentsDescK d0 = descs[0];
    for (int idx = threadIdx.x; idx < gm.nkc4 * 128; idx += MS_THREADS) {
      const int slot = idx & 1, l = (idx >> 1) & 31, pr = (idx >> 6) & 1, kc = idx >> 7;
      const int a = 2 * pr + slot, k = 4 * kc + (l & 3);
      double v = 0.0;
      if (k < gm.K) {
        const int s = k / gm.Ds, d = k - s * gm.Ds;
        const int o = 25 * sib + ((a < 3) ? 8 * a + (l >> 2) : 24);
        v = d0.w[s][(size_t)d * gm.M + o];
      }
      Ws[idx] = v;
    }
  }
  // rows of a stage beyond K and columns beyond the run are never written by a copy: they must hold finite values
  for (unsigned i = threadIdx.x; i < MS_NG * MS_STAGES * MS_STAGE_BYTES / 16; i += MS_THREADS)
    reinterpret_cast<double2 *>(msl + ws_b)[i] = make_double2(0.0, 0.0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();

  // this group's tiles: the groups of a CTA work on neighbouring tiles (KB200_MS_ADJ=1; default: a grid apart, measured equal or better)
  const int first = gm.adj ? MS_NG * slot + grp : slot + grp * gm.nslots, step = MS_NG * gm.nslots;
  if (warp >= NW) {
    // ---- producer of group grp: one bulk copy per direction row of a stage ----
    unsigned it = 0, done = 0;
    bool lockstep = gm.CS > 1;
    for (int t = first; t < gm.ntiles; t += step, ++done) {
      const int dsc = t / gm.per_desc, r = t - dsc * gm.per_desc, b = r / gm.ntn, tn = r - b * gm.ntn;
      if (lockstep && done > 0) {  // advisory: stay within a tile of the siblings so that their reads of this tile meet in L2
        bool ok = true;
        if (lane < gm.CS && lane != sib) {
          const volatile unsigned *p = tb.progress + ((size_t)slot * gm.CS + lane) * MS_NG + grp;
          int polls = 0;
          while (*p + 1u < done && ++polls < 512) {}
          ok = polls < 512;
        }
        if (!__all_sync(0xffffffffu, ok)) lockstep = false;
      }
      const long long n0 = (long long)tn * MS_NT;
      const long long rem = gm.N - n0;
      const unsigned rb = 8u * (unsigned)(rem < MS_NT ? rem : MS_NT);
      const double *const *rows = tb.inrow + (size_t)dsc * gm.nst * MS_KC;
      const long long off = (long long)b * gm.in_b + n0;
      for (int st = 0; st < gm.nst; ++st, ++it) {
        const unsigned s = it % MS_STAGES, ph = (it / MS_STAGES) & 1u;
        ms_mb_wait(empty0 + 8u * s, ph ^ 1u);
        const int nrows = min(MS_KC, gm.K - st * MS_KC);
        if (lane == 0) ms_mb_expect_tx(full0 + 8u * s, (unsigned)nrows * rb);
        __syncwarp();
        if (lane < nrows)
          ms_bulk_g2s(ring0 + s * MS_STAGE_BYTES + (unsigned)lane * MS_PITCH, rows[st * MS_KC + lane] + off, rb, full0 + 8u * s);
      }
      if (gm.CS > 1 && lane == 0) {
        __threadfence();
        *(volatile unsigned *)(tb.progress + ((size_t)slot * gm.CS + sib) * MS_NG + grp) = done + 1u;
      }
    }
    return;
  }

  // ---- consumers ----
  const int wn = warp - grp * MS_WN;
  const int col0 = wn * 16, jq = lane >> 2, kq = lane & 3;
  const unsigned a_lane = ms_smem(Ws) + 16u * lane;
  const unsigned b_lane = (unsigned)kq * MS_PITCH + (unsigned)(col0 + 2 * jq) * 8u;
  unsigned it = 0;
  for (int t = first; t < gm.ntiles; t += step) {
    const int dsc = t / gm.per_desc, r = t - dsc * gm.per_desc, b = r / gm.ntn, tn = r - b * gm.ntn;
    double acc[3][2][2];
#pragma unroll
    for (int a = 0; a < 3; ++a) acc[a][0][0] = acc[a][0][1] = acc[a][1][0] = acc[a][1][1] = 0.0;
    double px0 = 0.0, px1 = 0.0;  // moment 24 at columns 2jq, 2jq+1: partial sums over this lane's k residue class

    for (int st = 0; st < gm.nst; ++st, ++it) {
      const unsigned s = it % MS_STAGES, ph = (it / MS_STAGES) & 1u;
      ms_mb_wait(full0 + 8u * s, ph);
      const int kc_lo = st * (MS_KC / 4), nkc = min(MS_KC / 4, gm.nkc4 - kc_lo);
      unsigned ap = a_lane + (unsigned)kc_lo * 1024u;
      unsigned bp = ring0 + s * MS_STAGE_BYTES + b_lane;
#pragma unroll 3
      for (int kc = 0; kc < nkc; ++kc, ap += 1024u, bp += 4u * MS_PITCH) {
        const double2 bv = ms_lds128(bp);
        const double2 a01 = ms_lds128(ap), a2x = ms_lds128(ap + 512u);
        ms_dmma(acc[0][0][0], acc[0][0][1], a01.x, bv.x);
        ms_dmma(acc[0][1][0], acc[0][1][1], a01.x, bv.y);
        ms_dmma(acc[1][0][0], acc[1][0][1], a01.y, bv.x);
        ms_dmma(acc[1][1][0], acc[1][1][1], a01.y, bv.y);
        ms_dmma(acc[2][0][0], acc[2][0][1], a2x.x, bv.x);
        ms_dmma(acc[2][1][0], acc[2][1][1], a2x.x, bv.y);
        px0 = fma(a2x.y, bv.x, px0);
        px1 = fma(a2x.y, bv.y, px1);
      }
      __syncwarp();
      if (lane == 0) ms_mb_arrive(empty0 + 8u * s);
    }

    // epilogue: lane holds moments 8a + jq at the four consecutive columns col0 + 4*kq + {0,1,2,3} ...
    double *const *orow = tb.orow + (size_t)dsc * gm.M + 25 * sib;
    const long long ob = (long long)b * gm.out_b + (long long)tn * MS_NT + col0;
    const long long n4 = (long long)tn * MS_NT + col0 + 4 * kq;
    if (n4 < gm.N) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        double *p = orow[8 * a + jq] + ob + 4 * kq;
        double v0 = acc[a][0][0], v1 = acc[a][1][0], v2 = acc[a][0][1], v3 = acc[a][1][1];
        if (gm.accumulate) {
          double o0, o1, o2, o3;
          asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(o0), "=d"(o1), "=d"(o2), "=d"(o3) : "l"(p) : "memory");
          v0 += o0; v1 += o1; v2 += o2; v3 += o3;
        }
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v0), "d"(v1), "d"(v2), "d"(v3) : "memory");
      }
    }
    // ... and the k residue class kq of moment 24 at columns col0 + 2*jq + {0,1}: sum the four classes, class 0 stores
    px0 += __shfl_xor_sync(0xffffffffu, px0, 1);
    px1 += __shfl_xor_sync(0xffffffffu, px1, 1);
    px0 += __shfl_xor_sync(0xffffffffu, px0, 2);
    px1 += __shfl_xor_sync(0xffffffffu, px1, 2);
    if (kq == 0 && (long long)tn * MS_NT + col0 + 2 * jq < gm.N) {
      double2 *p = reinterpret_cast<double2 *>(orow[24] + ob + 2 * jq);
      double2 v = make_double2(px0, px1);
      if (gm.accumulate) { const double2 o = *p; v.x += o.x; v.y += o.y; }
      *p = v;
    }
  }
}


}  // namespace kb200

using namespace kb200;

static unsigned long long g_ltimes_slab_launches = 0;
extern "C" unsigned long long kb200_ltimes_slab_launches(void) { return g_ltimes_slab_launches; }

// LTimes, M = 25, contiguous (group, zone) columns.  h_ptrs: per descriptor the nsets psi chunks, then the phi chunk.
// Returns 0 if handled, -1 if this path does not apply, >0 on error.
int kb200_ltimes_slab_try(int M, int Ds, int nsets, int accumulate, long long B, long long N, long long in_b, long long in_r,
                          long long out_b, long long out_r, const void *d_views, int n, const void *const *h_ptrs, int same_w,
                          cudaStream_t st) {
  const char *env = getenv("KB200_LTIMES_SLAB");
  if (env && env[0] == '0') return -1;
  if (M % 25 != 0 || M > 200 || !same_w || N % 4 != 0 || n <= 0) return -1;
  const int CS = M / 25;
  if (CS > sm_count()) return -1;
  const int K = nsets * Ds, nkc4 = (K + 3) / 4, nst = (K + MS_KC - 1) / MS_KC;
  const size_t smem = (size_t)nkc4 * 1024 + (size_t)MS_NG * (MS_STAGES * MS_STAGE_BYTES + 16 * MS_STAGES);
  if (smem > 227 * 1024) return -1;
  if ((in_b * 8) % 16 != 0 || (in_r * 8) % 16 != 0 || (out_b * 8) % 32 != 0 || (out_r * 8) % 32 != 0) return -1;
  for (int i = 0; i < n; ++i) {
    if (((uintptr_t)h_ptrs[(size_t)i * (nsets + 1) + nsets] & 31) != 0) return -1;
  }
  MsGeom gm;
  memset(&gm, 0, sizeof(gm));
  gm.M = M; gm.Ds = Ds; gm.K = K; gm.nkc4 = nkc4; gm.nst = nst; gm.accumulate = accumulate;
  gm.N = N; gm.in_b = in_b; gm.out_b = out_b;
  { const char *ae = getenv("KB200_MS_ADJ"); gm.adj = (ae && ae[0] == '1'); }
  const long long ntn = (N + MS_NT - 1) / MS_NT;
  if (ntn * B * n > 0x3fffffffLL) return -1;
  gm.ntn = (int)ntn; gm.per_desc = (int)(ntn * B); gm.ntiles = gm.per_desc * n;

  // host-built row tables, uploaded through the descriptor cache
  const size_t Kp = (size_t)nst * MS_KC;
  std::vector<const void *> tab((size_t)n * Kp + (size_t)n * M);
  for (int i = 0; i < n; ++i) {
    const void *const *p = h_ptrs + (size_t)i * (nsets + 1);
    for (size_t k = 0; k < Kp; ++k)
      tab[(size_t)i * Kp + k] = k < (size_t)K ? (const void *)((const double *)p[k / Ds] + (long long)(k % Ds) * in_r) : nullptr;
    for (int o = 0; o < M; ++o) tab[(size_t)n * Kp + (size_t)i * M + o] = (const void *)((const double *)p[nsets] + (long long)o * out_r);
  }
  const void *d_tab = nullptr;
  int rc = device_descs(tab.data(), tab.size() * sizeof(void *), &d_tab, st);
  if (rc) return rc;
  MsTables tb;
  tb.inrow = reinterpret_cast<const double *const *>(d_tab);
  tb.orow = (double *const *)((const void *const *)d_tab + (size_t)n * Kp);

  KB_CUDA(cudaFuncSetAttribute(ltimes_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int nslots = sm_count() / CS;
  if (nslots > gm.ntiles) nslots = gm.ntiles;
  gm.CS = CS; gm.nslots = nslots;
  tb.progress = nullptr;
  if (CS > 1) {  // advisory progress counters of the siblings
    static unsigned *d_prog = nullptr;
    const size_t pb = (size_t)sm_count() * MS_NG * sizeof(unsigned);
    if (!d_prog) KB_CUDA(cudaMalloc(&d_prog, pb));
    KB_CUDA(cudaMemsetAsync(d_prog, 0, pb, st));
    tb.progress = d_prog;
  }
  ltimes_slab_kernel<<<nslots * CS, MS_THREADS, smem, st>>>((const MomentsDescK *)d_views, gm, tb);
  ++g_ltimes_slab_launches;
  return post_launch("ltimes_slab");
}
