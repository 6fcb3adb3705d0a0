// LTimes and LPlusTimes for sm_100a: the discrete-to-moments contraction and its transpose.
//
//   LTimes     phi(nm,g,z)   (+)= sum_q sum_d ell_q(nm,d)      * psi_q(d,g,z)   (Kernel/LTimes.cpp:54-65)
//   LPlusTimes rhs_q(d,g,z)  (+)= sum_nm    ell_plus_q(d,nm)   * phi_out(nm,g,z) (Kernel/LPlusTimes.cpp:49-60)
//
// Both are the same tall-skinny fp64 GEMM  out[o][n] = sum_k W[o][k] * in[k][n]  with the long
// dimension n = (group, zone):
//   LTimes     o = moment (M),            k = (dirset q, direction d)
//   LPlusTimes o = (dirset q, direction), k = moment
// The weights (ell / ell_plus, a few KB) live in shared memory and are read with broadcast 128-bit
// loads; the streamed operand is read with coalesced 128-bit fp64 loads exactly once per output
// tile; each thread keeps an [OT x 2] accumulator tile in registers (plain DFMA: on B200 the fp64
// tensor-core rate equals the DFMA rate, and M=25 does not tile into m8n8k4 without 22% padding).
// All direction sets that share a phi chunk are fused into one launch, so phi is written once
// (the reference RMWs phi once per direction set and once per d on its CUDA policy,
// Arch/LTimes.h:207-222).
//
// Two storage classes (src/Kripke/VarTypes.h:73-101):
//   "column" class  DGZ, DZG, GDZ, ZDG : the (d | nm) index is NOT the fastest one; for fixed batch b
//                   in[b][k][n] / out[b][o][n] with n contiguous            -> kernel moments_col
//   "row" class     GZD, ZGD           : d / nm is the fastest index: in[r][k], out[r][o] with
//                   r = flat (g,z) row                                        -> kernel moments_row
//
// Summation order (bitwise parity in EXACT mode): ascending q, then
// ascending d (LTimes) / ascending nm (LPlusTimes), separate multiply and add.
#include "kb200_common.cuh"
#include <stdlib.h>
#include <vector>

namespace kb200 {

template <bool EXACT>
__device__ __forceinline__ double mad(double a, double b, double c) {
  if (EXACT) return __dadd_rn(__dmul_rn(a, b), c);
  return fma(a, b, c);
}

struct MomentsGeom {
  int mode;          // 0 = LTimes, 1 = LPlusTimes
  int M, Ds, nsets;  // moments, directions per set, direction sets
  int accumulate;
  int OT;            // outputs per tile (<= OTP)
  int ntiles_o;      // LTimes: tiles over M ; LPlusTimes: tiles per direction set
  // column class
  long long B, N;                  // batches, contiguous run
  long long flux_b, flux_a;        // psi/rhs strides: batch, direction
  long long mom_b, mom_a;          // phi/phi_out strides: batch, moment
  // row class
  long long R;                     // rows
};

// ------------------------------------------------------------------------------------------------
// column class: thread = VEC consecutive n, block = 128 threads, grid = (n-blocks, o-tiles, descs)
// ------------------------------------------------------------------------------------------------
template <int OTP, int VEC, bool EXACT, typename Desc>
__global__ void __launch_bounds__(128) moments_col_kernel(const Desc *__restrict__ descs, MomentsGeom gm) {
  extern __shared__ __align__(16) double w_s[];  // [K][OTP]
  const Desc &dsc = descs[blockIdx.z];
  const int tile = blockIdx.y;
  const int M = gm.M, Ds = gm.Ds;
  int K, o0, oq = 0;
  if (gm.mode == 0) { K = gm.nsets * Ds; o0 = tile * gm.OT; }
  else { K = M; oq = tile / gm.ntiles_o; o0 = (tile % gm.ntiles_o) * gm.OT; }
  const int o_hi = (gm.mode == 0) ? M : Ds;
  const int on = min(gm.OT, o_hi - o0);  // valid outputs in this tile

  // stage the weight tile, zero padded: w_s[k][oo] = W[o0+oo][k]
  for (int idx = threadIdx.x; idx < K * OTP; idx += blockDim.x) {
    int k = idx / OTP, oo = idx - k * OTP;
    double v = 0.0;
    if (oo < on) {
      if (gm.mode == 0) {
        int q = k / Ds, d = k - q * Ds;
        v = ((const double *const *)dsc.w)[q][(size_t)d * M + (o0 + oo)];
      } else {
        v = ((const double *const *)dsc.w)[oq][(size_t)(o0 + oo) * M + k];
      }
    }
    w_s[idx] = v;
  }
  __syncthreads();

  const long long nvec = gm.N / VEC;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= gm.B * nvec) return;
  const long long b = p / nvec, n = (p - b * nvec) * VEC;

  double acc[OTP][VEC];
#pragma unroll
  for (int i = 0; i < OTP; ++i)
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[i][v] = 0.0;

  const long long in_ks = (gm.mode == 0) ? gm.flux_a : gm.mom_a;
  const long long in_b = (gm.mode == 0) ? gm.flux_b : gm.mom_b;
  const int nchunks = (gm.mode == 0) ? gm.nsets : 1;
  const int kper = (gm.mode == 0) ? Ds : M;
  int kk = 0;
  for (int c = 0; c < nchunks; ++c) {
    const double *__restrict__ src = ((const double *const *)dsc.in)[c] + b * in_b + n;
    int kc = 0;
    for (; kc + 4 <= kper; kc += 4, kk += 4) {
      double x[4][VEC];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (VEC == 2) {
          double2 t = __ldg(reinterpret_cast<const double2 *>(src + (long long)(kc + u) * in_ks));
          x[u][0] = t.x; x[u][VEC - 1] = t.y;
        } else {
          x[u][0] = __ldg(src + (long long)(kc + u) * in_ks);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double2 *w2 = reinterpret_cast<const double2 *>(w_s + (size_t)(kk + u) * OTP);
#pragma unroll
        for (int i = 0; i < OTP / 2; ++i) {
          double2 w = w2[i];
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            acc[2 * i][v] = mad<EXACT>(w.x, x[u][v], acc[2 * i][v]);
            acc[2 * i + 1][v] = mad<EXACT>(w.y, x[u][v], acc[2 * i + 1][v]);
          }
        }
      }
    }
    for (; kc < kper; ++kc, ++kk) {
      double x[VEC];
      if (VEC == 2) {
        double2 t = __ldg(reinterpret_cast<const double2 *>(src + (long long)kc * in_ks));
        x[0] = t.x; x[VEC - 1] = t.y;
      } else {
        x[0] = __ldg(src + (long long)kc * in_ks);
      }
      const double2 *w2 = reinterpret_cast<const double2 *>(w_s + (size_t)kk * OTP);
#pragma unroll
      for (int i = 0; i < OTP / 2; ++i) {
        double2 w = w2[i];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          acc[2 * i][v] = mad<EXACT>(w.x, x[v], acc[2 * i][v]);
          acc[2 * i + 1][v] = mad<EXACT>(w.y, x[v], acc[2 * i + 1][v]);
        }
      }
    }
  }

  const long long out_os = (gm.mode == 0) ? gm.mom_a : gm.flux_a;
  const long long out_b = (gm.mode == 0) ? gm.mom_b : gm.flux_b;
  double *__restrict__ dst = ((double *const *)dsc.out)[(gm.mode == 0) ? 0 : oq] + b * out_b + (long long)o0 * out_os + n;
#pragma unroll
  for (int i = 0; i < OTP; ++i) {
    if (i < on) {
      double *q = dst + (long long)i * out_os;
      if (VEC == 2) {
        double2 r = make_double2(acc[i][0], acc[i][VEC - 1]);
        if (gm.accumulate) {
          double2 o = *reinterpret_cast<double2 *>(q);
          // reference order: phi += ell*psi term by term starting from the old value; with an
          // old value of 0 (the only case the solver produces) both orders are identical
          r.x = __dadd_rn(o.x, r.x); r.y = __dadd_rn(o.y, r.y);
        }
        *reinterpret_cast<double2 *>(q) = r;
      } else {
        double r = acc[i][0];
        if (gm.accumulate) r = __dadd_rn(*q, r);
        *q = r;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// row class (GZD, ZGD): in[r][Kc] rows per chunk, out[r][Oc] rows per chunk.  A block owns RT rows;
// the streamed tile is staged through shared memory with coalesced 128-bit loads (rows padded to
// an odd stride so that thread=row accesses are bank-conflict free), each thread accumulates two
// rows x OT outputs in registers, and the output tile goes back through shared memory so the
// stores are coalesced as well.
// ------------------------------------------------------------------------------------------------
template <int OTP, bool EXACT, typename Desc, int RT>  // RT rows per block: 256 (2 per thread), 128, or 64 where the tile must shrink
__global__ void __launch_bounds__(128) moments_row_kernel(const Desc *__restrict__ descs, MomentsGeom gm) {
  extern __shared__ __align__(16) double smem[];
  constexpr bool HAS_B = RT == 256;
  const Desc &dsc = descs[blockIdx.y];
  const int M = gm.M, Ds = gm.Ds;
  const int Kc = (gm.mode == 0) ? Ds : M;        // inputs per row per in-chunk
  const int Oc = (gm.mode == 0) ? M : Ds;        // outputs per row per out-chunk
  const int KcP = Kc | 1;                        // odd row stride
  const int OcP = Oc | 1;
  double *in_s = smem;                           // [RT][KcP]
  double *out_s = in_s + (size_t)RT * KcP;       // [RT][OcP]
  double *w_s = out_s + (size_t)RT * OcP;        // [Kc][OTP]
  const long long r0 = (long long)blockIdx.x * RT;
  const int nrows = (int)min((long long)RT, gm.R - r0);
  const int n_in_chunks = (gm.mode == 0) ? gm.nsets : 1;
  const int n_out_chunks = (gm.mode == 0) ? 1 : gm.nsets;
  const int tiles_per_chunk = gm.ntiles_o;
  const bool has_a = (int)threadIdx.x < RT;               // RT = 64: the upper half of the block only helps with the copies
  const int ra = has_a ? (int)threadIdx.x : 0, rb = HAS_B ? (int)threadIdx.x + 128 : ra;

  for (int oc = 0; oc < n_out_chunks; ++oc) {
    for (int t = 0; t < tiles_per_chunk; ++t) {
      const int o0 = t * gm.OT;
      const int on = min(gm.OT, Oc - o0);
      double acc[OTP][2];
#pragma unroll
      for (int i = 0; i < OTP; ++i) acc[i][0] = acc[i][1] = 0.0;

      for (int ic = 0; ic < n_in_chunks; ++ic) {
        const bool resident = (n_in_chunks == 1) && (oc > 0 || t > 0);
        __syncthreads();  // previous users of in_s / w_s are done
        if (!resident) {
          const double *__restrict__ src = ((const double *const *)dsc.in)[ic] + r0 * Kc;
          const int total = nrows * Kc;
          for (int f = threadIdx.x; f < total; f += blockDim.x) {
            int r = f / Kc, c = f - r * Kc;
            in_s[r * KcP + c] = __ldg(src + f);
          }
        }
        for (int idx = threadIdx.x; idx < Kc * OTP; idx += blockDim.x) {
          int k = idx / OTP, oo = idx - k * OTP;
          double v = 0.0;
          if (oo < on) {
            if (gm.mode == 0) v = ((const double *const *)dsc.w)[ic][(size_t)k * M + (o0 + oo)];
            else v = ((const double *const *)dsc.w)[oc][(size_t)(o0 + oo) * M + k];
          }
          w_s[idx] = v;
        }
        __syncthreads();
        const double *xa = in_s + ra * KcP, *xb = in_s + rb * KcP;
        for (int k = 0; k < Kc; ++k) {
          double va = xa[k], vb = xb[k];
          const double2 *w2 = reinterpret_cast<const double2 *>(w_s + (size_t)k * OTP);
#pragma unroll
          for (int i = 0; i < OTP / 2; ++i) {
            double2 w = w2[i];
            acc[2 * i][0] = mad<EXACT>(w.x, va, acc[2 * i][0]);
            acc[2 * i + 1][0] = mad<EXACT>(w.y, va, acc[2 * i + 1][0]);
            acc[2 * i][1] = mad<EXACT>(w.x, vb, acc[2 * i][1]);
            acc[2 * i + 1][1] = mad<EXACT>(w.y, vb, acc[2 * i + 1][1]);
          }
        }
      }
      // scatter this output tile into the staging rows
#pragma unroll
      for (int i = 0; i < OTP; ++i)
        if (i < on) {
          if (has_a) out_s[ra * OcP + o0 + i] = acc[i][0];
          if (HAS_B) out_s[rb * OcP + o0 + i] = acc[i][1];
        }
    }
    __syncthreads();
    double *__restrict__ dst = ((double *const *)dsc.out)[oc] + r0 * Oc;
    const int total = nrows * Oc;
    for (int f = threadIdx.x; f < total; f += blockDim.x) {
      int r = f / Oc, c = f - r * Oc;
      double v = out_s[r * OcP + c];
      if (gm.accumulate) v = __dadd_rn(dst[f], v);
      dst[f] = v;
    }
  }
}

// internal common descriptor view (pointer tables live inside the ABI descriptors)
struct LtView {
  const void *w, *in, *out;
};

// The ABI descriptors keep their pointer tables inline; the kernels need pointers to those tables.
// We build a device-side "view" array next to the descriptor copy.
struct MomentsDescDev {
  const double *w[KB200_MAX_DIRSETS];
  const double *in[KB200_MAX_DIRSETS];
  double *out[KB200_MAX_DIRSETS];
};
// kernels index desc.w / desc.in / desc.out as pointer tables:
struct MomentsDescK {
  const double *const *w;
  const double *const *in;
  double *const *out;
};

}  // namespace kb200

using namespace kb200;

int kb200_moments_mma_try(int mode, int layout, int M, int Ds, int Gs, int Zs, int nsets, int accumulate, const void *d_views,
                          int n, const void *const *h_ptrs, int n_ptrs, int same_w, cudaStream_t st);  // kb200_moments_mma.cu
int kb200_moments_rowmma_try(int mode, int layout, int M, int Ds, int Gs, int Zs, int nsets, int accumulate, const void *d_views,
                             int n, const void *const *h_ptrs, int n_ptrs, cudaStream_t st);  // kb200_moments_rowmma.cu

namespace {

struct TileChoice { int OT, OTP, ntiles; };
static TileChoice choose_tile(int nout) {
  int nt = (nout + 25) / 26;
  int ot = (nout + nt - 1) / nt;
  int otp = (ot + 1) & ~1;
  static const int allowed[] = {2, 4, 10, 12, 16, 18, 22, 24, 26};
  for (int a : allowed)
    if (a >= otp) { otp = a; break; }
  return {ot, otp, nt};
}

template <int OTP, bool EXACT>
static int launch_col(const MomentsDescK *d_descs, int n, const MomentsGeom &gm, int vec, int ntiles_total, cudaStream_t st) {
  int K = (gm.mode == 0) ? gm.nsets * gm.Ds : gm.M;
  size_t smem = (size_t)K * OTP * sizeof(double);
  long long threads = gm.B * (gm.N / vec);
  dim3 grid((unsigned)((threads + 127) / 128), ntiles_total, n);
  if (vec == 2) {
    auto k = moments_col_kernel<OTP, 2, EXACT, MomentsDescK>;
    KB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    k<<<grid, 128, smem, st>>>(d_descs, gm);
  } else {
    auto k = moments_col_kernel<OTP, 1, EXACT, MomentsDescK>;
    KB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    k<<<grid, 128, smem, st>>>(d_descs, gm);
  }
  return post_launch("moments_col");
}

template <int OTP, bool EXACT, int RT>
static int launch_row_t(const MomentsDescK *d_descs, int n, const MomentsGeom &gm, size_t smem, cudaStream_t st) {
  dim3 grid((unsigned)((gm.R + RT - 1) / RT), n, 1);
  auto k = moments_row_kernel<OTP, EXACT, MomentsDescK, RT>;
  KB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  k<<<grid, 128, smem, st>>>(d_descs, gm);
  return post_launch("moments_row");
}
template <int OTP, bool EXACT>
static int launch_row(const MomentsDescK *d_descs, int n, const MomentsGeom &gm, cudaStream_t st) {
  int Kc = (gm.mode == 0) ? gm.Ds : gm.M, Oc = (gm.mode == 0) ? gm.M : gm.Ds;
  // the largest row tile whose staging fits: 256 rows, or 128 / 64 for long rows (M = 100 with 32 directions per set)
  auto need = [&](int rt) { return ((size_t)rt * (Kc | 1) + (size_t)rt * (Oc | 1) + (size_t)Kc * OTP) * sizeof(double); };
  if (need(256) <= 220 * 1024) return launch_row_t<OTP, EXACT, 256>(d_descs, n, gm, need(256), st);
  if (need(128) <= 220 * 1024) return launch_row_t<OTP, EXACT, 128>(d_descs, n, gm, need(128), st);
  KB_REQUIRE(need(64) <= 220 * 1024, "moments_row: tile needs %zu bytes of shared memory", need(64));
  return launch_row_t<OTP, EXACT, 64>(d_descs, n, gm, need(64), st);
}


template <bool EXACT>
static int dispatch(bool row, int otp, const MomentsDescK *d, int n, const MomentsGeom &gm, int vec, int ntiles_total, cudaStream_t st) {
#define CASE(X) case X: return row ? launch_row<X, EXACT>(d, n, gm, st) : launch_col<X, EXACT>(d, n, gm, vec, ntiles_total, st);
  switch (otp) {
    CASE(2) CASE(4) CASE(10) CASE(12) CASE(16) CASE(18) CASE(22) CASE(24) CASE(26)
  }
#undef CASE
  set_error("moments: unsupported tile %d", otp);
  return 2;
}

// Builds the device pointer tables + geometry and launches. `mode` 0 = LTimes, 1 = LPlusTimes.
static int run_moments(int mode, int layout, int M, int Ds, int Gs, int Zs, int nsets, int accumulate,
                       const void *h_descs, size_t desc_stride, int n, size_t off_w, size_t off_in, size_t off_out,
                       bool in_is_table, bool out_is_table, cudaStream_t st) {
  // 1. device copy of the ABI descriptors (cached)
  const void *d_abi = nullptr;
  int rc = device_descs(h_descs, desc_stride * n, &d_abi, st);
  if (rc) return rc;
  // 2. kernel-side views pointing into the device copy
  std::vector<MomentsDescK> views(n);
  // a single pointer (phi / phi_out) is itself a 1-entry table stored inside the descriptor
  for (int i = 0; i < n; ++i) {
    const char *base = (const char *)d_abi + desc_stride * i;
    views[i].w = (const double *const *)(base + off_w);
    views[i].in = (const double *const *)(base + off_in);
    views[i].out = (double *const *)(base + off_out);
  }
  (void)in_is_table; (void)out_is_table;
  const void *d_views = nullptr;
  rc = device_descs(views.data(), sizeof(MomentsDescK) * n, &d_views, st);
  if (rc) return rc;

  if (!exact_mode()) {  // default arithmetic: fp64 tensor-core path for the (group,zone)-contiguous layouts
    std::vector<const void *> ptrs;
    for (int i = 0; i < n; ++i) {
      const char *hb = (const char *)h_descs + desc_stride * i;
      const void *const *tin = (const void *const *)(hb + off_in), *const *tout = (const void *const *)(hb + off_out);
      for (int s = 0; s < (mode == 0 ? nsets : 1); ++s) ptrs.push_back(tin[s]);
      for (int s = 0; s < (mode == 0 ? 1 : nsets); ++s) ptrs.push_back(tout[s]);
    }
    int same_w = 1;  // every descriptor of the call uses the same weight tables (true for the host layer: one ell per direction set)
    for (int i = 1; i < n && same_w; ++i) {
      const void *const *w0 = (const void *const *)((const char *)h_descs + off_w);
      const void *const *wi = (const void *const *)((const char *)h_descs + desc_stride * i + off_w);
      for (int s = 0; s < nsets; ++s) same_w = same_w && (w0[s] == wi[s]);
    }
    rc = kb200_moments_mma_try(mode, layout, M, Ds, Gs, Zs, nsets, accumulate, d_views, n, ptrs.data(), (int)ptrs.size(), same_w, st);
    if (rc >= 0) return rc;
    rc = kb200_moments_rowmma_try(mode, layout, M, Ds, Gs, Zs, nsets, accumulate, d_views, n, ptrs.data(), (int)ptrs.size(), st);
    if (rc >= 0) return rc;
  }

  MomentsGeom gm;
  memset(&gm, 0, sizeof(gm));
  gm.mode = mode; gm.M = M; gm.Ds = Ds; gm.nsets = nsets; gm.accumulate = accumulate;
  TileChoice tc = choose_tile(mode == 0 ? M : Ds);
  gm.OT = tc.OT; gm.ntiles_o = tc.ntiles;
  bool row = (layout == 3 || layout == 5);
  int vec = 1, ntiles_total = tc.ntiles * (mode == 0 ? 1 : nsets);
  if (row) {
    gm.R = (long long)Gs * Zs;
  } else {
    Strides3 fs = strides_dgz(layout, Ds, Gs, Zs), ms = strides_dgz(layout, M, Gs, Zs);
    switch (layout) {
      case 0: case 1: gm.B = 1; gm.N = (long long)Gs * Zs; gm.flux_b = 0; gm.mom_b = 0; break;        // DGZ, DZG
      case 2: gm.B = Gs; gm.N = Zs; gm.flux_b = fs.g; gm.mom_b = ms.g; break;                          // GDZ
      default: gm.B = Zs; gm.N = Gs; gm.flux_b = fs.z; gm.mom_b = ms.z; break;                         // ZDG
    }
    gm.flux_a = fs.a; gm.mom_a = ms.a;
    vec = (gm.N % 2 == 0) ? 2 : 1;
  }
  if (exact_mode()) return dispatch<true>(row, tc.OTP, (const MomentsDescK *)d_views, n, gm, vec, ntiles_total, st);
  return dispatch<false>(row, tc.OTP, (const MomentsDescK *)d_views, n, gm, vec, ntiles_total, st);
}

}  // namespace

extern "C" {

int kb200_ltimes(const kb200_ltimes_desc *h, int n, kb200_stream_t stream) {
  if (n <= 0) return 0;
  KB_REQUIRE(h, "kb200_ltimes: null descriptors");
  for (int i = 0; i < n; ++i) {
    KB_REQUIRE(h[i].layout == h[0].layout && h[i].M == h[0].M && h[i].Ds == h[0].Ds && h[i].Gs == h[0].Gs &&
                   h[i].Zs == h[0].Zs && h[i].nsets == h[0].nsets && h[i].accumulate == h[0].accumulate,
               "kb200_ltimes: all descriptors of one call must share their dimensions");
    KB_REQUIRE(h[i].nsets >= 1 && h[i].nsets <= KB200_MAX_DIRSETS, "kb200_ltimes: nsets=%d out of range", h[i].nsets);
    KB_REQUIRE(h[i].layout >= 0 && h[i].layout < 6, "kb200_ltimes: bad layout %d", h[i].layout);
    KB_REQUIRE(h[i].phi, "kb200_ltimes: null phi");
  }
  if (h[0].M <= 0 || h[0].Ds <= 0 || h[0].Gs <= 0 || h[0].Zs <= 0) return 0;
  return run_moments(0, h[0].layout, h[0].M, h[0].Ds, h[0].Gs, h[0].Zs, h[0].nsets, h[0].accumulate, h,
                     sizeof(kb200_ltimes_desc), n, offsetof(kb200_ltimes_desc, ell), offsetof(kb200_ltimes_desc, psi),
                     offsetof(kb200_ltimes_desc, phi), true, false, resolve_stream(stream));
}

int kb200_lplustimes(const kb200_lplustimes_desc *h, int n, kb200_stream_t stream) {
  if (n <= 0) return 0;
  KB_REQUIRE(h, "kb200_lplustimes: null descriptors");
  for (int i = 0; i < n; ++i) {
    KB_REQUIRE(h[i].layout == h[0].layout && h[i].M == h[0].M && h[i].Ds == h[0].Ds && h[i].Gs == h[0].Gs &&
                   h[i].Zs == h[0].Zs && h[i].nsets == h[0].nsets && h[i].accumulate == h[0].accumulate,
               "kb200_lplustimes: all descriptors of one call must share their dimensions");
    KB_REQUIRE(h[i].nsets >= 1 && h[i].nsets <= KB200_MAX_DIRSETS, "kb200_lplustimes: nsets=%d out of range", h[i].nsets);
    KB_REQUIRE(h[i].layout >= 0 && h[i].layout < 6, "kb200_lplustimes: bad layout %d", h[i].layout);
    KB_REQUIRE(h[i].phi_out, "kb200_lplustimes: null phi_out");
  }
  if (h[0].M <= 0 || h[0].Ds <= 0 || h[0].Gs <= 0 || h[0].Zs <= 0) return 0;
  return run_moments(1, h[0].layout, h[0].M, h[0].Ds, h[0].Gs, h[0].Zs, h[0].nsets, h[0].accumulate, h,
                     sizeof(kb200_lplustimes_desc), n, offsetof(kb200_lplustimes_desc, ell_plus),
                     offsetof(kb200_lplustimes_desc, phi_out), offsetof(kb200_lplustimes_desc, rhs), false, true,
                     resolve_stream(stream));
}

}  // extern "C"
