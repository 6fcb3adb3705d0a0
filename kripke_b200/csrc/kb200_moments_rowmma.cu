// LTimes / LPlusTimes on the fp64 tensor-core pipe for the storage orders whose direction / moment
// index is the FASTEST one (GZD, ZGD):  psi[r][d], phi[r][nm] with r = the flat (group, zone) row.
//
//   LTimes      phi[r][nm]  (+)= sum_q sum_d  psi_q[r][d]   * ell_q[d][nm]        (Kernel/LTimes.cpp:54-65)
//   LPlusTimes  rhs_q[r][d] (+)= sum_nm       phi_out[r][nm] * ell_plus_q[d][nm]  (Kernel/LPlusTimes.cpp:49-60)
//
// Here the streamed field is the natural A operand of mma.sync.m8n8k4.f64 (row = r, k = its
// contiguous index): a warp owns 8 rows (one m-tile), reads them straight from global memory in
// fragment order (each row's k-run is contiguous, so the 8 x 32-byte / 8 x 64-byte pieces of a
// request are whole sectors of rows that are fully consumed), keeps every output tile of the row
// block in registers, and writes each output row segment once.  Nothing is staged through shared
// memory except the weights, which are converted once per CTA into fragment-major B operands.
//   * streaming regime (LTimes): many input chunks (direction sets) accumulate into one output
//     chunk; the A fragments of the next chunk are loaded while the current one is multiplied.
//     With Ds % 8 == 0 a lane loads 16 bytes (two consecutive directions) and feeds two MMAs whose
//     k indices are permuted accordingly (the reduction order is free).
//   * resident regime (LPlusTimes): one input chunk, A fragments of the whole reduction length in
//     registers, loop over the output chunks (direction sets).
//   * O = 8q+1 (M = 25 in LTimes): the last output column is accumulated by DFMA on the A-fragment
//     layout; K = 4s+1 (M = 25 in LPlusTimes): the last reduction index is added by DFMA on the
//     C-fragment layout -- instead of a tensor-core tile that would be 7/8 (3/4) padding.
// Not bit-ordered like the reference (the tensor core sums four products per step); EXACT mode
// uses the DFMA kernels in kb200_moments.cu.
#include "kb200_common.cuh"

namespace kb200 {

struct MomentsDescK {  // same as in kb200_moments.cu: pointer tables inside the device copy of the ABI descriptor
  const double *const *w;
  const double *const *in;
  double *const *out;
};

struct RowGeom {
  int mode;        // 0 = LTimes, 1 = LPlusTimes
  int M, Ds, nsets, accumulate;
  int Kc, Oc;      // reduction length per input chunk, outputs per output chunk
  int S;           // k4 steps per input chunk covered by tensor-core tiles
  int nin, nout;   // input / output chunks
  long long R;     // rows
};

__device__ __forceinline__ void rm_dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double rm_ldg(const double *p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double2 rm_ldg2(const double *p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

// weight element W(ic, oc)[k][o] of the generic product out_oc[r][o] = sum_ic sum_k in_ic[r][k] * W[k][o]
__device__ __forceinline__ double rm_weight(const MomentsDescK &dsc, const RowGeom &gm, int ic, int oc, int k, int o) {
  if (k >= gm.Kc || o >= gm.Oc) return 0.0;
  if (gm.mode == 0) return dsc.w[ic][(size_t)k * gm.M + o];   // ell_q[d][nm]:      k = d,  o = nm
  return dsc.w[oc][(size_t)o * gm.M + k];                     // ell_plus_q[d][nm]: k = nm, o = d
}

// reduction index served by lane-quad position kk in k4 step s (VEC: a lane holds two consecutive k)
template <bool VEC>
__device__ __forceinline__ int rm_kidx(int s, int kk) {
  return VEC ? 8 * (s >> 1) + 2 * kk + (s & 1) : 4 * s + kk;
}

// ------------------------------------------------------------------------------------------------
// streaming regime: out[r][o] = sum_ic sum_k in_ic[r][k] W_ic[k][o], all o-tiles resident (LTimes)
// ------------------------------------------------------------------------------------------------
template <int NT, bool XCOL, bool VEC, int SMAX>
__global__ void __launch_bounds__(256) moments_rowmma_stream_kernel(const MomentsDescK *__restrict__ descs, RowGeom gm) {
  extern __shared__ __align__(16) double rsm[];
  const MomentsDescK dsc = descs[blockIdx.y];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int S = gm.S, nin = gm.nin, Kc = gm.Kc, Oc = gm.Oc;
  double *Bs = rsm;                                  // [nin][S][NT][32]
  double *wxs = rsm + (size_t)nin * S * NT * 32;     // [nin][S][4]  weights of the extra output column
  for (int idx = threadIdx.x; idx < nin * S * NT * 32; idx += blockDim.x) {
    const int l = idx & 31, f = idx >> 5, nt = f % NT, s = (f / NT) % S, ic = f / (NT * S);
    Bs[idx] = rm_weight(dsc, gm, ic, 0, rm_kidx<VEC>(s, l & 3), 8 * nt + (l >> 2));
  }
  if (XCOL)
    for (int idx = threadIdx.x; idx < nin * S * 4; idx += blockDim.x) {
      const int kk = idx & 3, s = (idx >> 2) % S, ic = idx / (4 * S);
      wxs[idx] = rm_weight(dsc, gm, ic, 0, rm_kidx<VEC>(s, kk), Oc - 1);
    }
  __syncthreads();

  const long long ntiles = (gm.R + 7) / 8;
  const long long wstride = (long long)gridDim.x * 8;
  const int kk = lane & 3;
  double *__restrict__ out = dsc.out[0];

  for (long long tile = (long long)blockIdx.x * 8 + warp; tile < ntiles; tile += wstride) {
    const long long r = tile * 8 + (lane >> 2);
    const bool rv = r < gm.R;
    const long long rl = rv ? r : gm.R - 1;
    double acc[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
    double px = 0.0;
    double an[SMAX];  // A fragments of the chunk being loaded
    auto load = [&](int ic) {
      const double *row = dsc.in[ic] + rl * Kc;
      if (VEC) {
#pragma unroll
        for (int s = 0; s < SMAX; s += 2)
          if (s < S) { const double2 v = rm_ldg2(row + 4 * s + 2 * kk); an[s] = v.x; an[s + 1] = v.y; }
      } else {
#pragma unroll
        for (int s = 0; s < SMAX; ++s)
          if (s < S) { const int k = 4 * s + kk; an[s] = (k < Kc) ? rm_ldg(row + k) : 0.0; }
      }
    };
    load(0);
    for (int ic = 0; ic < nin; ++ic) {
      double ac[SMAX];
#pragma unroll
      for (int s = 0; s < SMAX; ++s) ac[s] = an[s];
      if (ic + 1 < nin) load(ic + 1);
      const double *bf = Bs + (size_t)ic * S * NT * 32 + lane;
      const double *wx = wxs + (size_t)ic * S * 4 + kk;
#pragma unroll
      for (int s = 0; s < SMAX; ++s)
        if (s < S) {
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) rm_dmma(acc[nt][0], acc[nt][1], ac[s], bf[(s * NT + nt) * 32]);
          if (XCOL) px = fma(ac[s], wx[s * 4], px);
        }
    }
    // epilogue: C fragment = rows lane>>2, columns 8*nt + 2*kk + {0,1}
    double *orow = out + r * Oc;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int o = 8 * nt + 2 * kk;
      if (rv) {
        if (o < Oc) { double v = acc[nt][0]; if (gm.accumulate) v += orow[o]; orow[o] = v; }
        if (o + 1 < Oc && !(XCOL && o + 1 == Oc - 1)) { double v = acc[nt][1]; if (gm.accumulate) v += orow[o + 1]; orow[o + 1] = v; }
      }
    }
    if (XCOL) {
      px += __shfl_xor_sync(0xffffffffu, px, 1);
      px += __shfl_xor_sync(0xffffffffu, px, 2);
      if (rv && kk == 0) { double v = px; if (gm.accumulate) v += orow[Oc - 1]; orow[Oc - 1] = v; }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// resident regime: out_oc[r][o] = sum_k in[r][k] W_oc[k][o], A fragments of the whole K in registers
// ------------------------------------------------------------------------------------------------
template <int NT, int KR, bool XK, bool VEC>
__global__ void __launch_bounds__(256) moments_rowmma_resident_kernel(const MomentsDescK *__restrict__ descs, RowGeom gm) {
  extern __shared__ __align__(16) double rsm[];
  const MomentsDescK dsc = descs[blockIdx.y];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int S = gm.S, nout = gm.nout, Kc = gm.Kc, Oc = gm.Oc;
  double *Bs = rsm;                                   // [nout][S][NT][32]
  double *wks = rsm + (size_t)nout * S * NT * 32;     // [nout][NT*8]  weights of the extra reduction index
  // with XK the tensor-core tiles cover k < 4*S = Kc - 1 and the last index goes through DFMA
  for (int idx = threadIdx.x; idx < nout * S * NT * 32; idx += blockDim.x) {
    const int l = idx & 31, f = idx >> 5, nt = f % NT, s = (f / NT) % S, oc = f / (NT * S);
    const int k = rm_kidx<VEC>(s, l & 3);
    Bs[idx] = (XK && k >= Kc - 1) ? 0.0 : rm_weight(dsc, gm, 0, oc, k, 8 * nt + (l >> 2));
  }
  if (XK)
    for (int idx = threadIdx.x; idx < nout * NT * 8; idx += blockDim.x) {
      const int o = idx % (NT * 8), oc = idx / (NT * 8);
      wks[idx] = rm_weight(dsc, gm, 0, oc, Kc - 1, o);
    }
  __syncthreads();

  const long long ntiles = (gm.R + 7) / 8;
  const long long wstride = (long long)gridDim.x * 8;
  const int kk = lane & 3;
  const double *__restrict__ in = dsc.in[0];

  for (long long tile = (long long)blockIdx.x * 8 + warp; tile < ntiles; tile += wstride) {
    const long long r = tile * 8 + (lane >> 2);
    const bool rv = r < gm.R;
    const long long rl = rv ? r : gm.R - 1;
    const double *row = in + rl * Kc;
    double a[KR];
    if (VEC) {
#pragma unroll
      for (int s = 0; s < KR; s += 2)
        if (s < S) { const double2 v = rm_ldg2(row + 4 * s + 2 * kk); a[s] = v.x; a[s + 1] = v.y; }
    } else {
#pragma unroll
      for (int s = 0; s < KR; ++s)
        if (s < S) { const int k = 4 * s + kk; a[s] = (k < (XK ? Kc - 1 : Kc)) ? rm_ldg(row + k) : 0.0; }
    }
    const double ax = XK ? rm_ldg(row + Kc - 1) : 0.0;
    for (int oc = 0; oc < nout; ++oc) {
      double acc[NT][2];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
      const double *bf = Bs + (size_t)oc * S * NT * 32 + lane;
#pragma unroll
      for (int s = 0; s < KR; ++s)
        if (s < S) {
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) rm_dmma(acc[nt][0], acc[nt][1], a[s], bf[(s * NT + nt) * 32]);
        }
      double *orow = dsc.out[oc] + r * Oc;
      const double *wk = wks + (size_t)oc * NT * 8 + 2 * kk;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int o = 8 * nt + 2 * kk;
        double v0 = acc[nt][0], v1 = acc[nt][1];
        if (XK) {
          const double2 w2 = *reinterpret_cast<const double2 *>(wk + 8 * nt);
          v0 = fma(ax, w2.x, v0);
          v1 = fma(ax, w2.y, v1);
        }
        if (rv && o < Oc) {
          if ((Oc & 1) == 0) {  // even row length: the pair is 16-byte aligned and inside the row
            double2 *p = reinterpret_cast<double2 *>(orow + o);
            if (gm.accumulate) { const double2 old = *p; v0 += old.x; v1 += old.y; }
            *p = make_double2(v0, v1);
          } else {
            if (gm.accumulate) v0 += orow[o];
            orow[o] = v0;
            if (o + 1 < Oc) { if (gm.accumulate) v1 += orow[o + 1]; orow[o + 1] = v1; }
          }
        }
      }
    }
  }
}

}  // namespace kb200

using namespace kb200;

template <int NT, bool XCOL, bool VEC>
static int launch_stream(const MomentsDescK *dv, int n, const RowGeom &gm, cudaStream_t st) {
  const size_t smem = ((size_t)gm.nin * gm.S * NT * 32 + (size_t)gm.nin * gm.S * 4) * sizeof(double);
  if (smem > 200 * 1024) return -1;
  auto k = moments_rowmma_stream_kernel<NT, XCOL, VEC, 8>;
  KB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm > 4) per_sm = 4;
  if (per_sm < 1) per_sm = 1;
  long long ctas = (long long)sm_count() * per_sm / n;
  const long long need = ((gm.R + 7) / 8 + 7) / 8;
  if (ctas > need) ctas = need;
  if (ctas < 1) ctas = 1;
  k<<<dim3((unsigned)ctas, n, 1), 256, smem, st>>>(dv, gm);
  return post_launch("moments_rowmma_stream");
}

template <int NT, int KR, bool XK, bool VEC>
static int launch_resident(const MomentsDescK *dv, int n, const RowGeom &gm, cudaStream_t st) {
  const size_t smem = ((size_t)gm.nout * gm.S * NT * 32 + (size_t)gm.nout * NT * 8) * sizeof(double);
  if (smem > 200 * 1024) return -1;
  auto k = moments_rowmma_resident_kernel<NT, KR, XK, VEC>;
  KB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm > 4) per_sm = 4;
  if (per_sm < 1) per_sm = 1;
  long long ctas = (long long)sm_count() * per_sm / n;
  const long long need = ((gm.R + 7) / 8 + 7) / 8;
  if (ctas > need) ctas = need;
  if (ctas < 1) ctas = 1;
  k<<<dim3((unsigned)ctas, n, 1), 256, smem, st>>>(dv, gm);
  return post_launch("moments_rowmma_resident");
}

// Returns 0 if handled, -1 if this path does not apply (caller falls back to the DFMA kernels), >0 on error.
int kb200_moments_rowmma_try(int mode, int layout, int M, int Ds, int Gs, int Zs, int nsets, int accumulate, const void *d_views,
                             int n, const void *const *h_ptrs, int n_ptrs, cudaStream_t st) {
  if (layout != 3 && layout != 5) return -1;
  const char *env = getenv("KB200_MOMENTS_DFMA");
  if (env && env[0] == '1') return -1;
  RowGeom gm;
  memset(&gm, 0, sizeof(gm));
  gm.mode = mode; gm.M = M; gm.Ds = Ds; gm.nsets = nsets; gm.accumulate = accumulate;
  gm.R = (long long)Gs * Zs;
  const MomentsDescK *dv = (const MomentsDescK *)d_views;
  bool al16 = true;
  for (int i = 0; i < n_ptrs; ++i)
    if (((uintptr_t)h_ptrs[i] & 15) != 0) al16 = false;
  if (mode == 0) {  // LTimes: K = Ds per direction set (streamed), O = M
    gm.Kc = Ds; gm.Oc = M; gm.nin = nsets; gm.nout = 1;
    const bool vec = al16 && (Ds % 8 == 0);
    gm.S = vec ? Ds / 4 : (Ds + 3) / 4;
    if (gm.S > 8) return -1;
    const bool xcol = (M % 8 == 1) && M > 8;
    const int nt = xcol ? M / 8 : (M + 7) / 8;
#define STREAM(NT_)                                                                          \
  if (nt <= NT_) {                                                                           \
    if (xcol) return vec ? launch_stream<NT_, true, true>(dv, n, gm, st) : launch_stream<NT_, true, false>(dv, n, gm, st); \
    return vec ? launch_stream<NT_, false, true>(dv, n, gm, st) : launch_stream<NT_, false, false>(dv, n, gm, st);         \
  }
    STREAM(1) STREAM(2) STREAM(3) STREAM(5) STREAM(8) STREAM(13)
#undef STREAM
    return -1;
  }
  // LPlusTimes: K = M (resident), O = Ds per direction set
  gm.Kc = M; gm.Oc = Ds; gm.nin = 1; gm.nout = nsets;
  const bool vec = al16 && (M % 8 == 0);
  const bool xk = !vec && (M % 4 == 1) && M > 4;
  gm.S = xk ? M / 4 : (M + 3) / 4;
  const int nt = (Ds + 7) / 8;
  if (nt > 4) return -1;
#define RESIDENT(NT_, KR_)                                                                   \
  if (nt <= NT_ && gm.S <= KR_) {                                                            \
    if (vec) return launch_resident<NT_, KR_, false, true>(dv, n, gm, st);                   \
    if (xk) return launch_resident<NT_, KR_, true, false>(dv, n, gm, st);                    \
    return launch_resident<NT_, KR_, false, false>(dv, n, gm, st);                           \
  }
  RESIDENT(1, 2) RESIDENT(1, 6) RESIDENT(1, 26)
  RESIDENT(2, 2) RESIDENT(2, 6) RESIDENT(2, 26)
  RESIDENT(3, 2) RESIDENT(3, 6) RESIDENT(3, 26)
  RESIDENT(4, 2) RESIDENT(4, 6) RESIDENT(4, 26)
#undef RESIDENT
  return -1;
}
