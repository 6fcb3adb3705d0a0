// Scattering on the fp64 tensor-core pipe (mma.sync m8n8k4 f64) for the zone-fastest storage orders
// (DGZ, GDZ).  Reference: src/Kripke/Kernel/Scattering.cpp:73-99.
//
//   phi_out(nm,g,z) (+)= sum_{src set s} sum_{gp} sigs_z(n(nm), g, gp, z) * phi_s(nm,gp,z)
//   sigs_z = sum_{mix in zone z} sigs(mat(mix), n, g, gp) * fraction(mix)
//
// For a fixed moment nm this is a [Gs x G] * [G x Zs] product whose matrix depends on the zone only
// through the material mix.  Writing f_m(z) for the volume fraction of material m in zone z (0 if
// absent) gives   phi_out(:,z) = sum_m S_{m,n} * (f_m(z) * phi(:,z)),   i.e. one tensor-core
// product per material PRESENT in the 8-zone column block of a warp, with the B fragment scaled by
// the per-zone fraction.  Pure single-material blocks -- all but the material interfaces of the
// Kobayashi geometry -- take exactly one product with f = 1.0.  At BASELINE config 2/3 the kernel
// is fp64-bound (G/8 = 8..16 flop/byte), so the contraction runs on the DMMA pipe; the moments are
// the batch dimension and the three S_{m,n} matrices of the current Legendre order live in shared
// memory in fragment-major form (rebuilt when n(nm) changes, which happens L times per CTA).
// The streamed operand uses the same 3-stage cp.async slab pipeline as kb200_moments_mma.cu.
// EXACT mode (bit-ordered like the reference) uses the DFMA kernel in kb200_scatter_pop.cu.
#include "kb200_common.cuh"

namespace kb200 {

struct ScatGeom {
  int layout, sigs_layout, M, L1, G, Gs, Zs, nsrc, accumulate;  // sigs may keep another nesting's order (transform path)
  int O, K, nkc4;          // outputs of one o-chunk launch slice (<= 32 per CTA), reduction length, k-chunks
  int KC, nst;             // slab rows, slabs per tile
  long long in_b, in_r;    // moment (batch) stride, group (row) stride of phi / phi_out
  int ntn;                 // zone tiles
  double source;           // != 0: Kernel::source folded in -- moment 0 also receives source * (volume fraction of material 0)
};

__device__ __forceinline__ void sc_cp_async16_zfill(void *smem_dst, const void *gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void sc_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void sc_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void sc_dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Volume fraction of each of the three materials in every zone of every descriptor: frac[(desc*3 + m)*Zs + z].
// The mixture tables are walked ONCE per call here (three dependent loads per zone) instead of once per (moment,
// zone tile) inside the contraction, where that latency chain was as long as the tile's tensor-core work.
__global__ void scatter_fractions_kernel(const kb200_scattering_desc *__restrict__ descs, int Zs, double *__restrict__ frac) {
  const kb200_scattering_desc &ds = descs[blockIdx.y];
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

static double *g_frac_scratch = nullptr;
static size_t g_frac_doubles = 0;

constexpr int SC_STAGES = 3;
constexpr int SC_NB = 2, SC_NT = 64 * SC_NB, SC_NTP = SC_NT + 4;  // QP (o-tiles of 8 destination groups per CTA) is a template parameter

template <int QP>
__global__ void __launch_bounds__(256, 2) scatter_mma_kernel(const kb200_scattering_desc *__restrict__ descs, ScatGeom gm,
                                                              const double *__restrict__ fractions) {
  extern __shared__ __align__(16) double ssm[];
  constexpr int NB = SC_NB, NT = SC_NT, NTP = SC_NTP;
  constexpr int PPR = NT / 2, RPP = 256 / PPR;
  const kb200_scattering_desc &ds = descs[blockIdx.y];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Gs = gm.Gs, K = gm.K, KC = gm.KC, nst = gm.nst, ntn = gm.ntn;
  const int o_base = blockIdx.z * 8 * QP;                  // first destination group of this o-chunk
  const int O = min(8 * QP, gm.O - o_base);                // valid destination groups in the chunk
  double *Ws = ssm;                                        // [3 materials][nkc4][QP][32] fragment-major
  double *slab = ssm + (size_t)3 * gm.nkc4 * QP * 32;      // [SC_STAGES][KC][NTP]
  const double **inrow = reinterpret_cast<const double **>(slab + (size_t)SC_STAGES * gm.KC * NTP);  // [nst*KC] row bases
  for (int k = threadIdx.x; k < gm.nst * gm.KC; k += 256) {
    const double *p = nullptr;
    if (k < gm.K) { const int sset = k / gm.Gs, gq = k - sset * gm.Gs; p = ds.phi_src[sset] + (long long)gq * gm.in_r; }
    inrow[k] = p;
  }
  __syncthreads();
  const Strides4 ss = strides_sigs(gm.sigs_layout, gm.L1, gm.G);

  const long long ntiles = (long long)gm.M * ntn;          // tile = (moment, zone tile), moment-major
  const int my_tiles = ((long long)blockIdx.x < ntiles) ? (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
  const int nitems = my_tiles * nst;

  // producer cursor
  int i_left = nitems, i_st = 0, i_tn = (int)(blockIdx.x % ntn), i_b = (int)(blockIdx.x / ntn), i_buf = 0;
  // every warp copies and consumes only its own 8*NB columns of a slab: no CTA-wide barrier in the pipeline
  constexpr int PW = 4 * NB, RPW = 32 / PW;
  const int c2 = lane % PW, r0 = lane / PW;
  const int wcol0 = warp * 8 * NB;
  auto issue = [&]() {
    if (i_left > 0) {
      double *dst = slab + (size_t)i_buf * KC * NTP + wcol0 + 2 * c2 + (size_t)r0 * NTP;
      const int n = i_tn * NT + wcol0 + 2 * c2;
      const bool ncol = n < gm.Zs;
      const long long boff = (long long)i_b * gm.in_b + n;
      const double *const *rp = inrow + i_st * KC + r0;
#pragma unroll 4
      for (int r = r0; r < KC; r += RPW, rp += RPW, dst += (size_t)RPW * NTP) {
        const double *base = *rp;
        const bool valid = ncol && base != nullptr;
        sc_cp_async16_zfill(dst, valid ? base + boff : ds.phi_out, valid);
      }
      --i_left;
      if (++i_st == nst) {
        i_st = 0;
        i_tn += gridDim.x;
        while (i_tn >= ntn) { i_tn -= ntn; ++i_b; }
      }
    }
    sc_cp_async_commit();
    if (++i_buf == SC_STAGES) i_buf = 0;
  };
#pragma unroll
  for (int s = 0; s < SC_STAGES - 1; ++s) issue();

  double acc[QP][NB][2];
  double frac[3][NB];
  unsigned present = 0;  // bit (3*nb + m): material m occurs in column block nb of this warp
  const int ncol0 = warp * 8 * NB;
  int st = 0, tn = (int)(blockIdx.x % ntn), b = (int)(blockIdx.x / ntn), cbuf = 0;
  int n_cur = -1;

  for (int item = 0; item < nitems; ++item) {
    sc_cp_async_wait<SC_STAGES - 2>();
    __syncwarp();  // per-warp pipeline: only this warp's columns are consumed below
    issue();

    if (st == 0) {
      // (re)build the three material matrices of this Legendre order, fragment-major, zero padded
      const int n_leg = ds.moment_to_legendre[b];
      if (n_leg != n_cur) {  // CTA-uniform: every warp reaches this item with the same moment index
        if (n_cur >= 0) __syncthreads();  // nobody still reads the previous matrices
        n_cur = n_leg;
        for (int idx = threadIdx.x; idx < 3 * gm.nkc4 * QP * 32; idx += 256) {
          const int l = idx & 31, f = idx >> 5, ot = f % QP, kc = (f / QP) % gm.nkc4, mat = f / (QP * gm.nkc4);
          const int o = 8 * ot + (l >> 2), k = 4 * kc + (l & 3);
          double v = 0.0;
          if (o < O && k < K) {
            const int sset = k / Gs, gp = k - sset * Gs;
            v = ds.sigs[(long long)mat * ss.mat + (long long)n_leg * ss.n + (long long)(ds.glower_dst + o_base + o) * ss.g +
                        (long long)(ds.glower_src[sset] + gp) * ss.gp];
          }
          Ws[idx] = v;
        }
        __syncthreads();
      }
      // material fractions of this lane's B-fragment columns
      present = 0;
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        const int z = tn * NT + ncol0 + 8 * nb + (lane >> 2);
        frac[0][nb] = frac[1][nb] = frac[2][nb] = 0.0;
        if (z < gm.Zs) {
          const double *fz = fractions + (size_t)blockIdx.y * 3 * gm.Zs + z;
          frac[0][nb] = __ldg(fz); frac[1][nb] = __ldg(fz + gm.Zs); frac[2][nb] = __ldg(fz + 2 * (size_t)gm.Zs);
        }
#pragma unroll
        for (int m = 0; m < 3; ++m)
          if (__any_sync(0xffffffffu, frac[m][nb] != 0.0)) present |= 1u << (3 * nb + m);
      }
#pragma unroll
      for (int a = 0; a < QP; ++a)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) acc[a][nb][0] = acc[a][nb][1] = 0.0;
    }

    const double *buf = slab + (size_t)cbuf * KC * NTP;
    const int kc_lo = st * (KC / 4);
    const int nkc = min(gm.nkc4 - kc_lo, KC / 4);
    const double *brow0 = buf + (size_t)(lane & 3) * NTP + ncol0 + (lane >> 2);
    // one pass over the slab per material present in this warp's columns (one pass for pure blocks, which is
    // nearly all of them): no branches inside the k loop; absent columns carry a zero fraction
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      if (!((present >> m) & 0x9u)) continue;  // warp-uniform
      const double *wf = Ws + (((size_t)m * gm.nkc4 + kc_lo) * QP) * 32 + lane;
      const double *brow = brow0;
      double fm[NB];
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) fm[nb] = frac[m][nb];
#pragma unroll 4
      for (int kc = 0; kc < nkc; ++kc, brow += 4 * NTP, wf += QP * 32) {
        double af[QP], bm[NB];
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) bm[nb] = brow[8 * nb] * fm[nb];
#pragma unroll
        for (int a = 0; a < QP; ++a) af[a] = wf[a * 32];
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
          for (int a = 0; a < QP; ++a) sc_dmma884(acc[a][nb][0], acc[a][nb][1], af[a], bm[nb]);
      }
    }

    if (st == nst - 1) {
      const int ncol = tn * NT + ncol0 + 2 * (lane & 3);
      // Kernel::source folded in (Kernel/Source.cpp:59-75): phi_out(0,g,z) += strength * fraction of material 0, the same for
      // every group; zero for the other moments (x + 0.0 leaves x's bits alone)
      double sx[NB][2];
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        sx[nb][0] = sx[nb][1] = 0.0;
        const int n = ncol + 8 * nb;
        if (gm.source != 0.0 && b == 0 && n < gm.Zs) {
          const double *f0 = fractions + (size_t)blockIdx.y * 3 * gm.Zs + n;
          sx[nb][0] = gm.source * __ldg(f0);
          sx[nb][1] = gm.source * __ldg(f0 + 1);
        }
      }
#pragma unroll
      for (int a = 0; a < QP; ++a) {
        const int o = 8 * a + (lane >> 2);
        if (o < O) {
          double *row = ds.phi_out + ((long long)b * gm.in_b + (long long)(o_base + o) * gm.in_r);
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            const int n = ncol + 8 * nb;
            if (n < gm.Zs) {
              double2 v = make_double2(acc[a][nb][0] + sx[nb][0], acc[a][nb][1] + sx[nb][1]);
              double2 *p = reinterpret_cast<double2 *>(row + n);
              if (gm.accumulate) { const double2 old = *p; v.x += old.x; v.y += old.y; }
              *p = v;
            }
          }
        }
      }
    }
    if (++cbuf == SC_STAGES) cbuf = 0;
    if (++st == nst) {
      st = 0;
      tn += gridDim.x;
      while (tn >= ntn) { tn -= ntn; ++b; }
    }
  }
  sc_cp_async_wait<0>();
}

}  // namespace kb200

using namespace kb200;

extern "C" const char *g_last_scattering_kernel;  // kb200_scatter_pop.cu
int kb200_scatter_slab_try(const kb200_scattering_desc *h, int n, const void *d_descs, int sigs_layout, double source, cudaStream_t st);  // kb200_scatter_slab.cu

// Returns 0 if handled, -1 if this path does not apply (caller falls back to the DFMA kernel), >0 on error.
int kb200_scatter_mma_try(const kb200_scattering_desc *h, int n, const void *d_descs, int sigs_layout, double source, cudaStream_t st) {
  const int layout = h[0].layout;
  if (layout != 0 && layout != 2) return -1;
  {  // the one-read kernel first (descriptor groups, whole-K stages); it declines what it does not cover
    const int rc = kb200_scatter_slab_try(h, n, d_descs, sigs_layout, source, st);
    if (rc != -1) { g_last_scattering_kernel = "slab"; return rc; }
  }
  const char *env = getenv("KB200_SCATTER_DFMA");
  if (env && env[0] == '1') return -1;
  if (h[0].Zs % 2 != 0) return -1;
  for (int i = 0; i < n; ++i) {
    if (h[i].nsrc != h[0].nsrc || h[i].accumulate != h[0].accumulate || h[i].G != h[0].G || h[i].L1 != h[0].L1) return -1;
    if (((uintptr_t)h[i].phi_out & 15) != 0) return -1;
    for (int s = 0; s < h[i].nsrc; ++s)
      if (((uintptr_t)h[i].phi_src[s] & 15) != 0) return -1;
  }
  ScatGeom gm;
  memset(&gm, 0, sizeof(gm));
  gm.layout = layout; gm.sigs_layout = sigs_layout >= 0 ? sigs_layout : layout; gm.M = h[0].M; gm.L1 = h[0].L1; gm.G = h[0].G; gm.Gs = h[0].Gs; gm.Zs = h[0].Zs;
  gm.nsrc = h[0].nsrc; gm.accumulate = h[0].accumulate;
  gm.source = source;
  gm.O = gm.Gs; gm.K = gm.nsrc * gm.Gs; gm.nkc4 = (gm.K + 3) / 4;
  const int Kp = gm.nkc4 * 4;
  gm.KC = Kp < 16 ? Kp : 16;
  gm.nst = (Kp + gm.KC - 1) / gm.KC;
  const Strides3 ms = strides_dgz(layout, gm.M, gm.Gs, gm.Zs);
  gm.in_b = ms.a; gm.in_r = ms.g;
  gm.ntn = (gm.Zs + SC_NT - 1) / SC_NT;
  const int QP = (gm.O <= 16) ? 2 : 4;  // narrow group sets (16 groups) would waste half of a 4-tile CTA on zero padding
  const size_t smem = ((size_t)3 * gm.nkc4 * QP * 32 + (size_t)SC_STAGES * gm.KC * SC_NTP + (size_t)gm.nst * gm.KC) * sizeof(double);
  if (smem > 200 * 1024) return -1;
  if (QP == 2) KB_CUDA(cudaFuncSetAttribute(scatter_mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else KB_CUDA(cudaFuncSetAttribute(scatter_mma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int nochunks = (gm.O + 8 * QP - 1) / (8 * QP);
  const long long ntiles = (long long)gm.M * gm.ntn;
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 2) per_sm = 2;
  long long ctas = (long long)sm_count() * per_sm / ((long long)n * nochunks);
  if (ctas < 1) ctas = 1;
  if (ctas > ntiles) ctas = ntiles;
  // per-zone material fractions of every descriptor (scratch kept between calls)
  const size_t need = (size_t)n * 3 * gm.Zs;
  if (g_frac_doubles < need) {
    if (g_frac_scratch) { KB_CUDA(cudaDeviceSynchronize()); cudaFree(g_frac_scratch); g_frac_scratch = nullptr; }
    KB_CUDA(cudaMalloc(&g_frac_scratch, need * sizeof(double)));
    g_frac_doubles = need;
  }
  scatter_fractions_kernel<<<dim3((gm.Zs + 255) / 256, n, 1), 256, 0, st>>>((const kb200_scattering_desc *)d_descs, gm.Zs, g_frac_scratch);
  {
    int rc = post_launch("scatter_fractions");
    if (rc) return rc;
  }
  dim3 grid((unsigned)ctas, n, nochunks);
  if (QP == 2) scatter_mma_kernel<2><<<grid, 256, smem, st>>>((const kb200_scattering_desc *)d_descs, gm, g_frac_scratch);
  else scatter_mma_kernel<4><<<grid, 256, smem, st>>>((const kb200_scattering_desc *)d_descs, gm, g_frac_scratch);
  g_last_scattering_kernel = "mma";
  return post_launch("scatter_mma");
}
