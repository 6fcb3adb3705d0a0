// SweepSubdomain for sm_100a: diamond-difference upwind sweep as a hyperplane wavefront.
//
// Reference: src/Kripke/Kernel/SweepSubdomain.cpp:34-109 (arithmetic), Arch/SweepSubdomains.h
// (the reference's CUDA policy: one block per (d,g), threads over (j,i), global-memory planes).
//
// Mapping used here (not the reference's):
//   * lanes of a warp  <->  32 consecutive phase-space elements e = (direction, group) of the
//     subdomain, enumerated in storage order, so that for the element-fastest layouts
//     (DZG, GZD, ZDG, ZGD) every rhs/psi/plane access of a warp is one coalesced 256-byte row;
//   * a warp owns LPW zone *lines* (fixed j,k, running along i) of a TJ x TK column tile and
//     marches them along i as a systolic array: at step h line (jj,kk) updates zone
//     i = h - jj - kk, i.e. all zones of the hyperplane i+jj+kk = h are updated concurrently;
//   * the i-face flux never leaves a register; the j- and k-face fluxes are handed to the
//     neighbouring lines through double-buffered shared memory (one __syncthreads per
//     hyperplane); only tile-boundary faces touch the global i/j/k_plane arrays, which are
//     updated in place exactly like the reference does (on entry incoming, on exit outgoing);
//   * column tiles of one subdomain form a 2-D wavefront (tile (tj,tk) needs (tj-1,tk) and
//     (tj,tk-1)); each tile diagonal of all subdomains/slices of a batch is one launch.
//
// Arithmetic order is the reference's; in EXACT mode (KB200_EXACT=1) multiplies and adds are
// kept separate so psi is bit-identical to the CPU reference, otherwise the three face products
// are contracted to DFMA (|difference| ~ 1e-16 relative).
#include "kb200_common.cuh"
#include <vector>

namespace kb200 {

constexpr int TJ = 8, TK = 8;      // column tile (lines)
constexpr int SWEEP_WARPS = 16;    // warps per block
constexpr int LPW = TJ * TK / SWEEP_WARPS;

struct SweepGeom {
  int layout, Ds, Gs, ni, nj, nk;
  int ntj, ntk, nslices, diag;
};

template <bool EXACT>
__device__ __forceinline__ double madd(double a, double b, double c) {
  if (EXACT) return __dadd_rn(__dmul_rn(a, b), c);
  return fma(a, b, c);
}

template <bool EXACT>
__global__ void __launch_bounds__(SWEEP_WARPS * 32) sweep_tile_kernel(const kb200_sweep_desc *__restrict__ descs, SweepGeom gm) {
  extern __shared__ __align__(16) double smem[];
  double *jf = smem;                              // [2][TK][TJ][32]
  double *kf = smem + 2 * TK * TJ * 32;           // [2][TK][TJ][32]

  const kb200_sweep_desc &ds = descs[blockIdx.z];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Ds = gm.Ds, Gs = gm.Gs, ni = gm.ni, nj = gm.nj, nk = gm.nk;

  // tile on this diagonal
  const int tj_lo = max(0, gm.diag - (gm.ntk - 1));
  const int tj = tj_lo + blockIdx.x;
  const int tk = gm.diag - tj;
  const int tjn = min(TJ, nj - tj * TJ), tkn = min(TK, nk - tk * TK);

  // phase-space element of this lane, enumerated with the faster of (d,g) first
  const Strides3 fs = strides_dgz(gm.layout, Ds, Gs, (long long)ni * nj * nk);
  const int E = Ds * Gs;
  int e = blockIdx.y * 32 + lane;
  const bool valid = e < E;
  if (!valid) e = E - 1;
  int d, g;
  if (fs.a < fs.g) { d = e % Ds; g = e / Ds; }
  else { g = e % Gs; d = e / Gs; }
  const StridesP ips = strides_plane(gm.layout, Ds, Gs, nj, nk);
  const StridesP jps = strides_plane(gm.layout, Ds, Gs, ni, nk);
  const StridesP kps = strides_plane(gm.layout, Ds, Gs, ni, nj);
  long long sg_sigt, sz_sigt;
  strides_gz(gm.layout, Gs, (long long)ni * nj * nk, sg_sigt, sz_sigt);

  const double *__restrict__ rhs = ds.rhs + (long long)d * fs.a + (long long)g * fs.g;
  double *__restrict__ psi = ds.psi + (long long)d * fs.a + (long long)g * fs.g;
  const double *__restrict__ sigt = ds.sigt + (long long)g * sg_sigt;
  double *__restrict__ ipl = ds.i_plane + (long long)d * ips.d + (long long)g * ips.g;
  double *__restrict__ jpl = ds.j_plane + (long long)d * jps.d + (long long)g * jps.g;
  double *__restrict__ kpl = ds.k_plane + (long long)d * kps.d + (long long)g * kps.g;
  double *__restrict__ iout = ds.out_plane[0] ? ds.out_plane[0] + (long long)d * ips.d + (long long)g * ips.g : nullptr;
  double *__restrict__ jout = (ds.out_plane[1] && tj == gm.ntj - 1) ? ds.out_plane[1] + (long long)d * jps.d + (long long)g * jps.g : nullptr;
  double *__restrict__ kout = (ds.out_plane[2] && tk == gm.ntk - 1) ? ds.out_plane[2] + (long long)d * kps.d + (long long)g * kps.g : nullptr;
  const bool i_zero = ds.inflow_zero[0] != 0;
  const bool j_zero = ds.inflow_zero[1] != 0 && tj == 0;
  const bool k_zero = ds.inflow_zero[2] != 0 && tk == 0;

  const double two_xcos = 2.0 * ds.xcos[d], two_ycos = 2.0 * ds.ycos[d], two_zcos = 2.0 * ds.zcos[d];
  const int id = ds.id, jd = ds.jd, kd = ds.kd;

  // per-line state
  double ifr[LPW], cy[LPW], cz[LPW];
  int jz[LPW], kz[LPW];  // actual zone indices of the line
#pragma unroll
  for (int m = 0; m < LPW; ++m) {
    const int l = warp + SWEEP_WARPS * m, jj = l % TJ, kk = l / TJ;
    const int jl = min(tj * TJ + jj, nj - 1), kl = min(tk * TK + kk, nk - 1);
    jz[m] = (jd > 0) ? jl : nj - 1 - jl;
    kz[m] = (kd > 0) ? kl : nk - 1 - kl;
    cy[m] = two_ycos / ds.dy[jz[m]];
    cz[m] = two_zcos / ds.dz[kz[m]];
    ifr[m] = 0.0;
  }
  const double cx0 = two_xcos / ds.dx[0];
  // uniform mesh in x (what the reference always generates): hoist the per-zone divide
  int uni = 1;
  for (int i = threadIdx.x; i < ni; i += blockDim.x) uni &= (ds.dx[i] == ds.dx[0]);
  const bool uniform_mesh = __syncthreads_and(uni) != 0;

  const int nsteps = ni + (tjn - 1) + (tkn - 1);
  for (int h = 0; h < nsteps; ++h) {
    const int buf = h & 1;
    double *jf_w = jf + buf * (TK * TJ * 32), *kf_w = kf + buf * (TK * TJ * 32);
    const double *jf_r = jf + (buf ^ 1) * (TK * TJ * 32), *kf_r = kf + (buf ^ 1) * (TK * TJ * 32);
#pragma unroll
    for (int m = 0; m < LPW; ++m) {
      const int l = warp + SWEEP_WARPS * m, jj = l % TJ, kk = l / TJ;
      const int ii = h - jj - kk;
      if (jj < tjn && kk < tkn && ii >= 0 && ii < ni) {
        const int iz = (id > 0) ? ii : ni - 1 - ii;
        const long long z = ((long long)kz[m] * nj + jz[m]) * ni + iz;
        const double r = __ldg(rhs + z * fs.z);
        const double st = __ldg(sigt + z * sz_sigt);
        const double cx = uniform_mesh ? cx0 : two_xcos / ds.dx[iz];
        // incoming faces
        double fi, fj, fk;
        if (ii == 0) fi = i_zero ? 0.0 : ipl[(long long)jz[m] * ips.a + (long long)kz[m] * ips.b];
        else fi = ifr[m];
        const long long jidx = (long long)iz * jps.a + (long long)kz[m] * jps.b;
        const long long kidx = (long long)iz * kps.a + (long long)jz[m] * kps.b;
        if (jj == 0) fj = j_zero ? 0.0 : jpl[jidx];
        else fj = jf_r[(kk * TJ + jj - 1) * 32 + lane];
        if (kk == 0) fk = k_zero ? 0.0 : kpl[kidx];
        else fk = kf_r[((kk - 1) * TJ + jj) * 32 + lane];
        // psi = (rhs + fi*cx + fj*cy + fk*cz) / (cx + cy + cz + sigt)   (SweepSubdomain.cpp:95-99)
        double num = madd<EXACT>(fi, cx, r);
        num = madd<EXACT>(fj, cy[m], num);
        num = madd<EXACT>(fk, cz[m], num);
        const double den = __dadd_rn(__dadd_rn(__dadd_rn(cx, cy[m]), cz[m]), st);
        const double p = __ddiv_rn(num, den);
        if (valid) psi[z * fs.z] = p;
        // diamond difference: outgoing = 2*psi - incoming (2*psi is exact, so fma == mul,sub)
        const double p2 = 2.0 * p;
        const double oi = p2 - fi, oj = p2 - fj, ok = p2 - fk;
        ifr[m] = oi;
        if (ii == ni - 1 && valid) {
          const long long iidx = (long long)jz[m] * ips.a + (long long)kz[m] * ips.b;
          ipl[iidx] = oi;
          if (iout) iout[iidx] = oi;
        }
        if (jj == tjn - 1) {
          if (valid) { jpl[jidx] = oj; if (jout) jout[jidx] = oj; }
        } else {
          jf_w[(kk * TJ + jj) * 32 + lane] = oj;
        }
        if (kk == tkn - 1) {
          if (valid) { kpl[kidx] = ok; if (kout) kout[kidx] = ok; }
        } else {
          kf_w[(kk * TJ + jj) * 32 + lane] = ok;
        }
      }
    }
    __syncthreads();
  }
}


}  // namespace kb200

using namespace kb200;

int kb200_sweep_zline_try(const kb200_sweep_desc *h, int n, const void *d_descs, cudaStream_t st);  // kb200_sweep_zline.cu
int kb200_sweep_elem_try(const kb200_sweep_desc *h, int n, const void *d_descs, cudaStream_t st);   // kb200_sweep_elem.cu
int kb200_sweep_irow_try(const kb200_sweep_desc *h, int n, const void *d_descs, const double *const *d_pop_w,
                         const double *const *d_pop_vol, const double *h_pop_vol_const, const double *d_pop_vol_const,
                         double *d_pop_partial, int pop_capacity, int *pop_count, cudaStream_t st);  // kb200_sweep_irow.cu
int kb200_sweep_pencil_try(const kb200_sweep_desc *h, int n, const void *d_descs, const double *const *d_pop_w,
                           const double *const *d_pop_vol, const double *d_pop_vol_const, double *d_pop_partial, int pop_capacity,
                           int *pop_count, cudaStream_t st);  // kb200_sweep_pencil.cu

// which kernel family handled the most recent kb200_sweep / kb200_sweep_population call of this process
static const char *g_last_sweep_kernel = "none";
extern "C" const char *kb200_last_sweep_kernel(void) { return g_last_sweep_kernel; }

static int sweep_impl(const kb200_sweep_desc *h, int n, const double *const *h_w, const double *const *h_volume,
                      const double *h_vol_const, double *d_partials, int capacity, int *count, kb200_stream_t stream) {
  if (count) *count = 0;
  if (n <= 0) return 0;
  KB_REQUIRE(h, "kb200_sweep: null descriptors");
  for (int i = 0; i < n; ++i) {
    KB_REQUIRE(h[i].layout == h[0].layout && h[i].Ds == h[0].Ds && h[i].Gs == h[0].Gs && h[i].ni == h[0].ni &&
                   h[i].nj == h[0].nj && h[i].nk == h[0].nk,
               "kb200_sweep: all subdomains of one call must share layout and extents");
    KB_REQUIRE(h[i].layout >= 0 && h[i].layout < 6, "kb200_sweep: bad layout %d", h[i].layout);
    KB_REQUIRE((h[i].id == 1 || h[i].id == -1) && (h[i].jd == 1 || h[i].jd == -1) && (h[i].kd == 1 || h[i].kd == -1),
               "kb200_sweep: sweep directions must be +1/-1");
    KB_REQUIRE(h[i].rhs && h[i].psi && h[i].sigt && h[i].i_plane && h[i].j_plane && h[i].k_plane && h[i].xcos &&
                   h[i].ycos && h[i].zcos && h[i].dx && h[i].dy && h[i].dz, "kb200_sweep: null pointer in descriptor %d", i);
  }
  if (h[0].Ds <= 0 || h[0].Gs <= 0 || h[0].ni <= 0 || h[0].nj <= 0 || h[0].nk <= 0) return 0;
  cudaStream_t st = resolve_stream(stream);
  const void *d = nullptr;
  int rc = device_descs(h, sizeof(*h) * n, &d, st);
  if (rc) return rc;
  {  // fast paths, optionally with fused population partials
    const void *d_w = nullptr, *d_v = nullptr, *d_vc = nullptr;
    if (d_partials && h_w && h_volume && capacity > 0) {
      rc = device_descs(h_w, sizeof(double *) * n, &d_w, st);
      if (rc) return rc;
      rc = device_descs(h_volume, sizeof(double *) * n, &d_v, st);
      if (rc) return rc;
      if (h_vol_const) {
        rc = device_descs(h_vol_const, sizeof(double) * n, &d_vc, st);
        if (rc) return rc;
      }
    }
    // element-fastest layouts: pencils of zone lines, lanes over elements
    rc = kb200_sweep_pencil_try(h, n, d, (const double *const *)d_w, (const double *const *)d_v, (const double *)d_vc,
                                d_w ? d_partials : nullptr, capacity, count, st);
    if (rc >= 0) { g_last_sweep_kernel = (count && *count > 0) ? "pencil+population" : "pencil"; return rc; }
    // zone-fastest layouts, ni = 4 * 2^k, default arithmetic
    rc = kb200_sweep_irow_try(h, n, d, (const double *const *)d_w, (const double *const *)d_v, d_vc ? h_vol_const : nullptr,
                              (const double *)d_vc, d_w ? d_partials : nullptr, capacity, count, st);
    if (rc >= 0) { g_last_sweep_kernel = (count && *count > 0) ? "irow+population" : "irow"; return rc; }
  }
  rc = kb200_sweep_zline_try(h, n, d, st);  // zone-fastest layouts with ni % 4 == 0
  if (rc >= 0) { g_last_sweep_kernel = "zline"; return rc; }
  rc = kb200_sweep_elem_try(h, n, d, st);  // element-fastest layouts
  if (rc >= 0) { g_last_sweep_kernel = "elem"; return rc; }
  g_last_sweep_kernel = "tile";
  SweepGeom gm;
  gm.layout = h[0].layout; gm.Ds = h[0].Ds; gm.Gs = h[0].Gs; gm.ni = h[0].ni; gm.nj = h[0].nj; gm.nk = h[0].nk;
  gm.ntj = (gm.nj + TJ - 1) / TJ; gm.ntk = (gm.nk + TK - 1) / TK;
  gm.nslices = (gm.Ds * gm.Gs + 31) / 32;
  const size_t smem = (size_t)2 * 2 * TK * TJ * 32 * sizeof(double);
  auto kern = exact_mode() ? sweep_tile_kernel<true> : sweep_tile_kernel<false>;
  KB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int diag = 0; diag < gm.ntj + gm.ntk - 1; ++diag) {
    gm.diag = diag;
    int tj_lo = diag - (gm.ntk - 1) > 0 ? diag - (gm.ntk - 1) : 0;
    int tj_hi = diag < gm.ntj - 1 ? diag : gm.ntj - 1;
    dim3 grid(tj_hi - tj_lo + 1, gm.nslices, n);
    kern<<<grid, SWEEP_WARPS * 32, smem, st>>>((const kb200_sweep_desc *)d, gm);
    rc = post_launch("sweep_tile");
    if (rc) return rc;
  }
  return 0;
}

extern "C" int kb200_sweep(const kb200_sweep_desc *h, int n, kb200_stream_t stream) {
  return sweep_impl(h, n, nullptr, nullptr, nullptr, nullptr, 0, nullptr, stream);
}

extern "C" int kb200_sweep_population(const kb200_sweep_desc *h, int n, const double *const *h_w, const double *const *h_volume,
                                      double *d_partials, int capacity, int *count, kb200_stream_t stream) {
  KB_REQUIRE(count, "kb200_sweep_population: null count");
  return sweep_impl(h, n, h_w, h_volume, nullptr, d_partials, capacity, count, stream);
}

extern "C" int kb200_sweep_population_uniform(const kb200_sweep_desc *h, int n, const double *const *h_w, const double *const *h_volume,
                                              const double *h_vol_const, double *d_partials, int capacity, int *count,
                                              kb200_stream_t stream) {
  KB_REQUIRE(count, "kb200_sweep_population_uniform: null count");
  return sweep_impl(h, n, h_w, h_volume, h_vol_const, d_partials, capacity, count, stream);
}
