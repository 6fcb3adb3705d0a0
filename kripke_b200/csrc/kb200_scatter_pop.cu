// Scattering, Source and Population for sm_100a.
//
//   Scattering  phi_out(nm,g,z) (+)= sum_src sum_gp sigs_z(n(nm),g,gp,z) * phi_src(nm,gp,z)
//               sigs_z = sum_{mix in z} sigs(mat(mix), n, g+glower_dst, gp+glower_src) * fraction(mix)
//               (src/Kripke/Kernel/Scattering.cpp:73-99)
//   Source      phi_out(0,g,zone(mix)) += strength*fraction(mix) for material 0 (Kernel/Source.cpp:59-75)
//   Population  sum (w(d)*psi(d,g,z))*volume(z)   (Kernel/Population.cpp:49-63)
//
// Scattering is kept dense over (g,gp) like the reference.  A thread owns one
// (moment, zone) pair and a register tile of GT destination groups; the per-zone material mix
// (<= 3 entries) is held in registers so pure zones -- the overwhelming majority -- take one
// multiply per sigs entry, which makes sigs_z bit-identical to the reference's 0.0 + sigs*1.0.
#include "kb200_common.cuh"
#include <stdlib.h>
#include <vector>

namespace kb200 {

template <bool EXACT>
__device__ __forceinline__ double mad2(double a, double b, double c) {
  if (EXACT) return __dadd_rn(__dmul_rn(a, b), c);
  return fma(a, b, c);
}

template <int GT, bool EXACT>
__global__ void __launch_bounds__(128) scattering_kernel(const kb200_scattering_desc *__restrict__ descs) {
  const kb200_scattering_desc &ds = descs[blockIdx.z];
  const int M = ds.M, Gs = ds.Gs, Zs = ds.Zs, G = ds.G;
  const Strides3 ms = strides_dgz(ds.layout, M, Gs, Zs);
  const Strides4 ss = strides_sigs(ds.layout, ds.L1, G);
  // flatten (nm, z) with the faster-varying of the two as the fast thread index
  const long long total = (long long)M * Zs;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  int nm, z;
  if (ms.a < ms.z) { nm = (int)(t % M); z = (int)(t / M); }
  else { z = (int)(t % Zs); nm = (int)(t / Zs); }
  const int n = ds.moment_to_legendre[nm];
  const int m0 = ds.zone_to_mixelem[z];
  const int nmix = ds.zone_to_num_mixelem[z];
  int mat[3] = {0, 0, 0};
  double frac[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int k = 0; k < 3; ++k)
    if (k < nmix) { mat[k] = ds.mixelem_to_material[m0 + k]; frac[k] = ds.mixelem_to_fraction[m0 + k]; }

  const int g0 = blockIdx.y * GT;
  double acc[GT];
#pragma unroll
  for (int i = 0; i < GT; ++i) acc[i] = 0.0;
  const long long base = (long long)nm * ms.a + (long long)z * ms.z;
  for (int s = 0; s < ds.nsrc; ++s) {
    const double *__restrict__ phi = ds.phi_src[s] + base;
    const int gl_src = ds.glower_src[s];
    for (int gp = 0; gp < Gs; ++gp) {
      const double x = __ldg(phi + (long long)gp * ms.g);
      const double *__restrict__ sp = ds.sigs + (long long)n * ss.n + (long long)(gp + gl_src) * ss.gp +
                                      (long long)(g0 + ds.glower_dst) * ss.g;
#pragma unroll
      for (int i = 0; i < GT; ++i) {
        if (g0 + i < Gs) {
          double sigs_z = 0.0;
#pragma unroll
          for (int k = 0; k < 3; ++k)
            if (k < nmix) sigs_z = mad2<EXACT>(__ldg(sp + (long long)mat[k] * ss.mat + (long long)i * ss.g), frac[k], sigs_z);
          acc[i] = mad2<EXACT>(sigs_z, x, acc[i]);
        }
      }
    }
  }
  double *__restrict__ out = ds.phi_out + base;
#pragma unroll
  for (int i = 0; i < GT; ++i)
    if (g0 + i < Gs) {
      double *q = out + (long long)(g0 + i) * ms.g;
      *q = ds.accumulate ? __dadd_rn(*q, acc[i]) : acc[i];
    }
}

__global__ void source_kernel(const kb200_source_desc *__restrict__ descs) {
  const kb200_source_desc &ds = descs[blockIdx.z];
  const Strides3 ms = strides_dgz(ds.layout, ds.M, ds.Gs, ds.Zs);
  const long long total = (long long)ds.num_mixelem * ds.Gs;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    int mix, g;
    if (ms.g == 1) { g = (int)(t % ds.Gs); mix = (int)(t / ds.Gs); }
    else { mix = (int)(t % ds.num_mixelem); g = (int)(t / ds.num_mixelem); }
    if (ds.mixelem_to_material[mix] == 0) {
      const int z = ds.mixelem_to_zone[mix];
      double *q = ds.phi_out + (long long)g * ms.g + (long long)z * ms.z;  // nm = 0
      *q = __dadd_rn(*q, __dmul_rn(ds.strength, ds.mixelem_to_fraction[mix]));
    }
  }
}

// Population: each block reduces a contiguous segment of one psi chunk; index decode is done on
// 32-bit segment-relative indices.  Block partials -> d_scratch, then one block sums them in a
// fixed order (deterministic for a given launch geometry).
constexpr int kPopThreads = 256;
constexpr int kPopSeg = 8192;  // elements per block-segment
constexpr int kPopMaxBlocks = 148 * 16;

__global__ void __launch_bounds__(kPopThreads) population_kernel(const kb200_population_desc *__restrict__ descs, int ndesc,
                                                                  long long segs_per_desc, double *__restrict__ partials) {
  __shared__ double red[kPopThreads / 32];
  double local = 0.0;
  const long long nseg_total = segs_per_desc * ndesc;
  for (long long seg = blockIdx.x; seg < nseg_total; seg += gridDim.x) {
    const int di = (int)(seg / segs_per_desc);
    const long long sbase = (seg - (long long)di * segs_per_desc) * kPopSeg;
    const kb200_population_desc &ds = descs[di];
    const long long total = (long long)ds.Ds * ds.Gs * ds.Zs;
    const Strides3 fs = strides_dgz(ds.layout, ds.Ds, ds.Gs, ds.Zs);
    // extents in storage order: n2 fastest
    long long n1, n2;
    int role1, role2, role0;  // 0 = direction, 1 = group, 2 = zone
    {
      long long st[3] = {fs.a, fs.g, fs.z};
      long long ex[3] = {ds.Ds, ds.Gs, ds.Zs};
      int o[3] = {0, 1, 2};
      for (int a = 0; a < 3; ++a)
        for (int b = a + 1; b < 3; ++b)
          if (st[o[b]] > st[o[a]]) { int tmp = o[a]; o[a] = o[b]; o[b] = tmp; }
      role0 = o[0]; role1 = o[1]; role2 = o[2];
      n1 = ex[role1]; n2 = ex[role2];
    }
    const long long inner = n1 * n2;
    for (int e = threadIdx.x; e < kPopSeg; e += kPopThreads) {
      const long long f = sbase + e;
      if (f >= total) break;
      const long long i0 = f / inner;
      const long long rem = f - i0 * inner;
      const long long i1 = rem / n2, i2 = rem - i1 * n2;
      long long idx[3];
      idx[role0] = i0; idx[role1] = i1; idx[role2] = i2;
      const double v = __ldg(ds.psi + f);
      local += __dmul_rn(__dmul_rn(__ldg(ds.w + idx[0]), v), __ldg(ds.volume + idx[2]));
    }
  }
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = (threadIdx.x < kPopThreads / 32) ? red[threadIdx.x] : 0.0;
    v = warp_sum(v);
    if (threadIdx.x == 0) partials[blockIdx.x] = v;
  }
}

// Vectorised population for the common case (fastest extent divisible by 4, 32-byte aligned chunks): a
// block owns a fixed slowest index i0 and a segment of the flattened inner (i1,i2) range, threads
// stream 32-byte vectors with 32-bit index arithmetic (one division per vector instead of two
// 64-bit divisions per element).  DR / ZR say which storage index (0 = slowest .. 2 = fastest) is
// the direction / the zone, i.e. the six storage orders of src/Kripke/VarTypes.h:73-101.
constexpr int kPopVecSeg = 8192;  // elements per work item
template <int DR, int ZR>
__global__ void __launch_bounds__(kPopThreads) population_vec_kernel(const kb200_population_desc *__restrict__ descs, int ndesc,
                                                                      int n0, int n1, int n2, double *__restrict__ partials) {
  __shared__ double red[kPopThreads / 32];
  const unsigned inner = (unsigned)n1 * (unsigned)n2;
  const unsigned segs = (inner + kPopVecSeg - 1) / kPopVecSeg;
  const unsigned items_per_desc = (unsigned)n0 * segs;
  const unsigned long long nitems = (unsigned long long)items_per_desc * ndesc;
  double local = 0.0;
  // zone-slowest orders have only Gs*Ds elements per slowest index: a whole block per item would idle most threads
  // and pay the item decode per 6 KB, so there a WARP owns an item
  const bool warp_items = inner <= 2048u;
  const unsigned tid = warp_items ? (threadIdx.x & 31u) : threadIdx.x, nthr = warp_items ? 32u : (unsigned)kPopThreads;
  const unsigned long long first = warp_items ? (unsigned long long)blockIdx.x * (kPopThreads / 32) + (threadIdx.x >> 5) : blockIdx.x;
  const unsigned long long stride = warp_items ? (unsigned long long)gridDim.x * (kPopThreads / 32) : gridDim.x;
  for (unsigned long long item = first; item < nitems; item += stride) {
    const unsigned di = (unsigned)(item / items_per_desc);
    const unsigned rem = (unsigned)(item - (unsigned long long)di * items_per_desc);
    const unsigned i0 = rem / segs, seg = rem - i0 * segs;
    const kb200_population_desc &ds = descs[di];
    const double *__restrict__ psi = ds.psi + (size_t)i0 * inner;
    const unsigned lo = seg * kPopVecSeg, hi = min(inner, lo + kPopVecSeg);
    for (unsigned j = lo + 4 * tid; j < hi; j += 4 * nthr) {
      const unsigned q = j / (unsigned)n2, r = j - q * (unsigned)n2;
      double v[4];
      asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                   : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(psi + j));
      double wv[4], zv[4];
      if (DR == 2) {
#pragma unroll
        for (int u = 0; u < 4; ++u) wv[u] = __ldg(ds.w + r + u);
      } else {
        const double w1 = __ldg(ds.w + (DR == 0 ? i0 : q));
        wv[0] = wv[1] = wv[2] = wv[3] = w1;
      }
      if (ZR == 2) {
#pragma unroll
        for (int u = 0; u < 4; ++u) zv[u] = __ldg(ds.volume + r + u);
      } else {
        const double z1 = __ldg(ds.volume + (ZR == 0 ? i0 : q));
        zv[0] = zv[1] = zv[2] = zv[3] = z1;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) local += __dmul_rn(__dmul_rn(wv[u], v[u]), zv[u]);  // (w*psi)*volume, Population.cpp:58
    }
  }
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = (threadIdx.x < kPopThreads / 32) ? red[threadIdx.x] : 0.0;
    v = warp_sum(v);
    if (threadIdx.x == 0) partials[blockIdx.x] = v;
  }
}

__global__ void population_final_kernel(const double *__restrict__ partials, int n, double *__restrict__ result, int accumulate) {
  __shared__ double red[32];
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v += partials[i];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    double s = (threadIdx.x < (int)(blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
    s = warp_sum(s);
    if (threadIdx.x == 0) *result = accumulate ? (*result + s) : s;
  }
}

__global__ void layout_transform_kernel(int src_layout, int dst_layout, int na, int ng, int nz,
                                        const double *__restrict__ src, double *__restrict__ dst) {
  const Strides3 a = strides_dgz(src_layout, na, ng, nz), b = strides_dgz(dst_layout, na, ng, nz);
  const long long total = (long long)na * ng * nz;
  // iterate in destination storage order (coalesced writes)
  long long st[3] = {b.a, b.g, b.z};
  long long ex[3] = {na, ng, nz};
  int o[3] = {0, 1, 2};
  for (int x = 0; x < 3; ++x)
    for (int y = x + 1; y < 3; ++y)
      if (st[o[y]] > st[o[x]]) { int tmp = o[x]; o[x] = o[y]; o[y] = tmp; }
  const long long n2 = ex[o[2]], n1 = ex[o[1]];
  for (long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x; f < total; f += (long long)gridDim.x * blockDim.x) {
    long long i0 = f / (n1 * n2), rem = f - i0 * n1 * n2, i1 = rem / n2, i2 = rem - i1 * n2;
    long long idx[3];
    idx[o[0]] = i0; idx[o[1]] = i1; idx[o[2]] = i2;
    dst[f] = src[idx[0] * a.a + idx[1] * a.g + idx[2] * a.z];
  }
}


// Tiled variant for the case that source and destination differ in their fastest index: a block moves one
// 32 x 32 tile of the (source-fastest, destination-fastest) plane through shared memory, so both the reads and
// the writes are 256-byte rows.  ext/ss/ds: extents and source/destination strides of (f = source-fastest index,
// d = destination-fastest index, t = the third one).
__global__ void __launch_bounds__(256) layout_transform_tiled_kernel(int nf, int nd, int nt, long long ss_d, long long ss_t,
                                                                     long long ds_f, long long ds_t,
                                                                     const double *__restrict__ src, double *__restrict__ dst) {
  __shared__ double tile[32][33];
  const int tf = (nf + 31) / 32, td = (nd + 31) / 32;
  const long long ntiles = (long long)tf * td * nt;
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;  // 32 x 8 threads
  for (long long b = blockIdx.x; b < ntiles; b += gridDim.x) {
    const int t = (int)(b / ((long long)tf * td));
    const int r = (int)(b - (long long)t * tf * td);
    const int f0 = (r % tf) * 32, d0 = (r / tf) * 32;
#pragma unroll
    for (int k = 0; k < 4; ++k) {  // rows of the source: f contiguous
      const int dd = d0 + ly + 8 * k, ff = f0 + lx;
      if (dd < nd && ff < nf) tile[ly + 8 * k][lx] = __ldg(src + (long long)t * ss_t + (long long)dd * ss_d + ff);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {  // rows of the destination: d contiguous
      const int ff = f0 + ly + 8 * k, dd = d0 + lx;
      if (ff < nf && dd < nd) dst[(long long)t * ds_t + (long long)ff * ds_f + dd] = tile[lx][ly + 8 * k];
    }
    __syncthreads();
  }
}

// Is every element of v[0..n) equal to v[0]?  flag[0] = number of blocks that saw a different value.
__global__ void uniform_check_kernel(const double *__restrict__ v, size_t n, unsigned *__restrict__ flag) {
  const double first = __ldg(v);
  bool differ = false;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) differ |= (__ldg(v + i) != first);
  if (__syncthreads_or(differ) && threadIdx.x == 0) atomicAdd(flag, 1u);
}

}  // namespace kb200

using namespace kb200;

int kb200_scatter_mma_try(const kb200_scattering_desc *h, int n, const void *d_descs, int sigs_layout, double source, cudaStream_t st);  // kb200_scatter_mma.cu
int kb200_scatter_row_try(const kb200_scattering_desc *h, int n, double source, cudaStream_t st);  // kb200_scatter_row.cu

extern "C" {

static int scattering_impl(const kb200_scattering_desc *h, int n, double source, int *folded, kb200_stream_t stream);
// which kernel family served the last scattering call: "slab" (one-read tensor-core kernel), "mma" (per-descriptor
// tensor-core kernel), "transposed+slab" / "transposed+mma" (moment-fastest nestings), "dfma" (bit-exact / odd shapes)
const char *g_last_scattering_kernel = "none";
const char *kb200_last_scattering_kernel(void) { return g_last_scattering_kernel; }
int kb200_scattering(const kb200_scattering_desc *h, int n, kb200_stream_t stream) {
  return scattering_impl(h, n, 0.0, nullptr, stream);
}
// Scattering with Kernel::source folded into its epilogue where the kernel in use supports it: *folded = 1
// means phi_out(0,g,z) has also received strength * (volume fraction of material 0 in zone z) and kb200_source must NOT be
// called; *folded = 0 means only the scattering was done.
int kb200_scattering_source(const kb200_scattering_desc *h, int n, double strength, int *folded, kb200_stream_t stream) {
  KB_REQUIRE(folded, "kb200_scattering_source: null argument");
  return scattering_impl(h, n, strength, folded, stream);
}
static int scattering_impl(const kb200_scattering_desc *h, int n, double source, int *folded, kb200_stream_t stream) {
  if (folded) *folded = 0;
  if (n <= 0) return 0;
  KB_REQUIRE(h, "kb200_scattering: null descriptors");
  for (int i = 0; i < n; ++i) {
    KB_REQUIRE(h[i].layout == h[0].layout && h[i].M == h[0].M && h[i].Gs == h[0].Gs && h[i].Zs == h[0].Zs,
               "kb200_scattering: all descriptors of one call must share their dimensions");
    KB_REQUIRE(h[i].nsrc >= 1 && h[i].nsrc <= KB200_MAX_DIRSETS, "kb200_scattering: nsrc=%d out of range", h[i].nsrc);
    KB_REQUIRE(h[i].layout >= 0 && h[i].layout < 6, "kb200_scattering: bad layout %d", h[i].layout);
    KB_REQUIRE(h[i].phi_out && h[i].sigs && h[i].moment_to_legendre && h[i].zone_to_mixelem && h[i].zone_to_num_mixelem &&
                   h[i].mixelem_to_material && h[i].mixelem_to_fraction, "kb200_scattering: null pointer in descriptor %d", i);
  }
  if (h[0].M <= 0 || h[0].Gs <= 0 || h[0].Zs <= 0) return 0;
  cudaStream_t st = resolve_stream(stream);
  const void *d = nullptr;
  int rc = device_descs(h, sizeof(*h) * n, &d, st);
  if (rc) return rc;
  if (!exact_mode()) {  // default arithmetic: fp64 tensor-core path for the zone-fastest layouts
    rc = kb200_scatter_mma_try(h, n, d, -1, source, st);  // sets g_last_scattering_kernel to "slab" or "mma"
    if (rc >= 0) { if (folded && rc == 0 && source != 0.0) *folded = 1; return rc; }
    rc = kb200_scatter_row_try(h, n, source, st);  // moment-fastest layouts: transposed through the zone-fastest kernel
    if (rc >= 0) {
      g_last_scattering_kernel = !strcmp(g_last_scattering_kernel, "slab") ? "transposed+slab" : "transposed+mma";
      if (folded && rc == 0 && source != 0.0) *folded = 1;
      return rc;
    }
  }
  g_last_scattering_kernel = "dfma";
  constexpr int GT = 8;
  long long total = (long long)h[0].M * h[0].Zs;
  dim3 grid((unsigned)((total + 127) / 128), (h[0].Gs + GT - 1) / GT, n);
  if (exact_mode()) scattering_kernel<GT, true><<<grid, 128, 0, st>>>((const kb200_scattering_desc *)d);
  else scattering_kernel<GT, false><<<grid, 128, 0, st>>>((const kb200_scattering_desc *)d);
  return post_launch("scattering");
}

int kb200_source(const kb200_source_desc *h, int n, kb200_stream_t stream) {
  if (n <= 0) return 0;
  KB_REQUIRE(h, "kb200_source: null descriptors");
  long long maxwork = 0;
  for (int i = 0; i < n; ++i) {
    KB_REQUIRE(h[i].layout >= 0 && h[i].layout < 6, "kb200_source: bad layout %d", h[i].layout);
    long long w = (long long)h[i].num_mixelem * h[i].Gs;
    if (w > maxwork) maxwork = w;
  }
  if (maxwork == 0) return 0;
  cudaStream_t st = resolve_stream(stream);
  const void *d = nullptr;
  int rc = device_descs(h, sizeof(*h) * n, &d, st);
  if (rc) return rc;
  long long blocks = (maxwork + 255) / 256;
  if (blocks > 4096) blocks = 4096;
  dim3 grid((unsigned)blocks, 1, n);
  source_kernel<<<grid, 256, 0, st>>>((const kb200_source_desc *)d);
  return post_launch("source");
}

size_t kb200_population_scratch_doubles(void) { return kPopMaxBlocks; }

int kb200_population_reduce(const double *d_partials, int n, double *d_result, kb200_stream_t stream) {
  KB_REQUIRE(d_result && (d_partials || n <= 0), "kb200_population_reduce: null pointer");
  population_final_kernel<<<1, 256, 0, resolve_stream(stream)>>>(d_partials, n > 0 ? n : 0, d_result, 0);
  return post_launch("population_final");
}

// *h_value = the common value of d_v[0..n) if all n doubles are bit-equal and positive, else 0.  Synchronises the stream
// (12 bytes come back instead of the whole array): used by the host layer to recognise a uniform zone-volume field.
int kb200_uniform_positive_value(const double *d_v, size_t n, double *h_value, kb200_stream_t stream) {
  KB_REQUIRE(h_value && (d_v || n == 0), "kb200_uniform_positive_value: null pointer");
  *h_value = 0.0;
  if (n == 0) return 0;
  cudaStream_t st = resolve_stream(stream);
  static unsigned *d_flag = nullptr;
  if (!d_flag) KB_CUDA(cudaMalloc(&d_flag, sizeof(unsigned)));
  KB_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(unsigned), st));
  size_t blocks = (n + 1023) / 1024;
  if (blocks > 592) blocks = 592;
  uniform_check_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_v, n, d_flag);
  int rc = post_launch("uniform_check");
  if (rc) return rc;
  unsigned differ = 1;
  double first = 0.0;
  KB_CUDA(cudaMemcpyAsync(&differ, d_flag, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
  KB_CUDA(cudaMemcpyAsync(&first, d_v, sizeof(double), cudaMemcpyDeviceToHost, st));
  KB_CUDA(cudaStreamSynchronize(st));
  if (!differ && first > 0.0) *h_value = first;
  return 0;
}

int kb200_population(const kb200_population_desc *h, int n, double *d_scratch, double *d_result, kb200_stream_t stream) {
  KB_REQUIRE(d_result && d_scratch, "kb200_population: null result/scratch");
  cudaStream_t st = resolve_stream(stream);
  if (n <= 0) {
    KB_CUDA(cudaMemsetAsync(d_result, 0, sizeof(double), st));
    return 0;
  }
  KB_REQUIRE(h, "kb200_population: null descriptors");
  long long maxtotal = 0;
  for (int i = 0; i < n; ++i) {
    KB_REQUIRE(h[i].layout >= 0 && h[i].layout < 6, "kb200_population: bad layout %d", h[i].layout);
    long long t = (long long)h[i].Ds * h[i].Gs * h[i].Zs;
    if (t > maxtotal) maxtotal = t;
  }
  const void *d = nullptr;
  int rc = device_descs(h, sizeof(*h) * n, &d, st);
  if (rc) return rc;
  {  // vectorised path: identical extents and layout, fastest extent % 4 == 0, 32-byte aligned chunks
    bool same = true;
    for (int i = 0; i < n; ++i)
      same = same && h[i].layout == h[0].layout && h[i].Ds == h[0].Ds && h[i].Gs == h[0].Gs && h[i].Zs == h[0].Zs &&
             ((uintptr_t)h[i].psi & 31) == 0;
    const int ext[3] = {h[0].Ds, h[0].Gs, h[0].Zs};
    static const int order[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};  // roles slowest..fastest (0=d,1=g,2=z)
    const int *o = order[h[0].layout];
    const int n0 = ext[o[0]], n1 = ext[o[1]], n2 = ext[o[2]];
    if (same && n2 % 4 == 0 && (double)n1 * n2 < 2147483648.0 && maxtotal > 0) {
      int dr = 0, zr = 0;
      for (int k = 0; k < 3; ++k) { if (o[k] == 0) dr = k; if (o[k] == 2) zr = k; }
      const long long segs = ((long long)n1 * n2 + kPopVecSeg - 1) / kPopVecSeg;
      const long long items = (long long)n * n0 * segs;
      const int blocks = (int)(items < kPopMaxBlocks ? items : kPopMaxBlocks);
      const kb200_population_desc *dd = (const kb200_population_desc *)d;
#define POPV(D, Z) population_vec_kernel<D, Z><<<blocks, kPopThreads, 0, st>>>(dd, n, n0, n1, n2, d_scratch)
      if (dr == 0 && zr == 2) POPV(0, 2); else if (dr == 0 && zr == 1) POPV(0, 1); else if (dr == 1 && zr == 2) POPV(1, 2);
      else if (dr == 2 && zr == 1) POPV(2, 1); else if (dr == 1 && zr == 0) POPV(1, 0); else POPV(2, 0);
#undef POPV
      rc = post_launch("population_vec");
      if (rc) return rc;
      population_final_kernel<<<1, 256, 0, st>>>(d_scratch, blocks, d_result, 0);
      return post_launch("population_final");
    }
  }
  long long segs_per_desc = (maxtotal + kPopSeg - 1) / kPopSeg;
  if (segs_per_desc == 0) segs_per_desc = 1;
  long long nseg = segs_per_desc * n;
  int blocks = (int)(nseg < kPopMaxBlocks ? nseg : kPopMaxBlocks);
  population_kernel<<<blocks, kPopThreads, 0, st>>>((const kb200_population_desc *)d, n, segs_per_desc, d_scratch);
  rc = post_launch("population");
  if (rc) return rc;
  population_final_kernel<<<1, 256, 0, st>>>(d_scratch, blocks, d_result, 0);
  return post_launch("population_final");
}

int kb200_layout_transform(int src_layout, int dst_layout, int na, int ng, int nz, const double *src, double *dst,
                           kb200_stream_t stream) {
  KB_REQUIRE(src_layout >= 0 && src_layout < 6 && dst_layout >= 0 && dst_layout < 6, "kb200_layout_transform: bad layout");
  KB_REQUIRE(src && dst && src != dst, "kb200_layout_transform: bad pointers (out-of-place only)");
  long long total = (long long)na * ng * nz;
  if (total <= 0) return 0;
  {  // different fastest index on the two sides: tiled transpose, both sides coalesced
    const Strides3 a = strides_dgz(src_layout, na, ng, nz), b = strides_dgz(dst_layout, na, ng, nz);
    const long long sst[3] = {a.a, a.g, a.z}, dst_[3] = {b.a, b.g, b.z};
    const int ext[3] = {na, ng, nz};
    int f = -1, d = -1;
    for (int x = 0; x < 3; ++x) { if (sst[x] == 1) f = x; if (dst_[x] == 1) d = x; }
    if (f >= 0 && d >= 0 && f != d) {
      const int t = 3 - f - d;
      const long long ntiles = (long long)((ext[f] + 31) / 32) * ((ext[d] + 31) / 32) * ext[t];
      long long blocks = ntiles < 148LL * 64 ? ntiles : 148LL * 64;
      layout_transform_tiled_kernel<<<(unsigned)blocks, 256, 0, resolve_stream(stream)>>>(ext[f], ext[d], ext[t], sst[d], sst[t], dst_[f],
                                                                                        dst_[t], src, dst);
      return post_launch("layout_transform_tiled");
    }
  }
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  layout_transform_kernel<<<(unsigned)blocks, 256, 0, resolve_stream(stream)>>>(src_layout, dst_layout, na, ng, nz, src, dst);
  return post_launch("layout_transform");
}

}  // extern "C"
