// Problem generator on the device: the zone loops of Kripke's setup that take seconds on the host at
// 128^3 zones -- material sub-sampling, the mixed-zone tables and the zonal total cross section
// (src/Kripke/Generate/Space.cpp:189-363).  Outputs are BIT-identical to the host generator: every floating-point
// expression is evaluated in the reference's order with explicit round-to-nearest multiplies, adds and divides (the
// reference's x86-64 build contracts nothing into FMA), and every table is filled in the reference's order.
#include "kb200_common.cuh"

namespace kb200 {

// Kobayashi-3i regions (Generate/Space.cpp:142-177)
__device__ __forceinline__ int material_at(double x, double y, double z) {
  x = fabs(x); y = fabs(y); z = fabs(z);
  if (x <= 10.0 && y <= 10.0 && z <= 10.0) return 0;
  if (x <= 10.0 && y <= 60.0 && z <= 10.0) return 1;
  if (x <= 40.0 && y >= 50.0 && y <= 60.0 && z <= 10.0) return 1;
  if (x >= 30.0 && x <= 40.0 && y >= 50.0 && y <= 60.0 && z <= 40.0) return 1;
  if (x >= 30.0 && x <= 40.0 && y >= 50.0 && z >= 30.0 && z <= 40.0) return 1;
  return 2;
}

// one thread per zone: ns^3 sub-samples at offsets (s+1)/(ns+1) of the zone width (Space.cpp:189-233)
__global__ void gen_fractions_kernel(kb200_zoneset_desc zs, double *__restrict__ frac, int *__restrict__ z2n) {
  const int Z = zs.ni * zs.nj * zs.nk;
  const int z = blockIdx.x * blockDim.x + threadIdx.x;
  if (z >= Z) return;
  const int i = z % zs.ni, j = (z / zs.ni) % zs.nj, k = z / (zs.ni * zs.nj);
  const double x0 = __dadd_rn(zs.x_min, __dmul_rn(zs.dx, (double)zs.i0));
  const double y0 = __dadd_rn(zs.y_min, __dmul_rn(zs.dy, (double)zs.j0));
  const double z0 = __dadd_rn(zs.z_min, __dmul_rn(zs.dz, (double)zs.k0));
  const double xi = __dadd_rn(x0, __dmul_rn(zs.dx, (double)i));
  const double yi = __dadd_rn(y0, __dmul_rn(zs.dy, (double)j));
  const double zi = __dadd_rn(z0, __dmul_rn(zs.dz, (double)k));
  const int ns = zs.ns;
  const double svf = __ddiv_rn(1.0, (double)(ns * ns * ns));
  double f[3] = {0.0, 0.0, 0.0};
  for (int si = 0; si < ns; ++si) {
    const double x = __dadd_rn(xi, __ddiv_rn(__dmul_rn(zs.dx, (double)(si + 1)), (double)(ns + 1)));
    for (int sj = 0; sj < ns; ++sj) {
      const double y = __dadd_rn(yi, __ddiv_rn(__dmul_rn(zs.dy, (double)(sj + 1)), (double)(ns + 1)));
      for (int sk = 0; sk < ns; ++sk) {
        const double zc = __dadd_rn(zi, __ddiv_rn(__dmul_rn(zs.dz, (double)(sk + 1)), (double)(ns + 1)));
        const int m = material_at(x, y, zc);
        // the same running sums, in the same order, as the host loop
        if (m == 0) f[0] = __dadd_rn(f[0], svf);
        else if (m == 1) f[1] = __dadd_rn(f[1], svf);
        else f[2] = __dadd_rn(f[2], svf);
      }
    }
  }
  frac[3 * (size_t)z] = f[0]; frac[3 * (size_t)z + 1] = f[1]; frac[3 * (size_t)z + 2] = f[2];
  z2n[z] = (f[0] > 0.0) + (f[1] > 0.0) + (f[2] > 0.0);
}

// exclusive scan of z2n into z2m, the total, and the material volumes: ONE block, every thread owns a contiguous run of
// zones (fixed summation order: run by run, then thread by thread)
constexpr int GEN_T = 1024;
__global__ void __launch_bounds__(GEN_T) gen_scan_kernel(int Z, const int *__restrict__ z2n, int *__restrict__ z2m,
                                                       const double *__restrict__ frac, double zone_volume, int *__restrict__ total,
                                                       double *__restrict__ volumes) {
  __shared__ int part[GEN_T];
  __shared__ double vpart[3][GEN_T];
  const int t = threadIdx.x;
  const int per = (Z + GEN_T - 1) / GEN_T;
  const int lo = min(t * per, Z), hi = min(lo + per, Z);
  int s = 0;
  double v[3] = {0.0, 0.0, 0.0};
  for (int z = lo; z < hi; ++z) {
    s += z2n[z];
#pragma unroll
    for (int m = 0; m < 3; ++m) v[m] = __dadd_rn(v[m], __dmul_rn(frac[3 * (size_t)z + m], zone_volume));
  }
  part[t] = s;
#pragma unroll
  for (int m = 0; m < 3; ++m) vpart[m][t] = v[m];
  __syncthreads();
  if (t == 0) {
    int run = 0;
    double tv[3] = {0.0, 0.0, 0.0};
    for (int q = 0; q < GEN_T; ++q) {
      const int c = part[q];
      part[q] = run;
      run += c;
#pragma unroll
      for (int m = 0; m < 3; ++m) tv[m] = __dadd_rn(tv[m], vpart[m][q]);
    }
    *total = run;
    volumes[0] = tv[0]; volumes[1] = tv[1]; volumes[2] = tv[2];
  }
  __syncthreads();
  int run = part[t];
  for (int z = lo; z < hi; ++z) {
    z2m[z] = run;
    run += z2n[z];
  }
}

// mixed-zone tables in zone order, materials ascending inside a zone (Space.cpp:285-320)
__global__ void gen_fill_kernel(int Z, const double *__restrict__ frac, const int *__restrict__ z2m, int *__restrict__ m2z,
                                int *__restrict__ m2m, double *__restrict__ m2f) {
  const int z = blockIdx.x * blockDim.x + threadIdx.x;
  if (z >= Z) return;
  int me = z2m[z];
#pragma unroll
  for (int m = 0; m < 3; ++m) {
    const double f = frac[3 * (size_t)z + m];
    if (f > 0.0) {
      m2z[me] = z;
      m2m[me] = m;
      m2f[me] = f;
      ++me;
    }
  }
}

// sigt_zonal(g, z) = sum over the zone's mixelems, in order, of fraction * sigt[material] (Space.cpp:331-363); the same
// value for every group of the set.  One thread per (zone, group) in storage order.
__global__ void gen_sigt_kernel(int Gs, int Z, int z_fastest, const int *__restrict__ z2m, const int *__restrict__ z2n,
                                const int *__restrict__ m2m, const double *__restrict__ m2f, double s0, double s1, double s2,
                                double *__restrict__ sigt) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)Gs * Z) return;
  const int z = z_fastest ? (int)(idx % Z) : (int)(idx / Gs);
  const int m0 = z2m[z], n = z2n[z];
  double acc = 0.0;
  for (int k = 0; k < n; ++k) {
    const int mat = m2m[m0 + k];
    acc = __dadd_rn(acc, __dmul_rn(m2f[m0 + k], mat == 0 ? s0 : (mat == 1 ? s1 : s2)));
  }
  sigt[idx] = acc;
}

}  // namespace kb200

using namespace kb200;

extern "C" int kb200_generate_mix_count(const kb200_zoneset_desc *zs, double *d_frac, int *d_z2n, int *d_z2m, int *h_num_mixelem,
                                        double *h_material_volume, kb200_stream_t stream) {
  KB_REQUIRE(zs && d_frac && d_z2n && d_z2m && h_num_mixelem && h_material_volume, "kb200_generate_mix_count: null argument");
  const int Z = zs->ni * zs->nj * zs->nk;
  KB_REQUIRE(Z > 0 && zs->ns > 0, "kb200_generate_mix_count: empty zone set");
  cudaStream_t st = resolve_stream(stream);
  gen_fractions_kernel<<<(Z + 255) / 256, 256, 0, st>>>(*zs, d_frac, d_z2n);
  int rc = post_launch("gen_fractions");
  if (rc) return rc;
  struct Out { int total; int pad; double vol[3]; };
  Out *d_out = nullptr;
  KB_CUDA(cudaMalloc(&d_out, sizeof(Out)));
  const double zone_volume = zs->dx * zs->dy * zs->dz;  // Space.cpp:134-136 (host arithmetic: evaluated left to right)
  gen_scan_kernel<<<1, GEN_T, 0, st>>>(Z, d_z2n, d_z2m, d_frac, zone_volume, &d_out->total, d_out->vol);
  rc = post_launch("gen_scan");
  if (rc) { cudaFree(d_out); return rc; }
  Out h;
  KB_CUDA(cudaMemcpyAsync(&h, d_out, sizeof(Out), cudaMemcpyDeviceToHost, st));
  KB_CUDA(cudaStreamSynchronize(st));
  KB_CUDA(cudaFree(d_out));
  *h_num_mixelem = h.total;
  h_material_volume[0] = h.vol[0]; h_material_volume[1] = h.vol[1]; h_material_volume[2] = h.vol[2];
  return 0;
}

extern "C" int kb200_generate_mix_fill(int Zs, const double *d_frac, const int *d_z2m, int *d_m2z, int *d_m2m, double *d_m2f,
                                       kb200_stream_t stream) {
  if (Zs <= 0) return 0;
  gen_fill_kernel<<<(Zs + 255) / 256, 256, 0, resolve_stream(stream)>>>(Zs, d_frac, d_z2m, d_m2z, d_m2m, d_m2f);
  return post_launch("gen_fill");
}

extern "C" int kb200_generate_sigt(int layout, int Gs, int Zs, const int *d_z2m, const int *d_z2n, const int *d_m2m,
                                   const double *d_m2f, const double *h_sigt3, double *d_sigt, kb200_stream_t stream) {
  if (Gs <= 0 || Zs <= 0) return 0;
  KB_REQUIRE(layout >= 0 && layout < 6 && h_sigt3, "kb200_generate_sigt: bad argument");
  long long sg, sz;
  strides_gz(layout, Gs, Zs, sg, sz);
  const size_t n = (size_t)Gs * Zs;
  gen_sigt_kernel<<<(unsigned)((n + 255) / 256), 256, 0, resolve_stream(stream)>>>(Gs, Zs, sz == 1 ? 1 : 0, d_z2m, d_z2n, d_m2m, d_m2f,
                                                                                 h_sigt3[0], h_sigt3[1], h_sigt3[2], d_sigt);
  return post_launch("gen_sigt");
}
