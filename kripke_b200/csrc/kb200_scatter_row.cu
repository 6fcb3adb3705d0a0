// Scattering for the storage orders whose zone index is not the fastest one (GZD, ZGD; DZG, ZDG), default arithmetic.
//
// In these nestings phi[.][.][nm] has the moment index fastest, so the group-to-group contraction
// (Kernel/Scattering.cpp:73-99) would have to gather its operand with a 200-byte stride.  The moments
// field is small (16*N_m bytes against 40*N_u for the whole iteration), so the cheapest correct
// design is the layout-transform boundary of SURVEY 8b2: each source chunk is transposed once into
// the zone-fastest order [nm][g][z] (tiled through shared memory, both sides coalesced), the
// tensor-core kernel of kb200_scatter_mma.cu runs on the transposed chunks (sigs stays in its own
// nesting's order), and phi_out is transposed back -- with the caller's "+=" folded into that
// last pass.  Extra traffic: 32*N_m bytes, about 2 x 1.1 ms at BASELINE config 2, instead of the
// 30-50 ms of the strided DFMA kernel.  EXACT mode keeps the bit-ordered DFMA kernel.
#include "kb200_common.cuh"
#include <algorithm>
#include <map>
#include <vector>

namespace kb200 {

// one (group, 64-zone) tile: rows[z][nm] (moment-fastest side) <-> cols[nm][z] (zone-fastest side)
constexpr int TRZ = 64;

template <bool TO_ZONE_FASTEST>
__global__ void __launch_bounds__(256) moments_transpose_kernel(const double *__restrict__ src, double *__restrict__ dst, int M, int Gs,
                                                                int Zs, long long mf_sg, long long mf_sz, int accumulate) {
  extern __shared__ double tsm[];  // [TRZ][MP]
  const int MP = M | 1;            // odd row stride: conflict-free transposed access
  const int g = blockIdx.y;
  const int z0 = blockIdx.x * TRZ;
  const int nz = min(TRZ, Zs - z0);
  const long long zf_base = (long long)g * Zs + z0;          // + nm * Gs * Zs + zl   (zone-fastest side, DGZ order)
  const long long mf_base = (long long)g * mf_sg + (long long)z0 * mf_sz;  // + zl * mf_sz + nm (moment-fastest side)
  const long long zf_sa = (long long)Gs * Zs;
  if (TO_ZONE_FASTEST) {
    for (int f = threadIdx.x; f < nz * M; f += blockDim.x) {
      const int zl = f / M, nm = f - zl * M;
      tsm[zl * MP + nm] = __ldg(src + mf_base + (long long)zl * mf_sz + nm);
    }
    __syncthreads();
    for (int f = threadIdx.x; f < M * TRZ; f += blockDim.x) {
      const int nm = f / TRZ, zl = f - nm * TRZ;
      if (zl < nz) dst[zf_base + (long long)nm * zf_sa + zl] = tsm[zl * MP + nm];
    }
  } else {
    for (int f = threadIdx.x; f < M * TRZ; f += blockDim.x) {
      const int nm = f / TRZ, zl = f - nm * TRZ;
      if (zl < nz) tsm[zl * MP + nm] = __ldg(src + zf_base + (long long)nm * zf_sa + zl);
    }
    __syncthreads();
    for (int f = threadIdx.x; f < nz * M; f += blockDim.x) {
      const int zl = f / M, nm = f - zl * M;
      double *q = dst + mf_base + (long long)zl * mf_sz + nm;
      double v = tsm[zl * MP + nm];
      if (accumulate) v += *q;
      *q = v;
    }
  }
}

// ZGD ([z][g][nm]): a zone's (group, moment) block is one contiguous run of Gs*M doubles, so a tile takes ZT zones x GH
// groups: the moment-fastest side moves in runs of GH*M doubles (6.4 KB at config 2) and the zone-fastest side in rows of
// ZT zones (128 bytes), instead of the 200-byte runs at a 12.8 KB stride the (group, 64 zones) tiles above would read
// in this nesting (measured through the generic tiled transform: 3.2 TB/s for the four passes of a scattering call).
// ZT = 64 (512-byte rows on the zone-fastest side) or 16.
struct ZgdChunk {  // one chunk pair of a batched launch (blockIdx.z)
  const double *src;
  double *dst;
  long long accumulate;
};
template <bool TO_ZONE_FASTEST, int ZT>
__global__ void __launch_bounds__(256) moments_transpose_zgd_kernel(const ZgdChunk *__restrict__ chunks, int M, int Gs, int Zs, int GH) {
  extern __shared__ double tsm[];            // [ZT][P]
  constexpr int RL = ZT < 32 ? ZT : 32, PPR = ZT / RL, RPP = 256 / RL;  // lanes per row piece, pieces per row, pieces per pass
  const double *__restrict__ src = chunks[blockIdx.z].src;
  double *__restrict__ dst = chunks[blockIdx.z].dst;
  const int accumulate = chunks[blockIdx.z].accumulate;
  const int X = GH * M, P = X | 1;           // odd pitch: the 16 zones of a row sit in 16 different bank pairs
  const int z0 = blockIdx.x * ZT, g0 = blockIdx.y * GH;
  const int nz = min(ZT, Zs - z0), ng = min(GH, Gs - g0), nx = ng * M;
  const long long zf_sa = (long long)Gs * Zs;                      // zone-fastest side: nm * Gs*Zs + g * Zs + z
  const long long mf_base = (long long)z0 * Gs * M + (long long)g0 * M;  // moment-fastest side: + zl * Gs*M + x
  const int hl = threadIdx.x % RL, hw = threadIdx.x / RL;          // lane inside a row piece, piece of the pass
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int U = 8;  // loads in flight per thread
  // every zone's run starts on a 16-byte boundary and holds an even number of doubles
  const bool vec2 = ((Gs * M) % 2 == 0) && ((GH * M) % 2 == 0) && (nx % 2 == 0) && ((reinterpret_cast<uintptr_t>(TO_ZONE_FASTEST ? (const void *)src : (const void *)dst) & 15) == 0);
  if (TO_ZONE_FASTEST) {
    for (int zl = warp; zl < nz; zl += 8) {  // a warp per zone: 512-byte (or 256-byte) pieces of the zone's run
      const double *run = src + mf_base + (long long)zl * Gs * M;
      double *row = tsm + zl * P;
      if (vec2) {  // 16-byte loads: twice the bytes in flight per thread
        const double2 *run2 = reinterpret_cast<const double2 *>(run);
        for (int x0 = lane; x0 < nx / 2; x0 += 32 * U) {
          double2 v[U];
#pragma unroll
          for (int u = 0; u < U; ++u) v[u] = (x0 + 32 * u < nx / 2) ? __ldg(run2 + x0 + 32 * u) : make_double2(0.0, 0.0);
#pragma unroll
          for (int u = 0; u < U; ++u)
            if (x0 + 32 * u < nx / 2) { row[2 * (x0 + 32 * u)] = v[u].x; row[2 * (x0 + 32 * u) + 1] = v[u].y; }
        }
        continue;
      }
      for (int x0 = lane; x0 < nx; x0 += 32 * U) {
        double v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = (x0 + 32 * u < nx) ? __ldg(run + x0 + 32 * u) : 0.0;
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (x0 + 32 * u < nx) row[x0 + 32 * u] = v[u];
      }
    }
    __syncthreads();
    for (int pi = hw; pi < nx * PPR; pi += RPP) {
      const int x = pi / PPR, zl = (pi - x * PPR) * RL + hl, gl = x / M, nm = x - gl * M;
      if (zl < nz) dst[(long long)nm * zf_sa + (long long)(g0 + gl) * Zs + z0 + zl] = tsm[zl * P + x];
    }
  } else {
    for (int p0 = hw; p0 < nx * PPR; p0 += RPP * U) {
      double v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pi = p0 + RPP * u, x = pi / PPR, zl = (pi - x * PPR) * RL + hl, gl = x / M, nm = x - gl * M;
        v[u] = (pi < nx * PPR && zl < nz) ? __ldg(src + (long long)nm * zf_sa + (long long)(g0 + gl) * Zs + z0 + zl) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pi = p0 + RPP * u, x = pi / PPR, zl = (pi - x * PPR) * RL + hl;
        if (pi < nx * PPR) tsm[zl * P + x] = v[u];
      }
    }
    __syncthreads();
    for (int zl = warp; zl < nz; zl += 8) {
      double *run = dst + mf_base + (long long)zl * Gs * M;
      const double *row = tsm + zl * P;
      if (accumulate) {
        for (int x0 = lane; x0 < nx; x0 += 32 * U) {
          double v[U];
#pragma unroll
          for (int u = 0; u < U; ++u) v[u] = (x0 + 32 * u < nx) ? run[x0 + 32 * u] : 0.0;
#pragma unroll
          for (int u = 0; u < U; ++u)
            if (x0 + 32 * u < nx) run[x0 + 32 * u] = v[u] + row[x0 + 32 * u];
        }
      } else if (vec2) {
        double2 *run2 = reinterpret_cast<double2 *>(run);
        for (int x = lane; x < nx / 2; x += 32) run2[x] = make_double2(row[2 * x], row[2 * x + 1]);
      } else {
        for (int x = lane; x < nx; x += 32) run[x] = row[x];
      }
    }
  }
}

// scratch chunks in the zone-fastest order, kept between calls (freed when the process ends)
struct RowScratch {
  std::vector<double *> bufs;
  size_t bytes_each = 0;
};
static RowScratch g_row_scratch;

static int row_scratch(size_t count, size_t bytes_each, std::vector<double *> &out) {
  RowScratch &rs = g_row_scratch;
  // too small, or sized for a much bigger problem of the past (many small chunks must not each claim a huge buffer)
  if (rs.bytes_each < bytes_each || rs.bytes_each > 2 * bytes_each) {
    if (!rs.bufs.empty()) KB_CUDA(cudaDeviceSynchronize());
    for (double *p : rs.bufs) cudaFree(p);
    rs.bufs.clear();
    rs.bytes_each = bytes_each;
  }
  while (rs.bufs.size() < count) {
    double *p = nullptr;
    KB_CUDA(cudaMalloc(&p, rs.bytes_each));
    rs.bufs.push_back(p);
  }
  out.assign(rs.bufs.begin(), rs.bufs.begin() + count);
  return 0;
}

}  // namespace kb200

using namespace kb200;

int kb200_scatter_mma_try(const kb200_scattering_desc *h, int n, const void *d_descs, int sigs_layout, double source, cudaStream_t st);  // kb200_scatter_mma.cu

// Returns 0 if handled, -1 if this path does not apply (caller falls back to the DFMA kernel), >0 on error.
int kb200_scatter_row_try(const kb200_scattering_desc *h, int n, double source, cudaStream_t st) {
  const int layout = h[0].layout;
  if (layout != 1 && layout != 3 && layout != 4 && layout != 5) return -1;
  // The tiled generic layout transform is the faster transposer (measured 5.2 against 5.9-7.1 ms per scattering call)
  // but cannot fold a "+=": the moment-fastest nestings keep their own kernel for that case, the group-fastest ones
  // (DZG, ZDG) fall back to the DFMA kernel.
  bool any_acc = false;
  for (int i = 0; i < n; ++i) any_acc |= (h[i].accumulate != 0);
  if (any_acc && (layout == 1 || layout == 4)) return -1;
  const bool zgd = layout == 5 && !(getenv("KB200_ZGD_TRANSPOSE") && getenv("KB200_ZGD_TRANSPOSE")[0] == '0');
  const bool generic = !any_acc && !zgd;
  const char *env = getenv("KB200_SCATTER_DFMA");
  if (env && env[0] == '1') return -1;
  const int M = h[0].M, Gs = h[0].Gs, Zs = h[0].Zs;
  if (Zs % 2 != 0) return -1;
  const size_t chunk_bytes = (size_t)M * Gs * Zs * sizeof(double);
  // distinct source chunks of the batch (the destination group sets of one zone set share them)
  std::map<const double *, int> src_index;
  for (int i = 0; i < n; ++i)
    for (int s = 0; s < h[i].nsrc; ++s)
      if (!src_index.count(h[i].phi_src[s])) { const int id = (int)src_index.size(); src_index[h[i].phi_src[s]] = id; }
  const int nsrc = (int)src_index.size();
  std::vector<double *> bufs;
  int rc = row_scratch((size_t)nsrc + n, chunk_bytes, bufs);
  if (rc) return rc;
  const Strides3 ms = strides_dgz(layout, M, Gs, Zs);  // the moment-fastest side: ms.a == 1
  const size_t smem = (size_t)TRZ * (M | 1) * sizeof(double);
  if (!zgd && !generic) {
    if (smem > 200 * 1024) return -1;
    KB_CUDA(cudaFuncSetAttribute(moments_transpose_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    KB_CUDA(cudaFuncSetAttribute(moments_transpose_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const dim3 grid((Zs + TRZ - 1) / TRZ, Gs, 1);
  // ZGD tiles: ZT zones x GH groups, sized so that four blocks fit an SM (KB200_ZGD_ZT=64: 512-byte rows on the zone-fastest side)
  int ZT = 16;  // measured at config 2: 5.69 ms per scattering call with 16 zones per tile, 5.92 ms with 64
  { const char *ze = getenv("KB200_ZGD_ZT"); if (ze && atoi(ze) == 64) ZT = 64; }
  if ((size_t)ZT * (M | 1) * sizeof(double) > 54 * 1024 || Zs < 64) ZT = 16;
  int GH = Gs;
  while ((size_t)ZT * ((GH * M) | 1) * sizeof(double) > 54 * 1024 && GH > 1) GH = (GH + 1) / 2;
  const size_t zsmem = (size_t)ZT * ((GH * M) | 1) * sizeof(double);
  const dim3 zgrid((Zs + ZT - 1) / ZT, (Gs + GH - 1) / GH, 1);
  if (zgd) {
    KB_CUDA(cudaFuncSetAttribute(moments_transpose_zgd_kernel<true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zsmem));
    KB_CUDA(cudaFuncSetAttribute(moments_transpose_zgd_kernel<false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zsmem));
    KB_CUDA(cudaFuncSetAttribute(moments_transpose_zgd_kernel<true, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zsmem));
    KB_CUDA(cudaFuncSetAttribute(moments_transpose_zgd_kernel<false, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zsmem));
  }
  if (zgd) {  // all source chunks in one launch
    std::vector<ZgdChunk> zc;
    for (auto &kv : src_index) zc.push_back(ZgdChunk{kv.first, bufs[kv.second], 0});
    const void *dz = nullptr;
    rc = device_descs(zc.data(), zc.size() * sizeof(ZgdChunk), &dz, st);
    if (rc) return rc;
    for (size_t c0 = 0; c0 < zc.size(); c0 += 65535) {
      const unsigned nc = (unsigned)std::min<size_t>(65535, zc.size() - c0);
      if (ZT == 64) moments_transpose_zgd_kernel<true, 64><<<dim3(zgrid.x, zgrid.y, nc), 256, zsmem, st>>>((const ZgdChunk *)dz + c0, M, Gs, Zs, GH);
      else moments_transpose_zgd_kernel<true, 16><<<dim3(zgrid.x, zgrid.y, nc), 256, zsmem, st>>>((const ZgdChunk *)dz + c0, M, Gs, Zs, GH);
      rc = post_launch("moments_transpose_zgd");
      if (rc) return rc;
    }
  }
  for (auto &kv : src_index) {
    if (zgd) break;
    if (generic) {
      rc = kb200_layout_transform(layout, 0, M, Gs, Zs, kv.first, bufs[kv.second], st);
    } else {
      moments_transpose_kernel<true><<<grid, 256, smem, st>>>(kv.first, bufs[kv.second], M, Gs, Zs, ms.g, ms.z, 0);
      rc = post_launch("moments_transpose");
    }
    if (rc) return rc;
  }
  std::vector<kb200_scattering_desc> t(h, h + n);
  for (int i = 0; i < n; ++i) {
    t[i].layout = 0;
    t[i].accumulate = 0;
    for (int s = 0; s < t[i].nsrc; ++s) t[i].phi_src[s] = bufs[src_index[h[i].phi_src[s]]];
    t[i].phi_out = bufs[nsrc + i];
  }
  const void *d = nullptr;
  rc = device_descs(t.data(), sizeof(kb200_scattering_desc) * n, &d, st);
  if (rc) return rc;
  rc = kb200_scatter_mma_try(t.data(), n, d, layout, source, st);
  if (rc != 0) return rc < 0 ? -1 : rc;
  if (zgd) {  // all destination chunks in one launch
    std::vector<ZgdChunk> zc;
    for (int i = 0; i < n; ++i) zc.push_back(ZgdChunk{bufs[nsrc + i], h[i].phi_out, h[i].accumulate});
    const void *dz = nullptr;
    rc = device_descs(zc.data(), zc.size() * sizeof(ZgdChunk), &dz, st);
    if (rc) return rc;
    for (size_t c0 = 0; c0 < zc.size(); c0 += 65535) {
      const unsigned nc = (unsigned)std::min<size_t>(65535, zc.size() - c0);
      if (ZT == 64) moments_transpose_zgd_kernel<false, 64><<<dim3(zgrid.x, zgrid.y, nc), 256, zsmem, st>>>((const ZgdChunk *)dz + c0, M, Gs, Zs, GH);
      else moments_transpose_zgd_kernel<false, 16><<<dim3(zgrid.x, zgrid.y, nc), 256, zsmem, st>>>((const ZgdChunk *)dz + c0, M, Gs, Zs, GH);
      rc = post_launch("moments_transpose_zgd");
      if (rc) return rc;
    }
    return 0;
  }
  for (int i = 0; i < n; ++i) {
    if (generic) {
      rc = kb200_layout_transform(0, layout, M, Gs, Zs, bufs[nsrc + i], h[i].phi_out, st);
    } else {
      moments_transpose_kernel<false><<<grid, 256, smem, st>>>(bufs[nsrc + i], h[i].phi_out, M, Gs, Zs, ms.g, ms.z, h[i].accumulate);
      rc = post_launch("moments_transpose");
    }
    if (rc) return rc;
  }
  return 0;
}
