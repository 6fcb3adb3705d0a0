// LTimes / LPlusTimes on the fp64 tensor-core pipe (mma.sync m8n8k4 f64) for the storage orders whose
// long (group, zone) run is contiguous: DGZ, DZG, GDZ.
//
//   out[o][n] (+)= sum_k W[o][k] * in[k][n]        n = contiguous run of N elements, per batch b
//   LTimes      o = moment,              k = (direction set, direction)     (Kernel/LTimes.cpp:54-65)
//   LPlusTimes  o = (direction set, dir), k = moment                        (Kernel/LPlusTimes.cpp:49-60)
//
// Why tensor cores: at BASELINE config 2 both kernels sit on the HBM/fp64 ridge (5.5 flop/B with a
// measured 36.9 TFLOP/s DFMA peak and 6.5 TB/s HBM), so the contraction has to run near the fp64
// peak to stay bandwidth-bound.  A DFMA formulation needs one shared-memory weight operand per
// four DFMAs and saturates the LSU/issue ports first; DMMA issues 1/8 of the instructions and takes
// one 8-byte fragment load per 8x8x4 tile (ncu evidence in profiles/).
//
// Structure: persistent CTAs of 8 warps walk n-tiles of 64*NB columns.  The streamed operand is
// brought in by a 3-stage cp.async pipeline as [KC rows][tile] slabs (rows padded so that the
// 4x8 B-fragment read is bank-conflict free); the weights are converted once per CTA into
// fragment-major shared memory (one coalesced 8-byte read per A fragment).  Each warp owns 8*NB
// columns and keeps QP o-tiles of accumulators in registers.  Two regimes:
//   * K streamed, all outputs resident in registers   (LTimes: O = M small, K = D large)
//   * K resident (one slab per tile), outputs in passes of QP o-tiles over the same slab
//                                                      (LPlusTimes: K = M small, O = D large)
// Rows/columns beyond O, K are zero padded.  This path is not bit-ordered like the reference (the
// tensor core sums four products per step); EXACT mode uses the DFMA kernels in kb200_moments.cu.
#include "kb200_common.cuh"

namespace kb200 {

struct MomentsDescK {  // same as in kb200_moments.cu: pointer tables inside the device copy of the ABI descriptor
  const double *const *w;
  const double *const *in;
  double *const *out;
};

struct MmaGeom {
  int mode;  // 0 = LTimes, 1 = LPlusTimes
  int M, Ds, nsets;
  int O, K, q, nkc4;  // outputs, reduction length, o-tiles (8), k-chunks (4)
  int KC;             // rows per stage slab (multiple of 4)
  int nst;            // slabs per tile
  int npass;          // output passes per tile (npass > 1 only if nst == 1)
  int accumulate;
  long long B, N;     // batches, contiguous run
  long long in_b, in_r, out_b, out_r;  // batch / row strides of the streamed and the produced field
  long long ntn;      // tiles along n
  // packed mode (ZDG: runs of only pack_N = Gs columns per zone): the B = Zs runs are presented as ONE virtual run of
  // B * pack_N columns; virtual column v lives at (v / pack_N) * pack_{in,out}_b + v % pack_N
  int pack_N;
  long long pack_in_b, pack_out_b;
};

__device__ __forceinline__ void cp_async16_zfill(void *smem_dst, const void *gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// TMA bulk copy + mbarrier plumbing of the slab pipeline
__device__ __forceinline__ unsigned mm_smem(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mm_bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void mm_mb_init(unsigned addr, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mm_mb_expect_tx(unsigned addr, unsigned bytes) {
  asm volatile("{ .reg .b64 t; mbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1; }" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mm_mb_arrive(unsigned addr) {
  asm volatile("{ .reg .b64 t; mbarrier.arrive.shared::cta.b64 t, [%0]; }" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mm_mb_wait(unsigned addr, unsigned parity) {
  asm volatile("{\n .reg .pred p;\n MMW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @!p bra MMW;\n}" ::"r"(addr), "r"(parity) : "memory");
}

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}


// XMODE 1: O = 8*q + 1, the last output row is accumulated with DFMA on the B-fragment layout (LTimes, M = 25)
// XMODE 2: K = 4*nkc4 + 1, the last reduction row is added with DFMA on the C-fragment layout (LPlusTimes, M = 25)
template <int QP, int NB, int XMODE, int MMA_STAGES, bool TMA>
__global__ void __launch_bounds__(256, (QP <= 4 ? 2 : 1)) moments_mma_kernel(const MomentsDescK *__restrict__ descs, MmaGeom gm) {
  extern __shared__ __align__(16) double msm[];
  constexpr int NT = 64 * NB, NTP = NT + 4;  // row stride = 4 (mod 16) doubles: conflict-free B fragments
  constexpr int PPR = NT / 2;                // 16-byte pieces per slab row
  constexpr int RPP = 256 / PPR;             // slab rows filled per pass of the CTA
  double *Ws = msm;                                         // [nkc4][q][32] fragment-major weights
  double *slab = msm + (size_t)gm.nkc4 * gm.q * 32;         // [MMA_STAGES][KC][NTP]
  // row base pointers (batch 0, column 0) of the streamed rows / produced rows, null when padded
  const double **inrow = reinterpret_cast<const double **>(slab + (size_t)MMA_STAGES * gm.KC * NTP);  // [nst*KC]
  double **outrow = reinterpret_cast<double **>(slab + (size_t)MMA_STAGES * gm.KC * NTP + (size_t)gm.nst * gm.KC);  // [8*q+8]
  double *wx = reinterpret_cast<double *>(outrow + 8 * gm.q + 8);  // XMODE 1: [nst*KC] W[O-1][k] ; XMODE 2: [8*q] W[o][K-1]
  const MomentsDescK dsc = descs[blockIdx.y];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int M = gm.M, Ds = gm.Ds, O = gm.O, K = gm.K, q = gm.q, KC = gm.KC;

  // weights -> fragment-major, zero padded: Ws[(kc*q + ot)*32 + l] = W[8*ot + (l>>2)][4*kc + (l&3)]
  for (int idx = threadIdx.x; idx < gm.nkc4 * q * 32; idx += 256) {
    const int l = idx & 31, f = idx >> 5, ot = f % q, kc = f / q;
    const int o = 8 * ot + (l >> 2), k = 4 * kc + (l & 3);
    double v = 0.0;
    if (o < O && k < K) {
      if (gm.mode == 0) { const int s = k / Ds, d = k - s * Ds; v = dsc.w[s][(size_t)d * M + o]; }   // ell[d][nm]
      else { const int s = o / Ds, d = o - s * Ds; v = dsc.w[s][(size_t)d * M + k]; }                // ell_plus[d][nm]
    }
    Ws[idx] = v;
  }
  for (int k = threadIdx.x; k < gm.nst * KC; k += 256) {
    const double *p = nullptr;
    if (k < K) {
      if (gm.mode == 0) { const int s = k / Ds, d = k - s * Ds; p = dsc.in[s] + (long long)d * gm.in_r; }
      else p = dsc.in[0] + (long long)k * gm.in_r;
    }
    inrow[k] = p;
  }
  if (XMODE == 1)
    for (int k = threadIdx.x; k < gm.nst * KC; k += 256) {
      double v = 0.0;
      if (k < K) { const int s = k / Ds, d = k - s * Ds; v = dsc.w[s][(size_t)d * M + (O - 1)]; }
      wx[k] = v;
    }
  if (XMODE == 2)
    for (int o = threadIdx.x; o < 8 * q; o += 256) {
      double v = 0.0;
      if (o < O) { const int s = o / Ds, d = o - s * Ds; v = dsc.w[s][(size_t)d * M + (K - 1)]; }
      wx[o] = v;
    }
  for (int o = threadIdx.x; o < 8 * q + 8; o += 256) {
    double *p = nullptr;
    if (o < O) {
      if (gm.mode == 0) p = dsc.out[0] + (long long)o * gm.out_r;
      else { const int s = o / Ds, d = o - s * Ds; p = dsc.out[s] + (long long)d * gm.out_r; }
    }
    outrow[o] = p;
  }
  __syncthreads();

  // all loop control is 32-bit and incremental (no divisions inside the pipeline)
  const int ntn = (int)gm.ntn, nst = gm.nst;
  const long long ntiles = gm.B * gm.ntn;
  const int my_tiles = ((long long)blockIdx.x < ntiles) ? (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
  const int nitems = my_tiles * nst;

  // Slab pipeline: the rows of a slab ([KC rows][NT columns], one contiguous NT*8-byte run of the streamed field
  // each) are fetched by TMA bulk copies -- measured: per-lane cp.async (LDGSTS) saturates near 3.8 TB/s on this
  // part, the bulk-copy engine does not -- into MMA_STAGES buffers guarded by full/empty mbarriers.  Warp 0 is
  // the producer (lane r issues row r) in between its own tiles; every warp releases a buffer after its last read.
  // Rows beyond K in the last slab are not copied: their weights are zero and the stale (finite) contents of the
  // buffer contribute nothing; columns beyond N are not copied and never stored.
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(wx + (XMODE == 1 ? (size_t)gm.nst * KC : XMODE == 2 ? (size_t)8 * q : 0));
  const unsigned full0 = mm_smem(bars), empty0 = full0 + 8u * MMA_STAGES;
  if (TMA) {
    if (threadIdx.x < MMA_STAGES) { mm_mb_init(full0 + 8u * threadIdx.x, 1u); mm_mb_init(empty0 + 8u * threadIdx.x, 8u); }
    for (int i = threadIdx.x; i < MMA_STAGES * KC * NTP; i += 256) slab[i] = 0.0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the zero fill is ordered before the async-proxy writes
    __syncthreads();
  }
  int i_left = nitems, i_st = 0, i_tn = (int)(blockIdx.x % gm.ntn), i_b = (int)(blockIdx.x / gm.ntn), i_buf = 0, i_use = 0;
  const unsigned slab0 = mm_smem(slab);
  auto issue = [&]() {  // warp 0 only
    if (i_left > 0) {
      if (i_use >= MMA_STAGES) mm_mb_wait(empty0 + 8u * i_buf, ((unsigned)(i_use / MMA_STAGES) - 1u) & 1u);
      const long long n0 = (long long)i_tn * NT;
      const long long rem = gm.N - n0;
      const unsigned rowbytes = (unsigned)((rem < NT ? rem : NT) * 8);
      const long long boff = (long long)i_b * gm.in_b + n0;
      unsigned nvalid = 0;
      for (int r = lane; r < KC; r += 32) nvalid += (inrow[i_st * KC + r] != nullptr);
      nvalid = __reduce_add_sync(0xffffffffu, nvalid);
      if (lane == 0) mm_mb_expect_tx(full0 + 8u * i_buf, nvalid * rowbytes);
      __syncwarp();
      for (int r = lane; r < KC; r += 32) {
        const double *base = inrow[i_st * KC + r];
        if (base) mm_bulk_g2s(slab0 + (unsigned)((i_buf * KC + r) * NTP) * 8u, base + boff, rowbytes, full0 + 8u * i_buf);
      }
      --i_left;
      ++i_use;
      if (++i_st == nst) {
        i_st = 0;
        i_tn += gridDim.x;
        while (i_tn >= ntn) { i_tn -= ntn; ++i_b; }
      }
      if (++i_buf == MMA_STAGES) i_buf = 0;
    }
  };

  // K-streamed regime (many slabs per tile): per-warp cp.async pipelines, no coupling between the warps of a CTA
  // every warp copies and consumes only its own 8*NB columns of a slab: no CTA-wide barrier in the pipeline
  constexpr int PW = 4 * NB, RPW = 32 / PW;  // 16-byte pieces per row per warp, rows per warp instruction
  const int c2 = lane % PW, r0 = lane / PW;
  const int wcol0 = warp * 8 * NB;
  auto issue_cp = [&]() {
    if (i_left > 0) {
      double *dst = slab + (size_t)i_buf * KC * NTP + wcol0 + 2 * c2 + (size_t)r0 * NTP;
      const long long n = (long long)i_tn * NT + wcol0 + 2 * c2;
      const bool ncol = n < gm.N;
      const long long boff = (long long)i_b * gm.in_b + (gm.pack_N ? (n / gm.pack_N) * gm.pack_in_b + n % gm.pack_N : n);
      const double *const *rp = inrow + i_st * KC + r0;
      // interior slabs (every row exists, every column inside the run): plain 16-byte copies, no predicates
      const bool interior = ((i_st + 1) * KC <= K) && ((long long)(i_tn + 1) * NT <= gm.N);
      if (interior) {
        const unsigned d32 = (unsigned)__cvta_generic_to_shared(dst);
#pragma unroll 8
        for (int r = r0, o = 0; r < KC; r += RPW, rp += RPW, o += RPW * NTP * 8)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d32 + o), "l"(*rp + boff) : "memory");
      } else {
#pragma unroll 4
        for (int r = r0; r < KC; r += RPW, rp += RPW, dst += (size_t)RPW * NTP) {
          const double *base = *rp;
          const bool valid = ncol && base != nullptr;
          cp_async16_zfill(dst, valid ? base + boff : dsc.in[0], valid);
        }
      }
      --i_left;
      if (++i_st == nst) {
        i_st = 0;
        i_tn += gridDim.x;
        while (i_tn >= ntn) { i_tn -= ntn; ++i_b; }
      }
    }
    cp_async_commit();
    if (++i_buf == MMA_STAGES) i_buf = 0;
  };

  if (TMA) {
    if (warp == 0) {
#pragma unroll
      for (int s = 0; s < MMA_STAGES - 1; ++s) issue();
    }
  } else {
#pragma unroll
    for (int s = 0; s < MMA_STAGES - 1; ++s) issue_cp();
  }

  double acc[QP][NB][2];
  double px[NB];  // XMODE 1: partial sums of the extra output row (this lane's k residue class)
  const int ncol0 = warp * 8 * NB;  // this warp's first column inside the tile
  // consumer cursor
  int st = 0, tn = (int)(blockIdx.x % gm.ntn), b = (int)(blockIdx.x / gm.ntn), cbuf = 0;

  for (int item = 0; item < nitems; ++item) {
    if (TMA) {
      if (warp == 0) issue();  // refills the buffer every warp released after slab item-1
      mm_mb_wait(full0 + 8u * cbuf, (unsigned)(item / MMA_STAGES) & 1u);  // slab `item` has landed
    } else {
      cp_async_wait<MMA_STAGES - 2>();
      __syncwarp();  // this warp's columns of slab `item` have landed; its columns of slab item-1 are free
      issue_cp();
    }

    const double *buf = slab + (size_t)cbuf * KC * NTP;
    const int kc_lo = st * (KC / 4);
    const int nkc = min(gm.nkc4 - kc_lo, KC / 4);
    const double *bbase = buf + (size_t)(lane & 3) * NTP + ncol0 + (lane >> 2);

    double2 xin2[NB];  // XMODE 2: the reduction row K-1 at this lane's C-fragment columns, the same for every pass
    if (XMODE == 2) {
      const double *xr = buf + (size_t)(K - 1) * NTP + ncol0 + 2 * (lane & 3);
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) xin2[nb] = *reinterpret_cast<const double2 *>(xr + 8 * nb);
    }
    for (int pass = 0; pass < gm.npass; ++pass) {
      if (st == 0 || gm.npass > 1) {
#pragma unroll
        for (int a = 0; a < QP; ++a)
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) acc[a][nb][0] = acc[a][nb][1] = 0.0;
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) px[nb] = 0.0;
      }
      const int ot0 = pass * QP;
      double wk2[QP];  // XMODE 2: weights of the reduction row K-1 for this pass's output tiles
      if (XMODE == 2) {
#pragma unroll
        for (int a = 0; a < QP; ++a) wk2[a] = (ot0 + a < q) ? wx[8 * (ot0 + a) + (lane >> 2)] : 0.0;
      }
      const double *wf = Ws + ((size_t)kc_lo * q + ot0) * 32 + lane;
      const double *brow = bbase;
      if (ot0 + QP <= q) {  // full pass: no per-tile predicates
#pragma unroll 2
        for (int kc = 0; kc < nkc; ++kc, brow += 4 * NTP, wf += q * 32) {
          double bf[NB], af[QP];
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) bf[nb] = brow[8 * nb];
#pragma unroll
          for (int a = 0; a < QP; ++a) af[a] = wf[a * 32];
#pragma unroll
          for (int a = 0; a < QP; ++a)
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) dmma884(acc[a][nb][0], acc[a][nb][1], af[a], bf[nb]);
          if (XMODE == 1) {
            const double wxk = wx[4 * (kc_lo + kc) + (lane & 3)];
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) px[nb] = fma(wxk, bf[nb], px[nb]);
          }
        }
      } else {
        for (int kc = 0; kc < nkc; ++kc, brow += 4 * NTP, wf += q * 32) {
          double bf[NB];
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) bf[nb] = brow[8 * nb];
#pragma unroll
          for (int a = 0; a < QP; ++a) {
            if (ot0 + a < q) {
              const double af = wf[a * 32];
#pragma unroll
              for (int nb = 0; nb < NB; ++nb) dmma884(acc[a][nb][0], acc[a][nb][1], af, bf[nb]);
            }
          }
          if (XMODE == 1) {
            const double wxk = wx[4 * (kc_lo + kc) + (lane & 3)];
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) px[nb] = fma(wxk, bf[nb], px[nb]);
          }
        }
      }
      if (XMODE == 2) {  // the reduction row K-1 on the C-fragment layout (operands fetched before the k loop)
#pragma unroll
        for (int a = 0; a < QP; ++a) {
          if (ot0 + a < q) {
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
              acc[a][nb][0] = fma(wk2[a], xin2[nb].x, acc[a][nb][0]);
              acc[a][nb][1] = fma(wk2[a], xin2[nb].y, acc[a][nb][1]);
            }
          }
        }
      }
      if (XMODE == 1 && st == nst - 1) {  // extra output row: sum the four k residue classes, one lane per column stores
        double *row = outrow[O - 1] + (long long)b * gm.out_b;
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          double v = px[nb];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          const long long n = (long long)tn * NT + ncol0 + 8 * nb + (lane >> 2);
          if ((lane & 3) == 0 && n < gm.N) {
            const long long no = gm.pack_N ? (n / gm.pack_N) * gm.pack_out_b + n % gm.pack_N : n;
            if (gm.accumulate) v += row[no];
            row[no] = v;
          }
        }
      }
      if (st == nst - 1) {  // epilogue of this pass: C fragment (row = lane>>2, cols 2*(lane&3)+{0,1})
        const long long ncol = (long long)tn * NT + ncol0 + 2 * (lane & 3);
#pragma unroll
        for (int a = 0; a < QP; ++a) {
          const int o = 8 * (ot0 + a) + (lane >> 2);
          if (ot0 + a < q && o < O) {
            double *row = outrow[o] + (long long)b * gm.out_b;
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
              const long long n = ncol + 8 * nb;
              if (n < gm.N) {
                double2 v = make_double2(acc[a][nb][0], acc[a][nb][1]);
                double2 *p = reinterpret_cast<double2 *>(row + (gm.pack_N ? (n / gm.pack_N) * gm.pack_out_b + n % gm.pack_N : n));
                if (gm.accumulate) { const double2 old = *p; v.x += old.x; v.y += old.y; }
                *p = v;
              }
            }
          }
        }
      }
    }
    if (TMA) {
      __syncwarp();
      if (lane == 0) mm_mb_arrive(empty0 + 8u * cbuf);  // this warp is done with the buffer
    }
    if (++cbuf == MMA_STAGES) cbuf = 0;
    if (++st == nst) {
      st = 0;
      tn += gridDim.x;
      while (tn >= ntn) { tn -= ntn; ++b; }
    }
  }
  if (!TMA) cp_async_wait<0>();
}

}  // namespace kb200

using namespace kb200;

template <int QP, int NB, int XMODE, int MMA_STAGES, bool TMA>
static int launch_mma_t(const MomentsDescK *d_views, int n, const MmaGeom &gm, cudaStream_t st) {
  constexpr int NT = 64 * NB, NTP = NT + 4;
  const size_t smem = ((size_t)gm.nkc4 * gm.q * 32 + (size_t)MMA_STAGES * gm.KC * NTP + (size_t)gm.nst * gm.KC + 8 * gm.q + 8 +
                       (XMODE == 1 ? (size_t)gm.nst * gm.KC : XMODE == 2 ? (size_t)8 * gm.q : 0) + 2 * MMA_STAGES) * sizeof(double);
  if (smem > 200 * 1024) return -1;
  auto k = moments_mma_kernel<QP, NB, XMODE, MMA_STAGES, TMA>;
  KB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long ntiles = gm.B * gm.ntn;
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 2) per_sm = 2;
  long long ctas = (long long)sm_count() * per_sm / (n > 0 ? n : 1);
  if (ctas < 1) ctas = 1;
  if (ctas > ntiles) ctas = ntiles;
  dim3 grid((unsigned)ctas, n, 1);
  k<<<grid, 256, smem, st>>>(d_views, gm);
  return post_launch("moments_mma");
}

// One slab per tile (K resident, LPlusTimes): TMA bulk rows + CTA-level mbarriers (measured 6.4 -> 6.0 ms at config 2);
// many slabs per tile (K streamed, LTimes): the per-warp cp.async pipelines are faster (6.8 vs 8.7 ms), the CTA-level
// hand-off of a TMA pipeline costs more than the copy engine saves.
template <int QP, int NB, int XMODE, int MMA_STAGES>
static int launch_mma(const MomentsDescK *d_views, int n, const MmaGeom &gm, cudaStream_t st) {
  const char *env = getenv("KB200_MOMENTS_TMA");
  const bool tma = (env ? env[0] == '1' : (gm.nst == 1 && gm.K <= 32)) && gm.pack_N == 0;  // packed rows are not contiguous
  if (tma) return launch_mma_t<QP, NB, XMODE, MMA_STAGES, true>(d_views, n, gm, st);
  return launch_mma_t<QP, NB, XMODE, MMA_STAGES, false>(d_views, n, gm, st);
}

// Returns 0 if handled, -1 if this path does not apply (caller falls back to the DFMA kernels), >0 on error.
int kb200_ltimes_slab_try(int M, int Ds, int nsets, int accumulate, long long B, long long N, long long in_b, long long in_r,
                          long long out_b, long long out_r, const void *d_views, int n, const void *const *h_ptrs, int same_w,
                          cudaStream_t st);  // kb200_moments_slab.cu

int kb200_gemm_slab_try(int M, int Ds, int nsets, int accumulate, long long N, long long in_r, long long out_r, const void *d_views, int n,
                        const void *const *h_ptrs, int same_w, cudaStream_t st);  // kb200_scatter_slab.cu

int kb200_moments_mma_try(int mode, int layout, int M, int Ds, int Gs, int Zs, int nsets, int accumulate, const void *d_views,
                          int n, const void *const *h_ptrs, int n_ptrs, int same_w, cudaStream_t st) {
  if (layout != 0 && layout != 1 && layout != 2 && layout != 4) return -1;
  const char *env = getenv("KB200_MOMENTS_DFMA");
  if (env && env[0] == '1') return -1;
  for (int i = 0; i < n_ptrs; ++i)
    if (((uintptr_t)h_ptrs[i] & 15) != 0) return -1;
  MmaGeom gm;
  memset(&gm, 0, sizeof(gm));
  gm.mode = mode; gm.M = M; gm.Ds = Ds; gm.nsets = nsets; gm.accumulate = accumulate;
  const Strides3 fs = strides_dgz(layout, Ds, Gs, Zs), ms = strides_dgz(layout, M, Gs, Zs);
  long long flux_b, mom_b;
  if (layout == 2) { gm.B = Gs; gm.N = Zs; flux_b = fs.g; mom_b = ms.g; }     // GDZ: batch = group
  else if (layout == 4) {                                                     // ZDG: runs of Gs columns per zone, packed
    if (!(Gs % 16 == 0 && Gs <= 128 && 128 % Gs == 0)) return -1;             // a warp's 16 columns must stay inside one zone
    gm.B = 1; gm.N = (long long)Gs * Zs; flux_b = 0; mom_b = 0;
    gm.pack_N = Gs;
    gm.pack_in_b = (mode == 0) ? fs.z : ms.z;
    gm.pack_out_b = (mode == 0) ? ms.z : fs.z;
  } else { gm.B = 1; gm.N = (long long)Gs * Zs; flux_b = 0; mom_b = 0; }      // DGZ, DZG
  if (gm.N % 2 != 0) return -1;
  if (mode == 0) { gm.O = M; gm.K = nsets * Ds; gm.in_b = flux_b; gm.in_r = fs.a; gm.out_b = mom_b; gm.out_r = ms.a; }
  else { gm.O = nsets * Ds; gm.K = M; gm.in_b = mom_b; gm.in_r = ms.a; gm.out_b = flux_b; gm.out_r = fs.a; }
  gm.q = (gm.O + 7) / 8;
  gm.nkc4 = (gm.K + 3) / 4;
  const int Kp = gm.nkc4 * 4;
  const MomentsDescK *dv = (const MomentsDescK *)d_views;
  // LPlusTimes with a long reduction (M >= 36: fp64-bound, BASELINE config 3) and one contiguous run of columns per chunk:
  // the producer/consumer-group kernel of kb200_scatter_slab.cu as a plain product
  if (mode == 1 && gm.K >= 36 && gm.B == 1 && gm.pack_N == 0 && n_ptrs == n * (nsets + 1)) {
    const int rc = kb200_gemm_slab_try(M, Ds, nsets, accumulate, gm.N, gm.in_r, gm.out_r, d_views, n, h_ptrs, same_w, st);
    if (rc != -1) return rc;
  }
  // LTimes with M = 25 * c moments: producer warps + consumer groups, c sibling CTAs of 24 + 1 moments (kb200_moments_slab.cu)
  if (mode == 0 && M % 25 == 0 && gm.pack_N == 0 && n_ptrs == n * (nsets + 1)) {
    const int rc = kb200_ltimes_slab_try(M, Ds, nsets, accumulate, gm.B, gm.N, gm.in_b, gm.in_r, gm.out_b, gm.out_r, d_views, n, h_ptrs,
                                         same_w, st);
    if (rc != -1) return rc;
  }
  // regime: all outputs in registers with K streamed, or K resident with output passes
  if (gm.q <= 4 || (gm.q <= 13 && Kp > 32)) {
    gm.KC = Kp < 16 ? Kp : 16;
    gm.nst = (Kp + gm.KC - 1) / gm.KC;
    gm.npass = 1;
    gm.ntn = (gm.N + 127) / 128;
    if (gm.q == 4 && gm.O == 25) {  // 3 tensor-core tiles + one DFMA row instead of 4 tiles (22% less fp64 work)
      gm.q = 3;
      const char *e = getenv("KB200_LTIMES_KC");
      const int kc = e ? atoi(e) : 32;
      if (kc == 32 && Kp >= 64) {  // 32-row slabs, two stages: the per-slab bookkeeping (~200 instructions) is amortised over 48 DMMAs
        gm.KC = 32;
        gm.nst = (Kp + gm.KC - 1) / gm.KC;
        return launch_mma<3, 2, 1, 2>(dv, n, gm, st);
      }
      return launch_mma<3, 2, 1, 3>(dv, n, gm, st);
    }
    if (gm.q <= 4) return launch_mma<4, 2, 0, 3>(dv, n, gm, st);
    return launch_mma<13, 2, 0, 3>(dv, n, gm, st);
  }
  if (Kp <= 32) {  // K resident
    gm.KC = Kp; gm.nst = 1; gm.npass = (gm.q + 3) / 4;
    gm.ntn = (gm.N + 127) / 128;
    if (gm.K % 4 == 1 && gm.K > 4) {  // last reduction row by DFMA instead of a 3/4-empty k-chunk
      gm.nkc4 = gm.K / 4;
      return launch_mma<4, 2, 2, 2>(dv, n, gm, st);
    }
    return launch_mma<4, 2, 0, 2>(dv, n, gm, st);
  }
  if (gm.q <= 16) {  // both large (e.g. Legendre order 9): narrower tiles, all outputs in registers
    gm.KC = 16; gm.nst = (Kp + 15) / 16; gm.npass = 1;
    gm.ntn = (gm.N + 63) / 64;
    return launch_mma<16, 1, 0, 3>(dv, n, gm, st);
  }
  return -1;
}
