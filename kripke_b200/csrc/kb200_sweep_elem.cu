// SweepSubdomain for the element-fastest storage orders (DZG, GZD, ZDG, ZGD) on sm_100a.
//
// Reference arithmetic: src/Kripke/Kernel/SweepSubdomain.cpp:86-108.  In these layouts the
// (direction, group) elements of one zone are adjacent in memory, so
//   * lanes of a warp  <->  32 consecutive elements in storage order: every rhs / psi / plane access of
//     a warp is one coalesced 256-byte row, and all lanes advance in lock step (no divergence);
//   * a thread owns ELK = 4 zone lines stacked in k (same j), marching along i; the CTA's 16 warps
//     cover an 8 x 8 column tile of lines as a systolic array: line (jj,kk) is at zone
//     i = u - (kk mod 4) of its slice, where the thread-local step u runs over ni + 3 steps per
//     32-element slice so that the four lines of a thread always belong to the same slice;
//   * CTAs stream over slices (bx, bx + nx, ...) without draining the wavefront in between;
//   * i-face flux in a register, k-face flux between the stacked lines of a thread in registers,
//     j-face flux and the k-face flux between the two k-halves through double-buffered shared memory
//     (one __syncthreads per step); tile-boundary faces use the i/j/k_plane arrays in place;
//   * rhs and the tile-boundary j/k faces are fetched three steps ahead by cp.async into per-warp
//     rings (no registers, no scoreboard stalls); sigt comes one step ahead through L1.
// Column tiles of a subdomain form a 2-D wavefront of launches; all subdomains of a stage share them.
#include "kb200_common.cuh"

namespace kb200 {

constexpr int EW = 16;     // warps per CTA
constexpr int ETJ = 8, ETK = 8, ELK = 4;
constexpr int EPF = 3;     // cp.async distance in steps
constexpr int ERING = 4;   // ring slots

struct EGeom {
  int layout, Ds, Gs, ni, nj, nk;
  int ntj, ntk, diag, nslices;
};

__device__ __forceinline__ void e_cp_async8(void *smem_dst, const void *gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void e_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void e_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct SliceInfo {
  unsigned eoff, soff, ipo, jpo, kpo;  // element part of the psi/rhs, sigt and plane addresses of this lane
  int d;
  bool valid;
};

template <bool EXACT>
__global__ void __launch_bounds__(EW * 32, 1) sweep_elem_kernel(const kb200_sweep_desc *__restrict__ descs, EGeom gm) {
  extern __shared__ __align__(16) double esm[];
  double *fjx = esm;                                  // [2][ETK*ETJ][32]
  double *fkx = fjx + 2 * ETK * ETJ * 32;             // [2][EW][32]   (indexed by the receiving warp)
  double *rring = fkx + 2 * EW * 32;                  // [ERING][EW][ELK][32]
  double *jring = rring + ERING * EW * ELK * 32;      // [ERING][2][ELK][32]  boundary j faces of the jj == 0 warps
  double *kring = jring + ERING * 2 * ELK * 32;       // [ERING][ETJ][32]     boundary k faces of the kb == 0 warps (line 0)
  double *cxt = kring + ERING * ETJ * 32;             // [Ds] 2*xcos/dx[0]
  double *txc = cxt + gm.Ds;                          // [Ds] 2*xcos
  double *cyt = txc + gm.Ds;                          // [Ds][ETJ]
  double *czt = cyt + (size_t)gm.Ds * ETJ;            // [Ds][ETK]

  const kb200_sweep_desc &ds = descs[blockIdx.z];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int jj = warp & (ETJ - 1), kb = warp >> 3;
  const int Ds = gm.Ds, Gs = gm.Gs, ni = gm.ni, nj = gm.nj, nk = gm.nk;
  const int E = Ds * Gs;

  const int tj_lo = max(0, gm.diag - (gm.ntk - 1));
  const int tj = tj_lo + blockIdx.y, tk = gm.diag - tj;
  const int tjn = min(ETJ, nj - tj * ETJ), tkn = min(ETK, nk - tk * ETK);
  const int id = ds.id, jd = ds.jd, kd = ds.kd;
  const bool fwd = id > 0;
  const int jl = tj * ETJ + min(jj, tjn - 1);
  const int jz = (jd > 0) ? jl : nj - 1 - jl;
  int kz[ELK];
  bool line_ok[ELK];
#pragma unroll
  for (int m = 0; m < ELK; ++m) {
    const int kk = kb * ELK + m;
    const int kl = tk * ETK + min(kk, tkn - 1);
    kz[m] = (kd > 0) ? kl : nk - 1 - kl;
    line_ok[m] = jj < tjn && kk < tkn;
  }

  const long long Zs = (long long)ni * nj * nk;
  const Strides3 fs = strides_dgz(gm.layout, Ds, Gs, Zs);
  const StridesP ips = strides_plane(gm.layout, Ds, Gs, nj, nk);
  const StridesP jps = strides_plane(gm.layout, Ds, Gs, ni, nk);
  const StridesP kps = strides_plane(gm.layout, Ds, Gs, ni, nj);
  long long sg_l, sz_l;
  strides_gz(gm.layout, Gs, Zs, sg_l, sz_l);
  const unsigned fz = (unsigned)fs.z, sz = (unsigned)sz_l;
  const bool d_fast = fs.a < fs.g;  // which of (d,g) is the faster element index

  const bool i_zero = ds.inflow_zero[0] != 0;
  const bool j_zero = ds.inflow_zero[1] != 0 && tj == 0;
  const bool k_zero = ds.inflow_zero[2] != 0 && tk == 0;
  const bool j_first = jj == 0, j_last = jj == tjn - 1;
  const bool k_first = kb == 0;  // line 0 of this thread is the first k line of the tile
  double *__restrict__ iout = ds.out_plane[0];
  double *__restrict__ jout = (tj == gm.ntj - 1) ? ds.out_plane[1] : nullptr;
  double *__restrict__ kout = (tk == gm.ntk - 1) ? ds.out_plane[2] : nullptr;
  const double *__restrict__ rhs_b = ds.rhs;
  const double *__restrict__ sigt_b = ds.sigt;
  double *__restrict__ psi_b = ds.psi;
  double *ipl = ds.i_plane, *jpl = ds.j_plane, *kpl = ds.k_plane;

  // coefficient tables of this tile (2*cos/delta, SweepSubdomain.cpp:88-93)
  for (int d = threadIdx.x; d < Ds; d += blockDim.x) {
    const double t2 = 2.0 * ds.xcos[d];
    txc[d] = t2;
    cxt[d] = t2 / ds.dx[0];
  }
  for (int idx = threadIdx.x; idx < Ds * ETJ; idx += blockDim.x) {
    const int d = idx / ETJ, j = idx - d * ETJ;
    const int jl2 = tj * ETJ + min(j, tjn - 1);
    const int jz2 = (jd > 0) ? jl2 : nj - 1 - jl2;
    cyt[idx] = (2.0 * ds.ycos[d]) / ds.dy[jz2];
  }
  for (int idx = threadIdx.x; idx < Ds * ETK; idx += blockDim.x) {
    const int d = idx / ETK, k = idx - d * ETK;
    const int kl2 = tk * ETK + min(k, tkn - 1);
    const int kz2 = (kd > 0) ? kl2 : nk - 1 - kl2;
    czt[idx] = (2.0 * ds.zcos[d]) / ds.dz[kz2];
  }
  int uni = 1;
  for (int i = threadIdx.x; i < ni; i += blockDim.x) uni &= (ds.dx[i] == ds.dx[0]);
  const bool uniform_x = __syncthreads_and(uni) != 0;  // also orders the table writes

  // the slices of this CTA
  const int nx = gridDim.x;
  const int cnt = ((int)blockIdx.x < gm.nslices) ? (gm.nslices - blockIdx.x + nx - 1) / nx : 0;
  const int per = ni + ELK - 1;  // steps per slice for one thread
  const int Hend = cnt * per + tjn + tkn;

  auto slice_info = [&](int n) {
    SliceInfo s;
    const int sl = (n >= 0 && n < cnt) ? (int)blockIdx.x + n * nx : (int)blockIdx.x;
    int e = sl * 32 + lane;
    s.valid = e < E && n >= 0 && n < cnt;
    e = min(e, E - 1);
    int d, g;
    if (d_fast) { d = e % Ds; g = e / Ds; }
    else { g = e % Gs; d = e / Gs; }
    s.d = d;
    s.eoff = (unsigned)(d * fs.a + g * fs.g);
    s.soff = (unsigned)(g * sg_l);
    s.ipo = (unsigned)(d * ips.d + g * ips.g);
    s.jpo = (unsigned)(d * jps.d + g * jps.g);
    s.kpo = (unsigned)(d * kps.d + g * kps.g);
    return s;
  };

  // zone-index pieces of the four lines
  unsigned zrow[ELK], iprow[ELK], jprow[ELK], kprow;
#pragma unroll
  for (int m = 0; m < ELK; ++m) {
    zrow[m] = (unsigned)(((long long)kz[m] * nj + jz) * ni);
    iprow[m] = (unsigned)(jz * ips.a + kz[m] * ips.b);
    jprow[m] = (unsigned)(kz[m] * jps.b);
  }
  kprow = (unsigned)(jz * kps.b);
  const unsigned jpa = (unsigned)jps.a, kpa = (unsigned)kps.a;

  // thread-local schedule: n = slice ordinal, u = step inside the slice period
  int n = -2, u = -EPF - jj - ELK * kb + 2 * per;
  while (u >= per) { u -= per; ++n; }
  SliceInfo cur = slice_info(n), nxt = slice_info(n + 1);
  double cx = 0, cy = 0, cz[ELK], csum[ELK], tx = 0;
  auto load_coeffs = [&]() {
    cx = cxt[cur.d]; tx = txc[cur.d]; cy = cyt[cur.d * ETJ + jj];
#pragma unroll
    for (int m = 0; m < ELK; ++m) {
      cz[m] = czt[cur.d * ETK + kb * ELK + m];
      csum[m] = __dadd_rn(__dadd_rn(cx, cy), cz[m]);
    }
  };
  load_coeffs();

  double fi[ELK], okprev[ELK - 1], s_cur[ELK], fin[ELK];
#pragma unroll
  for (int m = 0; m < ELK; ++m) { fi[m] = 0.0; s_cur[m] = 1.0; fin[m] = 0.0; }
#pragma unroll
  for (int m = 0; m < ELK - 1; ++m) okprev[m] = 0.0;

  int rs = 1;  // ring slot of this step (the slot written now is rs + EPF mod ERING)
  double *my_r = rring + (size_t)warp * ELK * 32 + lane;                    // + slot*EW*ELK*32 + m*32
  double *my_j = jring + (size_t)kb * ELK * 32 + lane;                      // + slot*2*ELK*32 + m*32
  double *my_k = kring + (size_t)jj * 32 + lane;                            // + slot*ETJ*32
  const bool bnd_j = j_first && !j_zero, bnd_k = k_first && !k_zero;

#pragma unroll 1
  for (int h = -EPF; h < Hend; ++h) {
    const int par = h & 1;
    const int ws = (rs + EPF) & (ERING - 1);

    // ---- fetch the operands of step h + EPF ----
    {
      int u2 = u + EPF;
      const bool tonext = u2 >= per;
      if (tonext) u2 -= per;
      const unsigned eo = tonext ? nxt.eoff : cur.eoff;
      const bool sv = tonext ? nxt.valid : cur.valid;
#pragma unroll
      for (int m = 0; m < ELK; ++m) {
        const int i2 = u2 - m;
        if (line_ok[m] && sv && i2 >= 0 && i2 < ni) {
          const unsigned iz = fwd ? (unsigned)i2 : (unsigned)(ni - 1 - i2);
          e_cp_async8(my_r + (size_t)ws * EW * ELK * 32 + m * 32, rhs_b + (eo + (zrow[m] + iz) * fz));
          if (bnd_j) e_cp_async8(my_j + (size_t)ws * 2 * ELK * 32 + m * 32, jpl + ((tonext ? nxt.jpo : cur.jpo) + iz * jpa + jprow[m]));
          if (m == 0 && bnd_k) e_cp_async8(my_k + (size_t)ws * ETJ * 32, kpl + ((tonext ? nxt.kpo : cur.kpo) + iz * kpa + kprow));
          if (i2 == 0 && !i_zero) fin[m] = ipl[(tonext ? nxt.ipo : cur.ipo) + iprow[m]];
        }
      }
    }
    e_cp_async_commit();

    // ---- sigt of step h + 1 (through L1) ----
    double s_nxt[ELK];
    {
      int u1 = u + 1;
      const bool tonext = u1 >= per;
      if (tonext) u1 -= per;
      const unsigned so = tonext ? nxt.soff : cur.soff;
#pragma unroll
      for (int m = 0; m < ELK; ++m) {
        const int i1 = min(max(u1 - m, 0), ni - 1);
        const unsigned iz = fwd ? (unsigned)i1 : (unsigned)(ni - 1 - i1);
        s_nxt[m] = __ldg(sigt_b + (so + (zrow[m] + iz) * sz));
      }
    }

    e_cp_async_wait<EPF>();  // all but the EPF newest groups are complete: this thread's operands of step h have landed

    // ---- the four lines of this thread, highest k first (so that okprev[m-1] is still last step's value) ----
#pragma unroll
    for (int mm = 0; mm < ELK; ++mm) {
      const int m = ELK - 1 - mm;
      const int i = u - m;
      const bool act = line_ok[m] && cur.valid && i >= 0 && i < ni;
      const int ic = min(max(i, 0), ni - 1);
      const unsigned iz = fwd ? (unsigned)ic : (unsigned)(ni - 1 - ic);
      const int line = (kb * ELK + m) * ETJ + jj;
      // incoming faces
      double fj, fk;
      if (j_first) fj = j_zero ? 0.0 : my_j[(size_t)rs * 2 * ELK * 32 + m * 32];
      else fj = fjx[(par ^ 1) * (ETK * ETJ * 32) + (line - 1) * 32 + lane];
      if (m > 0) fk = okprev[m - 1];
      else if (k_first) fk = k_zero ? 0.0 : my_k[(size_t)rs * ETJ * 32];
      else fk = fkx[(par ^ 1) * (EW * 32) + warp * 32 + lane];
      if (i == 0) fi[m] = i_zero ? 0.0 : fin[m];
      const double r = my_r[(size_t)rs * EW * ELK * 32 + m * 32];
      const double st = s_cur[m];
      double cxu = cx, cs = csum[m];
      if (!uniform_x) {
        cxu = __ddiv_rn(tx, ds.dx[iz]);
        cs = __dadd_rn(__dadd_rn(cxu, cy), cz[m]);
      }
      const double den = __dadd_rn(cs, st);
      double p, oi, ojv, okv;
      if (EXACT) {
        double num = __dadd_rn(r, __dmul_rn(fi[m], cxu));
        num = __dadd_rn(num, __dmul_rn(fj, cy));
        num = __dadd_rn(num, __dmul_rn(fk, cz[m]));
        p = __ddiv_rn(num, den);
        const double p2 = 2.0 * p;
        oi = p2 - fi[m]; ojv = p2 - fj; okv = p2 - fk;
      } else {
        double y;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(den));
        double e = fma(-den, y, 1.0);
        y = fma(y, e, y);
        e = fma(-den, y, 1.0);
        const double rc = fma(y, e, y);
        const double part = fma(fk, cz[m], fma(fj, cy, r));
        const double A = cxu * rc, B = part * rc;
        p = fma(fi[m], A, B);
        oi = fma(fi[m], fma(2.0, A, -1.0), B + B);
        ojv = fma(2.0, p, -fj); okv = fma(2.0, p, -fk);
      }
      if (act) {
        fi[m] = oi;
        const unsigned zi = (zrow[m] + iz) * fz;
        if (cur.valid) psi_b[cur.eoff + zi] = p;
        if (i == ni - 1) {
          ipl[cur.ipo + iprow[m]] = oi;
          if (iout) iout[cur.ipo + iprow[m]] = oi;
        }
        if (j_last) {
          const unsigned ix = cur.jpo + iz * jpa + jprow[m];
          jpl[ix] = ojv;
          if (jout) jout[ix] = ojv;
        } else {
          fjx[par * (ETK * ETJ * 32) + line * 32 + lane] = ojv;
        }
        if (kb * ELK + m == tkn - 1) {
          const unsigned ix = cur.kpo + iz * kpa + kprow;
          kpl[ix] = okv;
          if (kout) kout[ix] = okv;
        } else if (m == ELK - 1) {
          fkx[par * (EW * 32) + (warp + ETJ) * 32 + lane] = okv;
        }
      }
      if (m < ELK - 1) okprev[m] = okv;
    }
#pragma unroll
    for (int m = 0; m < ELK; ++m) s_cur[m] = s_nxt[m];

    // ---- advance the thread-local schedule ----
    if (++u == per) {
      u = 0;
      ++n;
      cur = nxt;
      nxt = slice_info(n + 1);
      load_coeffs();
    }
    rs = (rs + 1) & (ERING - 1);
    __syncthreads();
  }
  e_cp_async_wait<0>();
}

}  // namespace kb200

using namespace kb200;

// Returns 0 if the batch was handled, -1 if this path does not apply (caller falls back), >0 on error.
int kb200_sweep_elem_try(const kb200_sweep_desc *h, int n, const void *d_descs, cudaStream_t st) {
  const int layout = h[0].layout;
  if (layout != 1 && layout != 3 && layout != 4 && layout != 5) return -1;
  const char *env = getenv("KB200_SWEEP_GENERIC");
  if (env && env[0] == '1') return -1;
  EGeom gm;
  gm.layout = layout; gm.Ds = h[0].Ds; gm.Gs = h[0].Gs; gm.ni = h[0].ni; gm.nj = h[0].nj; gm.nk = h[0].nk;
  if (gm.ni < 8) return -1;
  if ((double)gm.Ds * gm.Gs * gm.ni * gm.nj * gm.nk >= 2147483648.0) return -1;  // 32-bit element offsets
  gm.ntj = (gm.nj + ETJ - 1) / ETJ;
  gm.ntk = (gm.nk + ETK - 1) / ETK;
  const int E = gm.Ds * gm.Gs;
  gm.nslices = (E + 31) / 32;
  const size_t smem = ((size_t)2 * ETK * ETJ * 32 + 2 * EW * 32 + (size_t)ERING * EW * ELK * 32 + (size_t)ERING * 2 * ELK * 32 +
                       (size_t)ERING * ETJ * 32 + (size_t)gm.Ds * (2 + ETJ + ETK)) * sizeof(double);
  if (smem > 200 * 1024) return -1;
  auto kern = exact_mode() ? sweep_elem_kernel<true> : sweep_elem_kernel<false>;
  KB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int sms = sm_count();
  for (int diag = 0; diag < gm.ntj + gm.ntk - 1; ++diag) {
    gm.diag = diag;
    const int tj_lo = diag - (gm.ntk - 1) > 0 ? diag - (gm.ntk - 1) : 0;
    const int tj_hi = diag < gm.ntj - 1 ? diag : gm.ntj - 1;
    const int ntiles = tj_hi - tj_lo + 1;
    const long long pairs = (long long)ntiles * n;
    // CTAs per (tile, subdomain): minimise waves x steps per CTA (one CTA per SM)
    const int per = gm.ni + ELK - 1;
    int nx = 1;
    long long best = -1;
    for (int c = 1; c <= gm.nslices; ++c) {
      const long long waves = (c * pairs + sms - 1) / sms;
      const long long cost = waves * ((long long)((gm.nslices + c - 1) / c) * per + 40);
      if (best < 0 || cost < best) { best = cost; nx = c; }
    }
    dim3 grid(nx, ntiles, n);
    kern<<<grid, EW * 32, smem, st>>>((const kb200_sweep_desc *)d_descs, gm);
    int rc = post_launch("sweep_elem");
    if (rc) return rc;
  }
  return 0;
}
