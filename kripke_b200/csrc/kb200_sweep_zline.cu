// SweepSubdomain for the zone-fastest storage orders (DGZ, GDZ) on sm_100a.
//
// Reference arithmetic: src/Kripke/Kernel/SweepSubdomain.cpp:86-108.  In DGZ/GDZ the zones of one
// phase-space element (d,g) are contiguous (i fastest), so the only memory-efficient unit is an
// i-row; the kernel therefore maps
//   * one thread  <-> one zone line (fixed j,k of one element), marching along i in 32-byte
//     blocks of four zones (LDG.256 of rhs and sigt, STG.256 of psi: full sectors);
//   * lanes of a warp <-> consecutive j lines, warps <-> consecutive k lines of a TJ x TK column tile;
//   * the wavefront is block-skewed: at macro-step H line (jj,kk) works on block H - jj - kk of its
//     stream, where a stream is the sequence of elements e = s, s+S, s+2S.. assigned to the tile's
//     CTA -- lines never drain between elements, so the fill/drain cost of the hyperplane is paid
//     once per CTA and not once per element;
//   * the i-face flux stays in a register, the j-face flux moves to the next lane by warp shuffle,
//     the k-face flux moves to the next warp through double-buffered shared memory (one
//     __syncthreads per macro-step); only tile-boundary faces touch the i/j/k_plane arrays, which
//     are updated in place like the reference does;
//   * everything a block needs (rhs, sigt, boundary faces) is prefetched one macro-step ahead.
// Column tiles of a subdomain form a 2-D wavefront of launches (tile diagonals), all subdomains of
// the batch in the same launch.
//
// EXACT mode keeps the reference's multiply-then-add order and IEEE division; the default mode
// contracts to DFMA and divides with a correctly rounded reciprocal + one Markstein correction.
#include "kb200_common.cuh"

namespace kb200 {

constexpr int ZW = 16;  // warps per CTA

struct ZGeom {
  int layout, Ds, Gs, ni, nj, nk, nb;
  int ntj, ntk, diag;
  int jw, kw;  // lanes per j-group / warps per k-group (powers of two)
  int exp;     // timing experiments (KB200_ZEXP bitmask; results are wrong when non-zero)
};

__device__ __forceinline__ void ldg256_nc(const double *p, double (&v)[4]) {
  asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ void ldg256(const double *p, double (&v)[4]) {
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p) : "memory");
}
__device__ __forceinline__ void stg256(double *p, const double (&v)[4]) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void cp_async_wait_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// everything that is fixed for a thread during the whole launch
struct ZCtx {
  const double *rhs_b, *sigt_b;
  double *psi_b, *ipl, *jpl, *kpl, *iout, *jout, *kout;
  const double *dx;
  double2 *fkx, *fjx, *fkin;
  const double *cxt, *txc, *cyt, *czt;
  unsigned row_off, ip_idx, jp_row, kp_row;  // zone part of the line's addresses
  unsigned sa, sg, Zs;                       // element strides of rhs/psi, zones per subdomain
  unsigned ipd, ipg, jpd, jpg, kpd, kpg;     // element strides of the three planes
  int Ds, ni, nb, jw;
  int dv, gv, dstep, gstep;                  // virtual predecessor of the stream's first element, stream stride
  unsigned estep, ewrap, sstep;              // incremental element-offset updates (mod 2^32)
  int T, Hend, jjkk;
  bool line_ok, i_zero, j_zero, k_zero, j_first, k_first, j_last, k_last, uniform_x;
  int exp;
};

template <bool EXACT, bool FWD>
__device__ __forceinline__ void zline_run(const ZCtx &cx_) {
  const ZCtx c_ = cx_;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nb = c_.nb, ni = c_.ni;
  // Stream state = the block computed most recently.  Before the first block it is the last block of a
  // virtual predecessor element, so "the next block" is always derived by the same arithmetic.  All
  // element offsets are advanced incrementally modulo 2^32 (no multiplications in the loop).
  int c = nb - 1;
  int d = c_.dv, g = c_.gv;
  unsigned eoff = (unsigned)d * c_.sa + (unsigned)g * c_.sg;
  unsigned soff = (unsigned)g * c_.Zs;
  double cx = 0, cy = 0, cz = 0, csum = 0;
  double r4n[4] = {0, 0, 0, 0}, s4n[4] = {1, 1, 1, 1}, finn = 0.0;
  double fi = 0.0;
  double oj[4] = {0, 0, 0, 0};
  constexpr int XS = (ZW + 1) * 32;  // slots per half-buffer of the k exchange
  constexpr int JS = ZW * 32;        // slots per half-buffer of the boundary staging rings
  int ring = 0;                      // (H + 2) mod 3
  const bool bnd_j = c_.j_first && !c_.j_zero, bnd_k = c_.k_first && !c_.k_zero;

  int t = -2 - c_.jjkk;
#pragma unroll 1
  for (int H = -2; H < c_.Hend; ++H, ++t) {  // H < 0 only prefetches the first blocks of line (0,0)
    const bool act = c_.line_ok && (unsigned)t < (unsigned)c_.T;
    const bool pre = c_.line_ok && (unsigned)(t + 1) < (unsigned)c_.T;
    const bool pre2 = c_.line_ok && (unsigned)(t + 2) < (unsigned)c_.T;
    const int par = H & 1;

    double r4[4], s4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { r4[u] = r4n[u]; s4[u] = s4n[u]; }
    const double fin = finn;

    if (act) {
      if (c == nb - 1) {  // first block of the next element of the stream
        c = 0;
        const int dn = d + c_.dstep;
        const bool w = dn >= c_.Ds;
        d = w ? dn - c_.Ds : dn;
        g += c_.gstep + (w ? 1 : 0);
        eoff += c_.estep + (w ? c_.ewrap : 0u);
        soff += c_.sstep + (w ? c_.Zs : 0u);
        cx = c_.cxt[d]; cy = c_.cyt[d * 32 + lane]; cz = c_.czt[d * ZW + warp];
        csum = __dadd_rn(__dadd_rn(cx, cy), cz);
      } else {
        ++c;
      }
    }

    // the element after the current one (used when a prefetch crosses the element boundary)
    const int dn_ = d + c_.dstep;
    const bool wn = dn_ >= c_.Ds;
    const int dN = wn ? dn_ - c_.Ds : dn_, gN = g + c_.gstep + (wn ? 1 : 0);
    const unsigned eoffN = eoff + c_.estep + (wn ? c_.ewrap : 0u), soffN = soff + c_.sstep + (wn ? c_.Zs : 0u);

    // ---- prefetch block t+1 into registers ----
    if (pre) {
      const bool last = (c == nb - 1);
      const int c1 = last ? 0 : c + 1;
      const unsigned i0n = FWD ? 4u * c1 : (unsigned)(ni - 4 - 4 * c1);
      if (!(c_.exp & 4)) ldg256_nc(c_.rhs_b + ((last ? eoffN : eoff) + c_.row_off + i0n), r4n);
      if (!(c_.exp & 1)) ldg256_nc(c_.sigt_b + ((last ? soffN : soff) + c_.row_off + i0n), s4n);
      if (last) finn = c_.i_zero ? 0.0 : c_.ipl[(unsigned)dN * c_.ipd + (unsigned)gN * c_.ipg + c_.ip_idx];
    }

    // ---- tile-boundary faces of block t+2: cp.async into a 3-slot ring (waited for one macro-step later) ----
    if (pre2 && (bnd_j || bnd_k)) {
      const int cb = (t >= 0 ? c : nb + t) + 2;  // block index relative to the current (or virtual) element
      const bool nxt = cb >= nb;
      const int c2 = nxt ? cb - nb : cb;
      const int dB = nxt ? dN : d, gB = nxt ? gN : g;
      const unsigned i0f = FWD ? 4u * c2 : (unsigned)(ni - 4 - 4 * c2);
      if (bnd_j) {
        const double *src = c_.jpl + ((unsigned)dB * c_.jpd + (unsigned)gB * c_.jpg + c_.jp_row + i0f);
        cp_async16(c_.fjx + ring * (2 * JS) + threadIdx.x, src);
        cp_async16(c_.fjx + ring * (2 * JS) + JS + threadIdx.x, src + 2);
      }
      if (bnd_k) {
        const double *src = c_.kpl + ((unsigned)dB * c_.kpd + (unsigned)gB * c_.kpg + c_.kp_row + i0f);
        cp_async16(c_.fkin + ring * (2 * JS) + threadIdx.x, src);
        cp_async16(c_.fkin + ring * (2 * JS) + JS + threadIdx.x, src + 2);
      }
    }
    cp_async_commit();
    const int rcur = (ring == 2) ? 0 : ring + 1;  // H mod 3: the slot fetched two macro-steps ago

    // ---- incoming faces of block t: j from the previous lane (shuffle) or the tile boundary ring,
    //      k from the previous warp's exchange slot or the tile boundary ring ----
    double fj[4], fk[4];
    {
      double sj[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) sj[u] = __shfl_up_sync(0xffffffffu, oj[u], 1, c_.jw);
      if (c_.j_first) {
        if (c_.j_zero) { fj[0] = fj[1] = fj[2] = fj[3] = 0.0; }
        else {
          const double2 a = c_.fjx[rcur * (2 * JS) + threadIdx.x], b = c_.fjx[rcur * (2 * JS) + JS + threadIdx.x];
          fj[0] = a.x; fj[1] = a.y; fj[2] = b.x; fj[3] = b.y;
        }
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) fj[u] = sj[u];
      }
      if (c_.k_first) {
        if (c_.k_zero) { fk[0] = fk[1] = fk[2] = fk[3] = 0.0; }
        else {
          const double2 a = c_.fkin[rcur * (2 * JS) + threadIdx.x], b = c_.fkin[rcur * (2 * JS) + JS + threadIdx.x];
          fk[0] = a.x; fk[1] = a.y; fk[2] = b.x; fk[3] = b.y;
        }
      } else {
        const double2 a = c_.fkx[(par ^ 1) * (2 * XS) + warp * 32 + lane], b = c_.fkx[(par ^ 1) * (2 * XS) + XS + warp * 32 + lane];
        fk[0] = a.x; fk[1] = a.y; fk[2] = b.x; fk[3] = b.y;
      }
    }

    if (act) {
      if (c == 0) fi = fin;
      const unsigned i0 = FWD ? 4u * c : (unsigned)(ni - 4 - 4 * c);
      double p4[4], ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int m = FWD ? u : 3 - u;  // memory slot of the u-th zone in sweep order
        double cxu = cx, cs = csum;
        if (!c_.uniform_x) {
          cxu = __ddiv_rn(c_.txc[d], c_.dx[i0 + m]);
          cs = __dadd_rn(__dadd_rn(cxu, cy), cz);
        }
        const double den = __dadd_rn(cs, s4[m]);
        double p;
        if (EXACT) {
          double num = __dadd_rn(r4[m], __dmul_rn(fi, cxu));
          num = __dadd_rn(num, __dmul_rn(fj[m], cy));
          num = __dadd_rn(num, __dmul_rn(fk[m], cz));
          p = __ddiv_rn(num, den);
          const double p2 = 2.0 * p;
          fi = p2 - fi;
          p4[m] = p; oj[m] = p2 - fj[m]; ok[m] = p2 - fk[m];
        } else {
          // default arithmetic: psi = fi*A + B with A = cx/den, B = (rhs + fj*cy + fk*cz)/den and the
          // outgoing i face fi' = fi*(2A-1) + 2B, so the recurrence along i is ONE dependent DFMA per
          // zone; everything else (reciprocal by Newton from MUFU.RCP64H, A, B) is off the chain.
          double y;
          asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(den));
          double e = fma(-den, y, 1.0);
          y = fma(y, e, y);
          e = fma(-den, y, 1.0);
          const double rc = fma(y, e, y);
          const double part = fma(fk[m], cz, fma(fj[m], cy, r4[m]));
          const double A = cxu * rc, B = part * rc;
          p = fma(fi, A, B);
          fi = fma(fi, fma(2.0, A, -1.0), B + B);
          p4[m] = p; oj[m] = fma(2.0, p, -fj[m]); ok[m] = fma(2.0, p, -fk[m]);
        }
      }
      if (!(c_.exp & 2)) stg256(c_.psi_b + (eoff + c_.row_off + i0), p4);
      if (c == nb - 1) {
        const unsigned ix = (unsigned)d * c_.ipd + (unsigned)g * c_.ipg + c_.ip_idx;
        c_.ipl[ix] = fi;
        if (c_.iout) c_.iout[ix] = fi;
      }
      if (c_.j_last) {
        const unsigned ix = (unsigned)d * c_.jpd + (unsigned)g * c_.jpg + c_.jp_row + i0;
        stg256(c_.jpl + ix, oj);
        if (c_.jout) stg256(c_.jout + ix, oj);
      }
      if (c_.k_last) {
        const unsigned ix = (unsigned)d * c_.kpd + (unsigned)g * c_.kpg + c_.kp_row + i0;
        stg256(c_.kpl + ix, ok);
        if (c_.kout) stg256(c_.kout + ix, ok);
      } else {
        c_.fkx[par * (2 * XS) + (warp + 1) * 32 + lane] = make_double2(ok[0], ok[1]);
        c_.fkx[par * (2 * XS) + XS + (warp + 1) * 32 + lane] = make_double2(ok[2], ok[3]);
      }
    }
    if (!(c_.exp & 8)) cp_async_wait_but_one();  // the boundary faces of the NEXT macro-step have landed; the newest group may still fly
    if (!(c_.exp & 8)) __syncthreads();
    ring = rcur;
  }
  cp_async_wait_all();
}

template <bool EXACT>
__global__ void __launch_bounds__(ZW * 32, 1) sweep_zline_kernel(const kb200_sweep_desc *__restrict__ descs, ZGeom gm) {
  extern __shared__ __align__(16) unsigned char zsm[];
  // shared memory: k-face exchange [2 buffers][2 halves][(ZW+1)*32] double2, tile-boundary staging rings for the
  // j and k faces [3 slots][2 halves][ZW*32] double2 each, then the per-direction coefficient tables
  ZCtx c;
  c.fkx = reinterpret_cast<double2 *>(zsm);
  c.fjx = c.fkx + 2 * 2 * (ZW + 1) * 32;
  c.fkin = c.fjx + 3 * 2 * ZW * 32;
  double *cxt = reinterpret_cast<double *>(c.fkin + 3 * 2 * ZW * 32);  // [Ds]   2*xcos/dx[0]
  double *txc = cxt + gm.Ds;                                           // [Ds]   2*xcos
  double *cyt = txc + gm.Ds;                                           // [Ds][32]
  double *czt = cyt + (size_t)gm.Ds * 32;                              // [Ds][ZW]
  c.cxt = cxt; c.txc = txc; c.cyt = cyt; c.czt = czt;

  const kb200_sweep_desc &ds = descs[blockIdx.z];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Ds = gm.Ds, ni = gm.ni, nj = gm.nj, nk = gm.nk, nb = gm.nb;
  const int jw = gm.jw, kw = gm.kw;
  const int jj = lane & (jw - 1), jsub = lane / jw;
  const int kk = warp & (kw - 1), ksub = warp / kw;
  const int spw = (32 / jw) * (ZW / kw);
  const int stream = (blockIdx.x * (ZW / kw) + ksub) * (32 / jw) + jsub;
  const int nstreams = gridDim.x * spw;
  const int E = Ds * gm.Gs;

  // tile of this CTA on the current diagonal
  const int tj_lo = max(0, gm.diag - (gm.ntk - 1));
  const int tj = tj_lo + blockIdx.y, tk = gm.diag - tj;
  const int tjn = min(jw, nj - tj * jw), tkn = min(kw, nk - tk * kw);
  c.line_ok = jj < tjn && kk < tkn;
  const int jd = ds.jd, kd = ds.kd;
  const int jl = tj * jw + min(jj, tjn - 1), kl = tk * kw + min(kk, tkn - 1);
  const int jz = (jd > 0) ? jl : nj - 1 - jl, kz = (kd > 0) ? kl : nk - 1 - kl;

  const long long Zs = (long long)ni * nj * nk;
  const Strides3 fs = strides_dgz(gm.layout, Ds, gm.Gs, Zs);
  const StridesP ips = strides_plane(gm.layout, Ds, gm.Gs, nj, nk);
  const StridesP jps = strides_plane(gm.layout, Ds, gm.Gs, ni, nk);
  const StridesP kps = strides_plane(gm.layout, Ds, gm.Gs, ni, nj);
  c.row_off = (unsigned)(((long long)kz * nj + jz) * ni);  // first zone of this line
  c.ip_idx = (unsigned)(jz + kz * nj);                     // i_plane [k][j]
  c.jp_row = (unsigned)(kz * ni);                          // j_plane [k][i]
  c.kp_row = (unsigned)(jz * ni);                          // k_plane [j][i]
  c.sa = (unsigned)fs.a; c.sg = (unsigned)fs.g; c.Zs = (unsigned)Zs;
  c.ipd = (unsigned)ips.d; c.ipg = (unsigned)ips.g;
  c.jpd = (unsigned)jps.d; c.jpg = (unsigned)jps.g;
  c.kpd = (unsigned)kps.d; c.kpg = (unsigned)kps.g;
  c.Ds = Ds; c.ni = ni; c.nb = nb; c.jw = jw;

  c.i_zero = ds.inflow_zero[0] != 0;
  c.j_zero = ds.inflow_zero[1] != 0 && tj == 0;
  c.k_zero = ds.inflow_zero[2] != 0 && tk == 0;
  c.j_first = jj == 0; c.k_first = kk == 0;
  c.j_last = jj == tjn - 1; c.k_last = kk == tkn - 1;
  c.iout = ds.out_plane[0];
  c.jout = (tj == gm.ntj - 1) ? ds.out_plane[1] : nullptr;
  c.kout = (tk == gm.ntk - 1) ? ds.out_plane[2] : nullptr;
  c.rhs_b = ds.rhs; c.sigt_b = ds.sigt; c.psi_b = ds.psi;
  c.ipl = ds.i_plane; c.jpl = ds.j_plane; c.kpl = ds.k_plane;
  c.dx = ds.dx;
  c.exp = gm.exp;
  if (gm.exp & 16) { c.j_zero = c.k_zero = c.i_zero = true; c.j_last = c.k_last = false; }

  // per-direction coefficient tables (2*cos/delta, SweepSubdomain.cpp:88-93)
  for (int d = threadIdx.x; d < Ds; d += blockDim.x) {
    const double t2 = 2.0 * ds.xcos[d];
    txc[d] = t2;
    cxt[d] = t2 / ds.dx[0];
  }
  for (int idx = threadIdx.x; idx < Ds * 32; idx += blockDim.x) {
    const int d = idx >> 5, l = idx & 31;
    const int jl2 = tj * jw + min(l & (jw - 1), tjn - 1);
    const int jz2 = (jd > 0) ? jl2 : nj - 1 - jl2;
    cyt[idx] = (2.0 * ds.ycos[d]) / ds.dy[jz2];
  }
  for (int idx = threadIdx.x; idx < Ds * ZW; idx += blockDim.x) {
    const int d = idx / ZW, w = idx - d * ZW;
    const int kl2 = tk * kw + min(w & (kw - 1), tkn - 1);
    const int kz2 = (kd > 0) ? kl2 : nk - 1 - kl2;
    czt[idx] = (2.0 * ds.zcos[d]) / ds.dz[kz2];
  }
  int uni = 1;
  for (int i = threadIdx.x; i < ni; i += blockDim.x) uni &= (ds.dx[i] == ds.dx[0]);
  c.uniform_x = __syncthreads_and(uni) != 0;  // also orders the table writes

  // stream bookkeeping: this thread's line works on elements stream, stream+S, ...
  const int cnt = (stream < E) ? (E - stream + nstreams - 1) / nstreams : 0;
  c.T = cnt * nb;
  const int stream0 = blockIdx.x * spw;  // the longest stream of this CTA
  const int cnt0 = (stream0 < E) ? (E - stream0 + nstreams - 1) / nstreams : 0;
  c.Hend = cnt0 * nb + (tjn - 1) + (tkn - 1);
  c.jjkk = jj + kk;
  c.gstep = nstreams / Ds; c.dstep = nstreams - c.gstep * Ds;
  {
    const int g0 = stream / Ds, d0 = stream - g0 * Ds;  // first element; its virtual predecessor is one stream stride back
    int dv = d0 - c.dstep, gv = g0 - c.gstep;
    if (dv < 0) { dv += Ds; gv -= 1; }
    c.dv = dv; c.gv = gv;
  }
  c.estep = (unsigned)c.dstep * c.sa + (unsigned)c.gstep * c.sg;
  c.ewrap = c.sg - (unsigned)Ds * c.sa;
  c.sstep = (unsigned)c.gstep * c.Zs;

  if (ds.id > 0) zline_run<EXACT, true>(c);
  else zline_run<EXACT, false>(c);
}

}  // namespace kb200

using namespace kb200;

static int pow2_ceil(int v, int cap) {
  int p = 1;
  while (p < v && p < cap) p <<= 1;
  return p;
}

// Returns 0 if the batch was handled, -1 if this path does not apply (caller falls back), >0 on error.
int kb200_sweep_zline_try(const kb200_sweep_desc *h, int n, const void *d_descs, cudaStream_t st) {
  const int layout = h[0].layout;
  if (layout != 0 && layout != 2) return -1;
  if (h[0].ni % 4 != 0 || h[0].ni < 8) return -1;
  const char *env = getenv("KB200_SWEEP_GENERIC");
  if (env && env[0] == '1') return -1;
  for (int i = 0; i < n; ++i) {
    const void *ptrs[] = {h[i].rhs, h[i].psi, h[i].sigt, h[i].j_plane, h[i].k_plane, h[i].out_plane[1], h[i].out_plane[2]};
    for (const void *p : ptrs)
      if (((uintptr_t)p & 31) != 0) return -1;
  }
  ZGeom gm;
  gm.layout = layout; gm.Ds = h[0].Ds; gm.Gs = h[0].Gs; gm.ni = h[0].ni; gm.nj = h[0].nj; gm.nk = h[0].nk;
  gm.nb = gm.ni / 4;
  { const char *x = getenv("KB200_ZEXP"); gm.exp = x ? atoi(x) : 0; }
  gm.jw = pow2_ceil(gm.nj, 32);
  gm.kw = pow2_ceil(gm.nk, ZW);
  gm.ntj = (gm.nj + gm.jw - 1) / gm.jw;
  gm.ntk = (gm.nk + gm.kw - 1) / gm.kw;
  const size_t smem = ((size_t)2 * 2 * (ZW + 1) * 32 + (size_t)2 * 3 * 2 * ZW * 32) * sizeof(double2) + (size_t)gm.Ds * (2 + 32 + ZW) * sizeof(double);
  // the kernel indexes every chunk with 32-bit element offsets
  if ((double)gm.Ds * gm.Gs * gm.ni * gm.nj * gm.nk >= 2147483648.0) return -1;
  if (smem > 200 * 1024) return -1;
  auto kern = exact_mode() ? sweep_zline_kernel<true> : sweep_zline_kernel<false>;
  KB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int spw = (32 / gm.jw) * (ZW / gm.kw);
  const int E = gm.Ds * gm.Gs;
  const int max_nx = (E + spw - 1) / spw;
  const int sms = sm_count();
  for (int diag = 0; diag < gm.ntj + gm.ntk - 1; ++diag) {
    gm.diag = diag;
    const int tj_lo = diag - (gm.ntk - 1) > 0 ? diag - (gm.ntk - 1) : 0;
    const int tj_hi = diag < gm.ntj - 1 ? diag : gm.ntj - 1;
    const int ntiles = tj_hi - tj_lo + 1;
    const long long pairs = (long long)ntiles * n;
    // CTAs per (tile, subdomain): fill whole waves of one CTA per SM, keep streams long
    int nx = 1;
    if (pairs < sms) {
      nx = (int)(sms / pairs);
      // a second wave is worth it only if every stream stays long
      const int nx2 = (int)(2 * sms / pairs);
      if (nx2 > nx && (long long)nx2 * pairs * 2 > 3LL * sms && E / (nx2 * spw) >= 48) nx = nx2;
    }
    if (nx > max_nx) nx = max_nx;
    if (nx < 1) nx = 1;
    dim3 grid(nx, ntiles, n);
    kern<<<grid, ZW * 32, smem, st>>>((const kb200_sweep_desc *)d_descs, gm);
    int rc = post_launch("sweep_zline");
    if (rc) return rc;
  }
  return 0;
}
