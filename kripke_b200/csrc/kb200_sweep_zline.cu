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

template <bool EXACT>
__global__ void __launch_bounds__(ZW * 32, 1) sweep_zline_kernel(const kb200_sweep_desc *__restrict__ descs, ZGeom gm) {
  extern __shared__ __align__(16) unsigned char zsm[];
  // shared memory: k-face exchange [2 buffers][2 halves][ZW*32] double2, then the per-direction tables
  double2 *fkx = reinterpret_cast<double2 *>(zsm);
  double *cxt = reinterpret_cast<double *>(zsm + (size_t)2 * 2 * ZW * 32 * sizeof(double2));  // [Ds]   2*xcos/dx[0]
  double *txc = cxt + gm.Ds;                                                                  // [Ds]   2*xcos
  double *cyt = txc + gm.Ds;                                                                  // [Ds][32]
  double *czt = cyt + (size_t)gm.Ds * 32;                                                     // [Ds][ZW]

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
  const bool line_ok = jj < tjn && kk < tkn;
  const int id = ds.id, jd = ds.jd, kd = ds.kd;
  const int jl = tj * jw + min(jj, tjn - 1), kl = tk * kw + min(kk, tkn - 1);
  const int jz = (jd > 0) ? jl : nj - 1 - jl, kz = (kd > 0) ? kl : nk - 1 - kl;

  const long long Zs = (long long)ni * nj * nk;
  const Strides3 fs = strides_dgz(gm.layout, Ds, gm.Gs, Zs);
  const StridesP ips = strides_plane(gm.layout, Ds, gm.Gs, nj, nk);
  const StridesP jps = strides_plane(gm.layout, Ds, gm.Gs, ni, nk);
  const StridesP kps = strides_plane(gm.layout, Ds, gm.Gs, ni, nj);
  const long long row_off = ((long long)kz * nj + jz) * ni;  // first zone of this line
  const long long ip_idx = (long long)jz + (long long)kz * nj;  // i_plane [k][j]
  const long long jp_row = (long long)kz * ni;                  // j_plane [k][i]
  const long long kp_row = (long long)jz * ni;                  // k_plane [j][i]

  const bool i_zero = ds.inflow_zero[0] != 0;
  const bool j_zero = ds.inflow_zero[1] != 0 && tj == 0;
  const bool k_zero = ds.inflow_zero[2] != 0 && tk == 0;
  const bool j_first = jj == 0, k_first = kk == 0;
  const bool j_last = jj == tjn - 1, k_last = kk == tkn - 1;
  double *__restrict__ iout = ds.out_plane[0];
  double *__restrict__ jout = (tj == gm.ntj - 1) ? ds.out_plane[1] : nullptr;
  double *__restrict__ kout = (tk == gm.ntk - 1) ? ds.out_plane[2] : nullptr;
  const double *__restrict__ rhs_b = ds.rhs;
  const double *__restrict__ sigt_b = ds.sigt;
  double *__restrict__ psi_b = ds.psi;
  double *ipl = ds.i_plane, *jpl = ds.j_plane, *kpl = ds.k_plane;

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
  const bool uniform_x = __syncthreads_and(uni) != 0;  // also orders the table writes

  // stream bookkeeping
  const int cnt = (stream < E) ? (E - stream + nstreams - 1) / nstreams : 0;
  const int T = cnt * nb;
  const int stream0 = blockIdx.x * spw;  // the longest stream of this CTA
  const int cnt0 = (stream0 < E) ? (E - stream0 + nstreams - 1) / nstreams : 0;
  const int Hend = cnt0 * nb + (tjn - 1) + (tkn - 1);

  // element state of the block being computed
  int c = nb - 1, e = stream - nstreams;
  long long off = 0, sgo = 0, ipo = 0, jpo = 0, kpo = 0;
  double cx = 0, cy = 0, cz = 0, csum = 0, tx = 0;
  // prefetched operands of the next block
  double r4n[4] = {0, 0, 0, 0}, s4n[4] = {1, 1, 1, 1}, fjbn[4] = {0, 0, 0, 0}, fkbn[4] = {0, 0, 0, 0}, finn = 0.0;
  double fi = 0.0;
  double oj[4] = {0, 0, 0, 0};

  auto elem_offsets = [&](int en, long long &o, long long &sg, long long &ip, long long &jp, long long &kp) {
    const int g = en / Ds, d = en - g * Ds;
    o = (long long)d * fs.a + (long long)g * fs.g;
    sg = (long long)g * Zs;
    ip = (long long)d * ips.d + (long long)g * ips.g;
    jp = (long long)d * jps.d + (long long)g * jps.g;
    kp = (long long)d * kps.d + (long long)g * kps.g;
  };

  for (int H = -1; H < Hend; ++H) {  // H = -1 only prefetches the first block of line (0,0)
    const int t = H - jj - kk;
    const bool act = line_ok && t >= 0 && t < T;
    const bool pre = line_ok && t + 1 >= 0 && t + 1 < T;

    // operands loaded during the previous macro-step
    double r4[4], s4[4], fjb[4], fkb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { r4[u] = r4n[u]; s4[u] = s4n[u]; fjb[u] = fjbn[u]; fkb[u] = fkbn[u]; }
    const double fin = finn;

    if (act) {
      if (++c == nb) {  // next element of the stream
        c = 0;
        e += nstreams;
        elem_offsets(e, off, sgo, ipo, jpo, kpo);
        const int d = e % Ds;
        cx = cxt[d]; tx = txc[d]; cy = cyt[d * 32 + lane]; cz = czt[d * ZW + warp];
        csum = __dadd_rn(__dadd_rn(cx, cy), cz);
      }
    }

    // ---- prefetch block t+1 ----
    if (pre) {
      int cn = c + 1;
      long long offn = off, sgn = sgo, ipn = ipo, jpn = jpo, kpn = kpo;
      if (!act || cn == nb) {
        cn = 0;
        elem_offsets(act ? e + nstreams : stream, offn, sgn, ipn, jpn, kpn);
        finn = i_zero ? 0.0 : ipl[ipn + ip_idx];
      }
      const int i0n = (id > 0) ? 4 * cn : ni - 4 - 4 * cn;
      ldg256_nc(rhs_b + offn + row_off + i0n, r4n);
      ldg256_nc(sigt_b + sgn + row_off + i0n, s4n);
      if (j_first && !j_zero) ldg256(jpl + jpn + jp_row + i0n, fjbn);
      if (k_first && !k_zero) ldg256(kpl + kpn + kp_row + i0n, fkbn);
    }

    // ---- incoming j faces: previous lane's outgoing faces of the previous macro-step ----
    double fj[4], fk[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const double v = __shfl_up_sync(0xffffffffu, oj[u], 1, jw);
      fj[u] = j_first ? (j_zero ? 0.0 : fjb[u]) : v;
    }
    // ---- incoming k faces: previous warp's outgoing faces of the previous macro-step ----
    {
      const double2 *rd = fkx + (size_t)((H & 1) ^ 1) * (2 * ZW * 32);
      if (k_first) {
#pragma unroll
        for (int u = 0; u < 4; ++u) fk[u] = k_zero ? 0.0 : fkb[u];
      } else {
        const double2 a = rd[(warp - 1) * 32 + lane], b = rd[ZW * 32 + (warp - 1) * 32 + lane];
        fk[0] = a.x; fk[1] = a.y; fk[2] = b.x; fk[3] = b.y;
      }
    }

    if (act) {
      if (c == 0) fi = fin;
      const int i0 = (id > 0) ? 4 * c : ni - 4 - 4 * c;
      double p4[4], ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int m = (id > 0) ? u : 3 - u;  // memory slot of the u-th zone in sweep order
        // select without dynamic register indexing
        const double r = (id > 0) ? r4[u] : r4[3 - u];
        const double st = (id > 0) ? s4[u] : s4[3 - u];
        const double fju = (id > 0) ? fj[u] : fj[3 - u];
        const double fku = (id > 0) ? fk[u] : fk[3 - u];
        double cxu = cx, cs = csum;
        if (!uniform_x) {
          cxu = __ddiv_rn(tx, ds.dx[i0 + m]);
          cs = __dadd_rn(__dadd_rn(cxu, cy), cz);
        }
        const double den = __dadd_rn(cs, st);
        double p;
        if (EXACT) {
          double num = __dadd_rn(r, __dmul_rn(fi, cxu));
          num = __dadd_rn(num, __dmul_rn(fju, cy));
          num = __dadd_rn(num, __dmul_rn(fku, cz));
          p = __ddiv_rn(num, den);
        } else {
          const double rc = __drcp_rn(den);
          const double part = fma(fku, cz, fma(fju, cy, r));
          const double num = fma(fi, cxu, part);
          const double q = num * rc;
          const double rem = fma(-den, q, num);
          p = fma(rem, rc, q);
        }
        const double p2 = 2.0 * p;
        fi = p2 - fi;
        const double ojv = p2 - fju, okv = p2 - fku;
        if (id > 0) { p4[u] = p; oj[u] = ojv; ok[u] = okv; }
        else { p4[3 - u] = p; oj[3 - u] = ojv; ok[3 - u] = okv; }
      }
      stg256(psi_b + off + row_off + i0, p4);
      if (c == nb - 1) {
        ipl[ipo + ip_idx] = fi;
        if (iout) iout[ipo + ip_idx] = fi;
      }
      if (j_last) {
        stg256(jpl + jpo + jp_row + i0, oj);
        if (jout) stg256(jout + jpo + jp_row + i0, oj);
      }
      if (k_last) {
        stg256(kpl + kpo + kp_row + i0, ok);
        if (kout) stg256(kout + kpo + kp_row + i0, ok);
      } else {
        double2 *wr = fkx + (size_t)(H & 1) * (2 * ZW * 32);
        wr[warp * 32 + lane] = make_double2(ok[0], ok[1]);
        wr[ZW * 32 + warp * 32 + lane] = make_double2(ok[2], ok[3]);
      }
    }
    __syncthreads();
  }
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
  if (h[0].ni % 4 != 0) return -1;
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
  gm.jw = pow2_ceil(gm.nj, 32);
  gm.kw = pow2_ceil(gm.nk, ZW);
  gm.ntj = (gm.nj + gm.jw - 1) / gm.jw;
  gm.ntk = (gm.nk + gm.kw - 1) / gm.kw;
  const size_t smem = (size_t)2 * 2 * ZW * 32 * sizeof(double2) + (size_t)gm.Ds * (2 + 32 + ZW) * sizeof(double);
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
