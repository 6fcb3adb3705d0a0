// SweepSubdomain for the element-fastest storage orders (DZG, GZD, ZDG, ZGD) on sm_100a: "pencil" kernel.
//
// Reference arithmetic: src/Kripke/Kernel/SweepSubdomain.cpp:86-108.  In these layouts the (direction, group)
// elements of one zone are adjacent in memory, so
//   * lanes of a warp  <->  32 consecutive elements in storage order (a "slice"): every rhs / psi / plane access of
//     a warp is one coalesced 256-byte row, every lane runs its own independent recurrence (no shuffles, no scan);
//   * a warp owns a PENCIL of PJ x PK = 4 x 2 zone lines and marches it along i.  Inside a step the eight zones are
//     visited in dependency order, so the j- and k-face fluxes between the lines of a pencil never leave registers and
//     the i-face flux of every line is a register carried from step to step: eight independent i recurrences per lane
//     give the fp64 pipe instruction-level parallelism without more warps;
//   * the 16 warps of a CTA form a 4 x 4 systolic array of pencils (a 16 x 8 column tile): pencil (J,K) is one step
//     behind (J-1,K) and (J,K-1).  Faces cross pencils through small shared-memory rings; there is NO CTA-wide
//     barrier in the loop: every warp publishes a monotone step counter (st.release.cta) and its two consumers
//     poll it (ld.acquire.cta), so a late load stalls one warp, not sixteen;
//   * rhs and sigt of the NEXT step are requested right after the current step's values have been consumed (one
//     register per line and array, no staging buffer): with 16 warps per SM a step of a warp recurs about every
//     1.5 us, so the loads have landed when they are needed;
//   * CTAs stream over slices without draining the systolic array; column tiles of a subdomain form a 2-D wavefront
//     of launches (tile diagonals), all subdomains of a stage share them; tile-boundary faces use the i/j/k_plane
//     arrays in place like the reference;
//   * optionally Kernel::population's sum (src/Kripke/Kernel/Population.cpp:49-63) is accumulated while psi is in
//     registers: one partial per CTA, summed later in a fixed order by kb200_population_reduce.
// EXACT mode keeps the reference's multiply-then-add order and IEEE division (bit-identical psi); the default
// contracts to DFMA and divides by reciprocal + cubic Newton step.
#include "kb200_common.cuh"
#include <vector>

namespace kb200 {

constexpr int PW = 16;                   // warps per CTA
constexpr int PWJ = 4, PWK = 4;          // pencils per tile
constexpr int PJ = 4, PK = 2;            // lines per pencil
constexpr int PL = PJ * PK;
constexpr int PTJ = PWJ * PJ, PTK = PWK * PK;  // column tile in lines
constexpr int PR = 4;                    // ring slots between neighbouring pencils

struct PGeom {
  int layout, Ds, Gs, ni, nj, nk;
  int ntj, ntk, diag, nslices;
  int pop;  // accumulate population partials
};

__device__ __forceinline__ void p_st_release(unsigned *p, unsigned v) {
  asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void p_wait_ge(const unsigned *p, int target) {  // the counter only grows
  const unsigned a = (unsigned)__cvta_generic_to_shared(p);
  unsigned v;
  do {
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  } while ((int)v < target);
}
__device__ __forceinline__ double p_ldg(const double *p) {  // streamed once: do not keep in L1
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double p_ld_cg(const double *p) {  // faces written by other CTAs / earlier launches: L2
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

struct PSlice {
  unsigned eoff, soff, ipo, jpo, kpo;  // element part of the psi/rhs, sigt and plane addresses of this lane
  int d;
  bool valid;
};

template <bool EXACT, bool POP>
__global__ void __launch_bounds__(PW * 32, 1) sweep_pencil_kernel(const kb200_sweep_desc *__restrict__ descs, const PGeom gm,
                                                                const double *const *__restrict__ pop_w,
                                                                const double *const *__restrict__ pop_vol,
                                                                const double *__restrict__ pop_vol_const,
                                                                double *__restrict__ pop_partial) {
  extern __shared__ __align__(16) double psm[];
  double *jring = psm;                                  // [PW][PR][PK][32]  j faces leaving the pencil of a warp
  double *kring = jring + PW * PR * PK * 32;            // [PW][PR][PJ][32]  k faces leaving the pencil of a warp
  double *red = kring + PW * PR * PJ * 32;              // [PW]
  double *coef = red + PW;                              // [2 + PJ + PK][PW * 32] 2*cos/delta of every thread's direction
  unsigned *cnt = reinterpret_cast<unsigned *>(coef + (2 + PJ + PK) * PW * 32);  // [PW] global steps finished by each warp

  __shared__ kb200_sweep_desc ds;  // read now and then (plane pointers, mesh): one LDS away instead of pinned registers
  if (threadIdx.x < sizeof(kb200_sweep_desc) / sizeof(int))
    reinterpret_cast<int *>(&ds)[threadIdx.x] = reinterpret_cast<const int *>(&descs[blockIdx.z])[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int J = warp & (PWJ - 1), K = warp / PWJ;
  const int Ds = gm.Ds, Gs = gm.Gs, ni = gm.ni, nj = gm.nj, nk = gm.nk;
  const int E = Ds * Gs;

  const int tj_lo = max(0, gm.diag - (gm.ntk - 1));
  const int tj = tj_lo + blockIdx.y, tk = gm.diag - tj;
  const int tjn = min(PTJ, nj - tj * PTJ), tkn = min(PTK, nk - tk * PTK);
  const int jd = ds.jd, kd = ds.kd;
  const bool fwd = ds.id > 0;

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
  const bool j_out = (tj == gm.ntj - 1), k_out = (tk == gm.ntk - 1);  // the tile's outgoing faces leave the subdomain
  const double *__restrict__ rhs_b = ds.rhs;
  const double *__restrict__ sigt_b = ds.sigt;
  double *__restrict__ psi_b = ds.psi;
#define ipl (ds.i_plane)
#define jpl (ds.j_plane)
#define kpl (ds.k_plane)

  if (threadIdx.x < PW) cnt[threadIdx.x] = 0u;
  int uni = 1;
  for (int i = threadIdx.x; i < ni; i += blockDim.x) uni &= (ds.dx[i] == ds.dx[0]);
  const bool uniform_x = __syncthreads_and(uni) != 0;  // also orders the counter initialisation

  // ---- the lines of this warp's pencil ----
  int jz[PJ], kz[PK];
  unsigned lmask = 0;       // bit l = jl * PK + kl: the line exists in this (possibly ragged) tile
  unsigned zrow[PL];         // first zone of the line: (kz * nj + jz) * ni
#pragma unroll
  for (int jl = 0; jl < PJ; ++jl) {
    const int jj = J * PJ + jl;
    const int jlz = tj * PTJ + min(jj, tjn - 1);
    jz[jl] = (jd > 0) ? jlz : nj - 1 - jlz;
  }
#pragma unroll
  for (int kl = 0; kl < PK; ++kl) {
    const int kk = K * PK + kl;
    const int klz = tk * PTK + min(kk, tkn - 1);
    kz[kl] = (kd > 0) ? klz : nk - 1 - klz;
  }
#pragma unroll
  for (int jl = 0; jl < PJ; ++jl)
#pragma unroll
    for (int kl = 0; kl < PK; ++kl) {
      const int l = jl * PK + kl;
      if (J * PJ + jl < tjn && K * PK + kl < tkn) lmask |= 1u << l;
      zrow[l] = (unsigned)((kz[kl] * nj + jz[jl]) * ni);
    }
  const bool any_line = lmask != 0;
  // the tile's last j / k line lies in this pencil: its outgoing faces go to the plane arrays, not to a neighbour pencil
  const bool j_tile_last = (tjn - 1) / PJ == J, k_tile_last = (tkn - 1) / PK == K;
  const bool j_in_ring = J > 0, k_in_ring = K > 0;
  const bool jb_load = !j_in_ring && !j_zero, kb_load = !k_in_ring && !k_zero;  // incoming faces come from the plane arrays
  const unsigned ipa = (unsigned)ips.a, ipb = (unsigned)ips.b, jpa = (unsigned)jps.a, jpb = (unsigned)jps.b,
                 kpa = (unsigned)kps.a, kpb = (unsigned)kps.b;

  // ---- the slices of this CTA ----
  const int nx = gridDim.x;
  const int cnt_sl = ((int)blockIdx.x < gm.nslices) ? (gm.nslices - (int)blockIdx.x + nx - 1) / nx : 0;
  const int T = cnt_sl * ni;                       // local steps of every warp
  const int Hend = T + (PWJ - 1) + (PWK - 1);      // global steps of the CTA
  auto slice_info = [&](int n) {
    PSlice s;
    const bool in = n >= 0 && n < cnt_sl;
    const int sl = in ? (int)blockIdx.x + n * nx : (int)blockIdx.x;
    int e = sl * 32 + lane;
    s.valid = e < E && in;
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

  PSlice cur = slice_info(0);
  double *mycoef = coef + threadIdx.x;  // [0] 2*xcos/dx[0], [1] 2*xcos, [2..] 2*ycos/dy[j], then 2*zcos/dz[k]
  auto load_coeffs = [&]() {  // 2*cos/delta of this lane's direction (SweepSubdomain.cpp:88-93); only this thread reads them
    const double tx = 2.0 * ds.xcos[cur.d];
    mycoef[0] = tx / ds.dx[0];
    mycoef[PW * 32] = tx;
    const double ty = 2.0 * ds.ycos[cur.d], tz = 2.0 * ds.zcos[cur.d];
#pragma unroll
    for (int jl = 0; jl < PJ; ++jl) mycoef[(2 + jl) * PW * 32] = ty / ds.dy[jz[jl]];
#pragma unroll
    for (int kl = 0; kl < PK; ++kl) mycoef[(2 + PJ + kl) * PW * 32] = tz / ds.dz[kz[kl]];
  };
  load_coeffs();

  const double *vol = POP ? pop_vol[blockIdx.z] : nullptr;
  const double vol_c = (POP && pop_vol_const) ? pop_vol_const[blockIdx.z] : 0.0;  // > 0: every zone has this volume
  double pop = 0.0, acc = 0.0;

  // operands of the current step (requested one step earlier) and the i-face flux of every line
  double R[PL], S[PL], FI[PL], JB[PK], KB[PJ];
  const unsigned i_first = fwd ? 0u : (unsigned)(ni - 1);
  const int istep = fwd ? 1 : -1;
  auto request = [&](const PSlice &sl, unsigned iz) {  // rhs, sigt and tile-boundary faces of zone column iz of slice sl
    const unsigned b = sl.eoff + iz * fz, sb = sl.soff + iz * sz;
#pragma unroll
    for (int l = 0; l < PL; ++l)
      if (lmask >> l & 1u) {
        R[l] = p_ldg(rhs_b + (zrow[l] * fz + b));
        S[l] = __ldg(sigt_b + (zrow[l] * sz + sb));
      }
    if (jb_load) {
#pragma unroll
      for (int kl = 0; kl < PK; ++kl)
        if (lmask >> kl & 1u) JB[kl] = p_ld_cg(jpl + (sl.jpo + iz * jpa + (unsigned)kz[kl] * jpb));
    }
    if (kb_load) {
#pragma unroll
      for (int jl = 0; jl < PJ; ++jl)
        if (lmask >> (jl * PK) & 1u) KB[jl] = p_ld_cg(kpl + (sl.kpo + iz * kpa + (unsigned)jz[jl] * kpb));
    }
  };
  auto load_iface = [&](const PSlice &sl) {
#pragma unroll
    for (int jl = 0; jl < PJ; ++jl)
#pragma unroll
      for (int kl = 0; kl < PK; ++kl) {
        const int l = jl * PK + kl;
        FI[l] = 0.0;
        if (!i_zero && (lmask >> l & 1u)) FI[l] = p_ld_cg(ipl + (sl.ipo + (unsigned)jz[jl] * ipa + (unsigned)kz[kl] * ipb));
      }
  };
#pragma unroll
  for (int l = 0; l < PL; ++l) { R[l] = 0.0; S[l] = 1.0; FI[l] = 0.0; }
#pragma unroll
  for (int kl = 0; kl < PK; ++kl) JB[kl] = 0.0;
#pragma unroll
  for (int jl = 0; jl < PJ; ++jl) KB[jl] = 0.0;
  if (any_line && cnt_sl > 0) {
    request(cur, i_first);
    load_iface(cur);
  }

  int n = 0, i = 0;           // slice ordinal and step inside the slice of the NEXT local step
  unsigned iz = i_first;      // memory i index of that step
  const int skew = J + K;
  double *my_j = jring + (size_t)warp * PR * PK * 32 + lane;
  double *my_k = kring + (size_t)warp * PR * PJ * 32 + lane;
  const double *in_j = jring + (size_t)(warp - 1) * PR * PK * 32 + lane;      // pencil (J-1,K)
  const double *in_k = kring + (size_t)(warp - PWJ) * PR * PJ * 32 + lane;    // pencil (J,K-1)

#pragma unroll 1
  for (int H = 0; H < Hend; ++H) {
    const int t = H - skew;
    if (any_line && t >= 0 && t < T) {
      const int ws = H & (PR - 1), rs = (H - 1) & (PR - 1);
      // faces of the upwind pencils (they finished global step H-1), and room in my own ring slot
      if (j_in_ring) p_wait_ge(cnt + warp - 1, H);
      if (k_in_ring) p_wait_ge(cnt + warp - PWJ, H);
      if (J < PWJ - 1) p_wait_ge(cnt + warp + 1, H - PR + 2);
      if (K < PWK - 1) p_wait_ge(cnt + warp + PWJ, H - PR + 2);

      const bool last_i = (i == ni - 1);
      const bool more = !last_i || (n + 1 < cnt_sl);
      // slice and zone column of the next local step
      const unsigned izn = last_i ? i_first : iz + (unsigned)istep;
      unsigned nb = cur.eoff, nsb = cur.soff, njp = cur.jpo, nkp = cur.kpo;
      if (last_i) {  // once per slice: the element offsets of the next slice
        const PSlice nx_ = slice_info(n + 1);
        nb = nx_.eoff; nsb = nx_.soff; njp = nx_.jpo; nkp = nx_.kpo;
      }
      nb += izn * fz; nsb += izn * sz; njp += izn * jpa; nkp += izn * kpa;
      const unsigned b = cur.eoff + iz * fz;
      double cxu = mycoef[0];
      if (!uniform_x) cxu = __ddiv_rn(mycoef[PW * 32], ds.dx[iz]);
      double cz[PK];
#pragma unroll
      for (int kl = 0; kl < PK; ++kl) cz[kl] = mycoef[(2 + PJ + kl) * PW * 32];

      // j faces entering the pencil: from pencil (J-1,K) or the tile boundary (JB stays 0 for vacuum inflow)
      double fj[PK];
#pragma unroll
      for (int kl = 0; kl < PK; ++kl) fj[kl] = j_in_ring ? in_j[(rs * PK + kl) * 32] : JB[kl];
#pragma unroll
      for (int jl = 0; jl < PJ; ++jl) {
        if (!(lmask >> (jl * PK) & 1u)) continue;
        double fk = k_in_ring ? in_k[(rs * PJ + jl) * 32] : KB[jl];
        const double cyj = mycoef[(2 + jl) * PW * 32];
#pragma unroll
        for (int kl = 0; kl < PK; ++kl) {
          const int l = jl * PK + kl;
          if (kl > 0 && !(lmask >> l & 1u)) continue;
          const double r = R[l], st = S[l];
          const double fi = FI[l];
          if (more) {  // operands of the next step of this line
            R[l] = p_ldg(rhs_b + (zrow[l] * fz + nb));
            S[l] = __ldg(sigt_b + (zrow[l] * sz + nsb));
          }
          const double cs = __dadd_rn(__dadd_rn(cxu, cyj), cz[kl]);
          const double den = __dadd_rn(cs, st);
          double p;
          if (EXACT) {
            double num = __dadd_rn(r, __dmul_rn(fi, cxu));
            num = __dadd_rn(num, __dmul_rn(fj[kl], cyj));
            num = __dadd_rn(num, __dmul_rn(fk, cz[kl]));
            p = __ddiv_rn(num, den);
            const double p2 = 2.0 * p;
            FI[l] = p2 - fi; fj[kl] = p2 - fj[kl]; fk = p2 - fk;
          } else {
            double y;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(den));
            const double e1 = fma(-den, y, 1.0);  // cubic step: y*(1 + e + e^2), e ~ 2^-23 -> error ~ 2^-69
            const double e2 = fma(e1, e1, e1);
            const double rc = fma(y, e2, y);
            const double num = fma(fk, cz[kl], fma(fj[kl], cyj, fma(fi, cxu, r)));
            p = num * rc;
            FI[l] = fma(2.0, p, -fi); fj[kl] = fma(2.0, p, -fj[kl]); fk = fma(2.0, p, -fk);
          }
          if (cur.valid) {
            psi_b[zrow[l] * fz + b] = p;
            if (POP) {
              if (vol_c > 0.0) acc += p;
              else acc = fma(p, __ldg(vol + (zrow[l] + iz)), acc);
            }
          }
        }
        // k face leaving the pencil on this j line (fk belongs to the last existing k line)
        if (k_tile_last) {
          if (cur.valid) {
            const unsigned ix = cur.kpo + iz * kpa + (unsigned)jz[jl] * kpb;
            kpl[ix] = fk;
            if (k_out && ds.out_plane[2]) ds.out_plane[2][ix] = fk;
          }
        } else {
          my_k[(ws * PJ + jl) * 32] = fk;
        }
      }
      // j faces leaving the pencil (fj belongs to the last existing j line)
      if (j_tile_last) {
        if (cur.valid) {
#pragma unroll
          for (int kl = 0; kl < PK; ++kl)
            if (lmask >> kl & 1u) {
              const unsigned ix = cur.jpo + iz * jpa + (unsigned)kz[kl] * jpb;
              jpl[ix] = fj[kl];
              if (j_out && ds.out_plane[1]) ds.out_plane[1][ix] = fj[kl];
            }
        }
      } else {
#pragma unroll
        for (int kl = 0; kl < PK; ++kl) my_j[(ws * PK + kl) * 32] = fj[kl];
      }
      if (last_i && cur.valid) {  // i faces leaving the subdomain
#pragma unroll
        for (int jl = 0; jl < PJ; ++jl)
#pragma unroll
          for (int kl = 0; kl < PK; ++kl) {
            const int l = jl * PK + kl;
            if (lmask >> l & 1u) {
              const unsigned ix = cur.ipo + (unsigned)jz[jl] * ipa + (unsigned)kz[kl] * ipb;
              ipl[ix] = FI[l];
              if (ds.out_plane[0]) ds.out_plane[0][ix] = FI[l];
            }
          }
      }
      // tile-boundary faces of the next step
      if (more) {
        if (jb_load) {
#pragma unroll
          for (int kl = 0; kl < PK; ++kl)
            if (lmask >> kl & 1u) JB[kl] = p_ld_cg(jpl + (njp + (unsigned)kz[kl] * jpb));
        }
        if (kb_load) {
#pragma unroll
          for (int jl = 0; jl < PJ; ++jl)
            if (lmask >> (jl * PK) & 1u) KB[jl] = p_ld_cg(kpl + (nkp + (unsigned)jz[jl] * kpb));
        }
      }
      // ---- advance the local schedule ----
      if (last_i) {
        if (POP) {
          if (cur.valid) pop = fma(pop_w[blockIdx.z][cur.d] * (vol_c > 0.0 ? vol_c : 1.0), acc, pop);
          acc = 0.0;
        }
        i = 0; iz = i_first; ++n;
        cur = slice_info(n);
        if (n < cnt_sl) {
          load_coeffs();
          load_iface(cur);
        }
      } else {
        ++i; iz += (unsigned)istep;
      }
    }
    __syncwarp();
    if (lane == 0) p_st_release(cnt + warp, (unsigned)(H + 1));
  }

  if (POP) {  // fixed-order block reduction: lanes, then warps
    pop = warp_sum(pop);
    if (lane == 0) red[warp] = pop;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int w = 0; w < PW; ++w) s += red[w];
      pop_partial[((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = s;
    }
  }
}

#undef ipl
#undef jpl
#undef kpl

}  // namespace kb200

using namespace kb200;

// Returns 0 if the batch was handled, -1 if this path does not apply (caller falls back), >0 on error.
// pop_* (optional): device pointer tables of n entries (+ per-subdomain uniform volume, 0 = read the volume array) and a
// scratch of pop_capacity doubles for the fused population partials; *pop_count receives the number written.
int kb200_sweep_pencil_try(const kb200_sweep_desc *h, int n, const void *d_descs, const double *const *d_pop_w,
                           const double *const *d_pop_vol, const double *d_pop_vol_const, double *d_pop_partial, int pop_capacity,
                           int *pop_count, cudaStream_t st) {
  const int layout = h[0].layout;
  if (pop_count) *pop_count = 0;
  if (layout != 1 && layout != 3 && layout != 4 && layout != 5) return -1;
  const char *env = getenv("KB200_SWEEP_PENCIL");
  if (env && env[0] == '0') return -1;
  PGeom gm;
  gm.layout = layout; gm.Ds = h[0].Ds; gm.Gs = h[0].Gs; gm.ni = h[0].ni; gm.nj = h[0].nj; gm.nk = h[0].nk;
  if ((double)gm.Ds * gm.Gs * gm.ni * gm.nj * gm.nk >= 2147483648.0) return -1;  // 32-bit element offsets
  gm.ntj = (gm.nj + PTJ - 1) / PTJ;
  gm.ntk = (gm.nk + PTK - 1) / PTK;
  const int E = gm.Ds * gm.Gs;
  gm.nslices = (E + 31) / 32;
  const size_t smem = ((size_t)PW * PR * (PK + PJ) * 32 + PW + (size_t)(2 + PJ + PK) * PW * 32) * sizeof(double) + PW * sizeof(unsigned);
  const int sms = sm_count();
  const int ndiag = gm.ntj + gm.ntk - 1;
  // CTAs per (tile, subdomain) of every diagonal: minimise waves x steps per CTA (one CTA per SM)
  std::vector<int> nxs(ndiag);
  long long total_ctas = 0;
  for (int diag = 0; diag < ndiag; ++diag) {
    const int tj_lo = diag - (gm.ntk - 1) > 0 ? diag - (gm.ntk - 1) : 0;
    const int tj_hi = diag < gm.ntj - 1 ? diag : gm.ntj - 1;
    const long long pairs = (long long)(tj_hi - tj_lo + 1) * n;
    int nx = 1;
    long long best = -1;
    for (int c = 1; c <= gm.nslices; ++c) {
      const long long waves = (c * pairs + sms - 1) / sms;
      const long long cost = waves * ((long long)((gm.nslices + c - 1) / c) * gm.ni + 24);
      if (best < 0 || cost < best) { best = cost; nx = c; }
    }
    nxs[diag] = nx;
    total_ctas += nx * pairs;
  }
  double *pp = d_pop_partial;
  if (pp && (!d_pop_w || !d_pop_vol || total_ctas > pop_capacity)) pp = nullptr;
  gm.pop = pp ? 1 : 0;
  const bool exact = exact_mode();
  auto kern = pp ? (exact ? sweep_pencil_kernel<true, true> : sweep_pencil_kernel<false, true>)
                 : (exact ? sweep_pencil_kernel<true, false> : sweep_pencil_kernel<false, false>);
  KB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long used = 0;
  for (int diag = 0; diag < ndiag; ++diag) {
    gm.diag = diag;
    const int tj_lo = diag - (gm.ntk - 1) > 0 ? diag - (gm.ntk - 1) : 0;
    const int tj_hi = diag < gm.ntj - 1 ? diag : gm.ntj - 1;
    const int ntiles = tj_hi - tj_lo + 1;
    dim3 grid(nxs[diag], ntiles, n);
    kern<<<grid, PW * 32, smem, st>>>((const kb200_sweep_desc *)d_descs, gm, d_pop_w, d_pop_vol, d_pop_vol_const, pp ? pp + used : nullptr);
    int rc = post_launch("sweep_pencil");
    if (rc) return rc;
    used += (long long)nxs[diag] * ntiles * n;
  }
  if (pop_count) *pop_count = pp ? (int)used : 0;
  return 0;
}
