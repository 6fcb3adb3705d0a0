/* TEST INFRASTRUCTURE -- CPU oracle (see kripke_oracle.h).  Plain C99, scalar, no FMA
 * (-ffp-contract=off): the reference's CPU build is x86-64 -O3 without -march, hence no FMA
 * either, so every field this file produces is bit-identical to the reference's
 * Sequential path; only the population *reduction* is order-dependent.
 *
 * Each function cites the reference file:line it restates.  Nothing here is used by the product.
 */
#define _USE_MATH_DEFINES
#define _DEFAULT_SOURCE
#include "kripke_oracle.h"
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * Layout rules: src/Kripke/VarTypes.h:73-101 (DefaultOrder<Layout>) + Core/VarLayout.h:86-165.
 * A field's index types are (stably) sorted by their position in DefaultOrder; earlier = slower.
 * ---------------------------------------------------------------------------------------- */
enum { K_DIM, K_MAT, K_DIR, K_LEG, K_MOM, K_GG, K_GRP, K_ZONE, K_ZK, K_ZJ, K_ZI, K_MIX, K_NKINDS };

static int kind_rank(int layout, int kind) {
  static const int order[6][12] = {
      /* DGZ */ {K_DIM, K_MAT, K_DIR, K_LEG, K_MOM, K_GG, K_GRP, K_ZONE, K_ZK, K_ZJ, K_ZI, K_MIX},
      /* DZG */ {K_DIM, K_MAT, K_DIR, K_LEG, K_MOM, K_ZONE, K_ZK, K_ZJ, K_ZI, K_GG, K_GRP, K_MIX},
      /* GDZ */ {K_DIM, K_MAT, K_GG, K_GRP, K_DIR, K_LEG, K_MOM, K_ZONE, K_ZK, K_ZJ, K_ZI, K_MIX},
      /* GZD */ {K_DIM, K_MAT, K_GG, K_GRP, K_ZONE, K_ZK, K_ZJ, K_ZI, K_MIX, K_DIR, K_LEG, K_MOM},
      /* ZDG */ {K_DIM, K_MAT, K_ZONE, K_ZK, K_ZJ, K_ZI, K_MIX, K_DIR, K_LEG, K_MOM, K_GG, K_GRP},
      /* ZGD */ {K_DIM, K_MAT, K_ZONE, K_ZK, K_ZJ, K_ZI, K_MIX, K_GG, K_GRP, K_DIR, K_LEG, K_MOM}};
  for (int r = 0; r < 12; ++r)
    if (order[layout][r] == kind) return r;
  return -1;
}

/* strides[] (in elements) of an n-d field with index kinds[] and extents sizes[] */
static void field_strides(int layout, int n, const int *kinds, const size_t *sizes, size_t *strides) {
  int perm[8];
  for (int i = 0; i < n; ++i) perm[i] = i;
  for (int i = 1; i < n; ++i) { /* stable insertion sort by rank */
    int v = perm[i], j = i - 1;
    while (j >= 0 && kind_rank(layout, kinds[perm[j]]) > kind_rank(layout, kinds[v])) {
      perm[j + 1] = perm[j];
      --j;
    }
    perm[j + 1] = v;
  }
  size_t s = 1;
  for (int i = n - 1; i >= 0; --i) {
    strides[perm[i]] = s;
    s *= sizes[perm[i]];
  }
}

/* ------------------------------------------------------------------------------------------ */
typedef struct {
  double xcos, ycos, zcos, w;
  int id, jd, kd, octant;
} qpoint; /* Generate/Quadrature.cpp:20-33 */

struct ko_problem {
  ko_input in;
  int layout;
  int G, D, M, L1, Z;             /* global sizes */
  int SP, SQ, Sx, Sy, Sz;         /* local subdomain counts (procs folded into zone sets) */
  int Gs, Ds, ni, nj, nk, Zs;     /* per-subdomain extents */
  int nPQR, nPR, nQ, nR;
  /* strides */
  size_t fs_d, fs_g, fs_z;                      /* psi, rhs  <Direction,Group,Zone> */
  size_t ms_m, ms_g, ms_z;                      /* phi, phi_out <Moment,Group,Zone> */
  size_t ip_d, ip_g, ip_j, ip_k;                /* i_plane <Direction,Group,ZoneJ,ZoneK> */
  size_t jp_d, jp_g, jp_i, jp_k;                /* j_plane <Direction,Group,ZoneI,ZoneK> */
  size_t kp_d, kp_g, kp_i, kp_j;                /* k_plane <Direction,Group,ZoneI,ZoneJ> */
  size_t st_g, st_z;                            /* sigt_zonal <Group,Zone> */
  size_t ss_mat, ss_n, ss_g, ss_gp;             /* data/sigs <Material,Legendre,GlobalGroup,GlobalGroup> */
  /* fields: arrays of chunks */
  double **psi, **rhs;                          /* nPQR */
  double **i_plane, **j_plane, **k_plane;       /* nPQR */
  double **phi, **phi_out, **sigt_zonal;        /* nPR */
  double **ell, **ell_plus;                     /* nQ : [d][nm] */
  double **xcos, **ycos, **zcos, **w;           /* nQ */
  int **id, **jd, **kd;                         /* nQ */
  double **dx, **dy, **dz;                      /* Sx, Sy, Sz chunks */
  double **volume;                              /* nR */
  int **zone_to_mixelem, **zone_to_num_mixelem; /* nR */
  int **mixelem_to_zone, **mixelem_to_material; /* nR */
  double **mixelem_to_fraction;                 /* nR */
  int *num_mixelem;                             /* nR */
  double *sigs;                                 /* 1 chunk */
  int *moment_to_legendre;                      /* 1 chunk */
  int *upwind, *downwind;                       /* nPQR x 3 */
  double material_volume[3];
};

/* subdomain id <-> coords: Core/PartitionSpace.cpp:75-98,202-217 (row-major P,Q,Rx,Ry,Rz) */
static int sdom_id(const ko_problem *p, int gs, int ds, int x, int y, int z) {
  return (((gs * p->SQ + ds) * p->Sx + x) * p->Sy + y) * p->Sz + z;
}
static void sdom_coord(const ko_problem *p, int s, int c[5]) {
  c[4] = s % p->Sz; s /= p->Sz;
  c[3] = s % p->Sy; s /= p->Sy;
  c[2] = s % p->Sx; s /= p->Sx;
  c[1] = s % p->SQ; s /= p->SQ;
  c[0] = s;
}
/* projections PQR -> PR, Q, R chunk ids: Core/PartitionSpace.cpp:246-258 */
static int chunk_pr(const ko_problem *p, int s) {
  int c[5]; sdom_coord(p, s, c);
  return ((c[0] * p->Sx + c[2]) * p->Sy + c[3]) * p->Sz + c[4];
}
static int chunk_q(const ko_problem *p, int s) { int c[5]; sdom_coord(p, s, c); return c[1]; }
static int chunk_r(const ko_problem *p, int s) {
  int c[5]; sdom_coord(p, s, c);
  return (c[2] * p->Sy + c[3]) * p->Sz + c[4];
}

/* ---------------- defaults / validation ---------------- */
void ko_default_input(ko_input *in) { /* InputVariables.cpp:17-47 */
  memset(in, 0, sizeof(*in));
  in->nx = in->ny = in->nz = 16;
  in->num_directions = 96;
  in->num_groups = 32;
  in->legendre_order = 4;
  in->layout = KO_DGZ;
  in->npx = in->npy = in->npz = 1;
  in->num_dirsets = 8;
  in->num_groupsets = 2;
  in->num_zonesets_dim[0] = in->num_zonesets_dim[1] = in->num_zonesets_dim[2] = 1;
  in->sigt[0] = 0.1; in->sigt[1] = 0.0001; in->sigt[2] = 0.1;
  in->sigs[0] = 0.05; in->sigs[1] = 0.00005; in->sigs[2] = 0.05;
  in->num_material_subsamples = 4;
}

int ko_check_input(const ko_input *in) { /* InputVariables.cpp:52-109 */
  if (in->num_zonesets_dim[0] <= 0 || in->num_zonesets_dim[1] <= 0 || in->num_zonesets_dim[2] <= 0) return 1;
  if (in->num_groups < 1) return 2;
  if (in->num_groups % in->num_groupsets) return 3;
  if (in->num_directions < 8) return 4;
  if (in->num_dirsets % 8 && in->num_dirsets < 8) return 5;
  if (in->num_directions % in->num_dirsets) return 6;
  if (in->legendre_order < 0) return 7;
  return 0;
}

/* ---------------- quadrature: Generate/Quadrature.cpp ---------------- */
static void gauss_legendre(double x1, double x2, double *x, double *w, int n, double eps) { /* :39-71 */
  int m = (n + 1) / 2;
  double xm = 0.5 * (x2 + x1), xl = 0.5 * (x2 - x1);
  for (int i = 1; i <= m; i++) {
    double z = cos(M_PI * (i - 0.25) / (n + 0.5)), z1, pp, p1, p2, p3;
    do {
      p1 = 1.0; p2 = 0.0;
      for (int j = 1; j <= n; j++) {
        p3 = p2; p2 = p1;
        p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j;
      }
      pp = n * (z * p1 - p2) / (z * z - 1.0);
      z1 = z;
      z = z1 - p1 / pp;
    } while (fabs(z - z1) > eps);
    x[i - 1] = xm - xl * z;
    x[n - i] = xm + xl * z;
    w[i - 1] = 2.0 * xl / ((1.0 - z * z) * pp * pp);
    w[n - i] = w[i - 1];
  }
}

static double fact_fcn(int n) { /* :78-85 */
  double fact = 1.0;
  for (int i = n; i > 0; --i) fact *= (double)i;
  return fact;
}

static double pnm_fcn(int n, int m, double x) { /* :87-138 */
  double fact, pnn = 0, pmm, pmmp1, somx2;
  pmm = 1.0;
  if (m > 0) {
    somx2 = sqrt((1.0 - x) * (1.0 + x));
    fact = 1.0;
    for (int i = 1; i <= m; i++) {
      pmm *= -fact * somx2;
      fact += 2.0;
    }
  }
  if (n == m) return pmm;
  pmmp1 = x * (2 * m + 1) * pmm;
  if (n == (m + 1)) return pmmp1;
  for (int nn = m + 2; nn <= n; nn++) {
    pnn = (x * (2 * nn - 1) * pmmp1 - (nn + m - 1) * pmm) / (nn - m);
    pmm = pmmp1;
    pmmp1 = pnn;
  }
  return pnn;
}

static double ynm_fcn(int n, int m, double mu, double eta, double xi) { /* :140-183 */
  double fac1, fac2, anm, pnm, dm0, taum, tmp, phi, phi_tmp;
  double floor_ = 1.e-20;
  tmp = fabs(eta / (mu + floor_));
  phi_tmp = atan(tmp);
  if ((mu > 0) && (eta > 0)) phi = phi_tmp;
  else if ((mu < 0) && (eta > 0)) phi = M_PI - fabs(phi_tmp);
  else if ((mu < 0) && (eta < 0)) phi = M_PI + fabs(phi_tmp);
  else phi = 2.0 * M_PI - fabs(phi_tmp);
  fac1 = fact_fcn(n - abs(m));
  fac2 = fact_fcn(n + abs(m));
  pnm = pnm_fcn(n, abs(m), xi);
  tmp = ((double)m) * phi;
  if (m >= 0) taum = cos(tmp);
  else taum = sin(-tmp);
  dm0 = (m == 0) ? 1.0 : 0.0;
  tmp = ((2 * n + 1) * fac1) / (2.0 * (1.0 + dm0) * M_PI * fac2);
  anm = sqrt(tmp);
  return anm * pnm * taum;
}

/* Directions are sorted by DESCENDING octant with std::sort (:74-76,278).  libstdc++'s std::sort
 * on <= 16 elements is a plain (stable) insertion sort; above that it is an introsort whose
 * order among equal keys is implementation-defined, which this oracle does not restate:
 * Gauss-Legendre sets with more than 16 points are checked against golden vectors only. */
static int create_quadrature(const ko_input *in, qpoint *dirs) { /* :192-309 */
  int nd = in->num_directions;
  memset(dirs, 0, sizeof(qpoint) * (size_t)nd);
  if (in->quad_num_polar > 0) {
    int np = in->quad_num_polar, na = in->quad_num_azimuthal;
    if (np % 4 || na % 2 || np * na != nd) return 1;
    if (nd > 16) return 2;
    double *pc = malloc(sizeof(double) * np), *pw = malloc(sizeof(double) * np);
    double *aa = malloc(sizeof(double) * na), *aw = malloc(sizeof(double) * na);
    gauss_legendre(-1.0, 1.0, pc, pw, np, DBL_EPSILON);
    double dangle = 2.0 * M_PI / ((double)na);
    for (int i = 0; i < na; i++) {
      aa[i] = (i == 0) ? dangle / 2.0 : aa[i - 1] + dangle;
      aw[i] = dangle;
    }
    int d = 0;
    for (int i = 0; i < np; i++)
      for (int j = 0; j < na; j++, d++) {
        double xcos = sqrt(1.0 - pc[i] * pc[i]) * cos(aa[j]);
        double ycos = sqrt(1.0 - pc[i] * pc[i]) * sin(aa[j]);
        double zcos = pc[i];
        dirs[d].id = (xcos > 0.) ? 1 : -1;
        dirs[d].jd = (ycos > 0.) ? 1 : -1;
        dirs[d].kd = (zcos > 0.) ? 1 : -1;
        dirs[d].octant = (dirs[d].id == -1 ? 1 : 0) + (dirs[d].jd == -1 ? 2 : 0) + (dirs[d].kd == -1 ? 4 : 0);
        dirs[d].xcos = fabs(xcos);
        dirs[d].ycos = fabs(ycos);
        dirs[d].zcos = fabs(zcos);
        dirs[d].w = pw[i] * aw[j];
      }
    for (int i = 1; i < nd; ++i) { /* stable insertion sort, comparator b.octant < a.octant */
      qpoint v = dirs[i];
      int j = i - 1;
      while (j >= 0 && dirs[j].octant < v.octant) { dirs[j + 1] = dirs[j]; --j; }
      dirs[j + 1] = v;
    }
    free(pc); free(pw); free(aa); free(aw);
  } else { /* dummy S2 with repeated directions :280-306 */
    double mu = cos(M_PI / 4);
    double eta = sqrt(1 - mu * mu) * cos(M_PI / 4);
    double xi = sqrt(1 - mu * mu) * sin(M_PI / 4);
    int per = nd / 8, d = 0;
    for (int octant = 0; octant < 8; ++octant)
      for (int sd = 0; sd < per; sd++, d++) {
        dirs[d].id = (octant & 1) ? 1 : -1;
        dirs[d].jd = ((octant >> 1) & 1) ? 1 : -1;
        dirs[d].kd = ((octant >> 2) & 1) ? 1 : -1;
        dirs[d].w = 4.0 * M_PI / (double)nd;
        dirs[d].xcos = mu;
        dirs[d].ycos = eta;
        dirs[d].zcos = xi;
      }
  }
  return 0;
}

/* material_fcn: Generate/Space.cpp:142-177 */
static int material_fcn(double x, double y, double z) {
  x = fabs(x); y = fabs(y); z = fabs(z);
  if (x <= 10.0 && y <= 10.0 && z <= 10.0) return 0;
  if (x <= 10.0 && y <= 60.0 && z <= 10.0) return 1;
  if (x <= 40.0 && y >= 50.0 && y <= 60.0 && z <= 10.0) return 1;
  if (x >= 30.0 && x <= 40.0 && y >= 50.0 && y <= 60.0 && z <= 40.0) return 1;
  if (x >= 30.0 && x <= 40.0 && y >= 50.0 && z >= 30.0 && z <= 40.0) return 1;
  return 2;
}

static void *xcalloc(size_t n, size_t sz) {
  void *p = calloc(n ? n : 1, sz);
  if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
  return p;
}
static double **alloc_chunks(int nchunks, size_t n) {
  double **c = xcalloc((size_t)nchunks, sizeof(double *));
  for (int i = 0; i < nchunks; ++i) c[i] = xcalloc(n, sizeof(double));
  return c;
}
static int **alloc_ichunks(int nchunks, size_t n) {
  int **c = xcalloc((size_t)nchunks, sizeof(int *));
  for (int i = 0; i < nchunks; ++i) c[i] = xcalloc(n, sizeof(int));
  return c;
}

ko_problem *ko_create(const ko_input *in) {
  if (ko_check_input(in)) return NULL;
  ko_problem *p = xcalloc(1, sizeof(*p));
  p->in = *in;
  p->layout = in->layout;
  p->G = in->num_groups;
  p->D = in->num_directions;
  p->L1 = in->legendre_order + 1;
  p->M = p->L1 * p->L1;
  p->Z = in->nx * in->ny * in->nz;
  /* Generate/Decomp.cpp:23-56; --procs folded into the zone sets */
  p->SP = in->num_groupsets;
  p->SQ = in->num_dirsets;
  p->Sx = in->num_zonesets_dim[0] * in->npx;
  p->Sy = in->num_zonesets_dim[1] * in->npy;
  p->Sz = in->num_zonesets_dim[2] * in->npz;
  p->Gs = p->G / p->SP;                       /* Generate/Energy.cpp:29-33 */
  p->Ds = p->D / p->SQ;                       /* Generate/Quadrature.cpp:323-327 */
  p->ni = in->nx / p->Sx;                     /* Generate/Space.cpp:46-87 */
  p->nj = in->ny / p->Sy;
  p->nk = in->nz / p->Sz;
  if (p->ni * p->Sx != in->nx || p->nj * p->Sy != in->ny || p->nk * p->Sz != in->nz) { free(p); return NULL; }
  p->Zs = p->ni * p->nj * p->nk;
  p->nPQR = p->SP * p->SQ * p->Sx * p->Sy * p->Sz;
  p->nPR = p->SP * p->Sx * p->Sy * p->Sz;
  p->nQ = p->SQ;
  p->nR = p->Sx * p->Sy * p->Sz;
  int L = p->layout;

  { int k[3] = {K_DIR, K_GRP, K_ZONE}; size_t s[3] = {p->Ds, p->Gs, p->Zs}, st[3];
    field_strides(L, 3, k, s, st); p->fs_d = st[0]; p->fs_g = st[1]; p->fs_z = st[2]; }
  { int k[3] = {K_MOM, K_GRP, K_ZONE}; size_t s[3] = {p->M, p->Gs, p->Zs}, st[3];
    field_strides(L, 3, k, s, st); p->ms_m = st[0]; p->ms_g = st[1]; p->ms_z = st[2]; }
  { int k[4] = {K_DIR, K_GRP, K_ZJ, K_ZK}; size_t s[4] = {p->Ds, p->Gs, p->nj, p->nk}, st[4];
    field_strides(L, 4, k, s, st); p->ip_d = st[0]; p->ip_g = st[1]; p->ip_j = st[2]; p->ip_k = st[3]; }
  { int k[4] = {K_DIR, K_GRP, K_ZI, K_ZK}; size_t s[4] = {p->Ds, p->Gs, p->ni, p->nk}, st[4];
    field_strides(L, 4, k, s, st); p->jp_d = st[0]; p->jp_g = st[1]; p->jp_i = st[2]; p->jp_k = st[3]; }
  { int k[4] = {K_DIR, K_GRP, K_ZI, K_ZJ}; size_t s[4] = {p->Ds, p->Gs, p->ni, p->nj}, st[4];
    field_strides(L, 4, k, s, st); p->kp_d = st[0]; p->kp_g = st[1]; p->kp_i = st[2]; p->kp_j = st[3]; }
  { int k[2] = {K_GRP, K_ZONE}; size_t s[2] = {p->Gs, p->Zs}, st[2];
    field_strides(L, 2, k, s, st); p->st_g = st[0]; p->st_z = st[1]; }
  { int k[4] = {K_MAT, K_LEG, K_GG, K_GG}; size_t s[4] = {3, p->L1, p->G, p->G}, st[4];
    field_strides(L, 4, k, s, st); p->ss_mat = st[0]; p->ss_n = st[1]; p->ss_g = st[2]; p->ss_gp = st[3]; }

  /* ---- quadrature, ell, ell_plus, adjacency: Generate/Quadrature.cpp:315-519 ---- */
  qpoint *dirs = xcalloc((size_t)p->D, sizeof(qpoint));
  if (create_quadrature(in, dirs)) { free(dirs); free(p); return NULL; }
  p->moment_to_legendre = xcalloc((size_t)p->M, sizeof(int));
  { int nm = 0;
    for (int n = 0; n < p->L1; ++n)
      for (int m = -n; m <= n; ++m) p->moment_to_legendre[nm++] = n; }
  p->xcos = alloc_chunks(p->nQ, p->Ds); p->ycos = alloc_chunks(p->nQ, p->Ds);
  p->zcos = alloc_chunks(p->nQ, p->Ds); p->w = alloc_chunks(p->nQ, p->Ds);
  p->id = alloc_ichunks(p->nQ, p->Ds); p->jd = alloc_ichunks(p->nQ, p->Ds); p->kd = alloc_ichunks(p->nQ, p->Ds);
  p->ell = alloc_chunks(p->nQ, (size_t)p->Ds * p->M);
  p->ell_plus = alloc_chunks(p->nQ, (size_t)p->Ds * p->M);
  double SQRT4PI = sqrt(4 * M_PI);
  for (int q = 0; q < p->nQ; ++q) {
    int lower = q * p->Ds;
    for (int d = 0; d < p->Ds; ++d) {
      qpoint *pt = &dirs[d + lower];
      p->xcos[q][d] = pt->xcos; p->ycos[q][d] = pt->ycos; p->zcos[q][d] = pt->zcos; p->w[q][d] = pt->w;
      p->id[q][d] = pt->id; p->jd[q][d] = pt->jd; p->kd[q][d] = pt->kd;
    }
    int nm = 0;
    for (int n = 0; n < p->L1; n++)
      for (int m = -n; m <= n; m++, nm++)
        for (int d = 0; d < p->Ds; ++d) {
          qpoint *pt = &dirs[d + lower];
          double xc = (pt->id) * (pt->xcos), yc = (pt->jd) * (pt->ycos), zc = (pt->kd) * (pt->zcos);
          double ynm = ynm_fcn(n, m, xc, yc, zc);
          /* ell<Moment,Direction> and ell_plus<Direction,Moment> are both stored [d][nm] */
          p->ell[q][(size_t)d * p->M + nm] = pt->w * ynm / SQRT4PI;
          p->ell_plus[q][(size_t)d * p->M + nm] = ynm * SQRT4PI;
        }
  }
  free(dirs);
  p->upwind = xcalloc((size_t)p->nPQR * 3, sizeof(int));
  p->downwind = xcalloc((size_t)p->nPQR * 3, sizeof(int));
  for (int s = 0; s < p->nPQR; ++s) { /* :466-518 */
    int c[5]; sdom_coord(p, s, c);
    int q = c[1];
    int sweep_dir[3] = {p->id[q][0], p->jd[q][0], p->kd[q][0]};
    int nsd[3] = {p->Sx, p->Sy, p->Sz};
    for (int dim = 0; dim < 3; ++dim) {
      int up[5], dn[5];
      memcpy(up, c, sizeof(up)); memcpy(dn, c, sizeof(dn));
      up[2 + dim] -= sweep_dir[dim];
      dn[2 + dim] += sweep_dir[dim];
      p->upwind[s * 3 + dim] = (up[2 + dim] < 0 || up[2 + dim] >= nsd[dim]) ? -1 : sdom_id(p, up[0], up[1], up[2], up[3], up[4]);
      p->downwind[s * 3 + dim] = (dn[2 + dim] < 0 || dn[2 + dim] >= nsd[dim]) ? -1 : sdom_id(p, dn[0], dn[1], dn[2], dn[3], dn[4]);
    }
  }

  /* ---- mesh + materials: Generate/Space.cpp:38-365 ---- */
  double x_min = -60.0, x_max = 60.0, y_min = -100.0, y_max = 100.0, z_min = -60.0, z_max = 60.0;
  double dx = (x_max - x_min) / in->nx, dy = (y_max - y_min) / in->ny, dz = (z_max - z_min) / in->nz;
  p->dx = alloc_chunks(p->Sx, p->ni); p->dy = alloc_chunks(p->Sy, p->nj); p->dz = alloc_chunks(p->Sz, p->nk);
  for (int c = 0; c < p->Sx; ++c) for (int i = 0; i < p->ni; ++i) p->dx[c][i] = dx;
  for (int c = 0; c < p->Sy; ++c) for (int i = 0; i < p->nj; ++i) p->dy[c][i] = dy;
  for (int c = 0; c < p->Sz; ++c) for (int i = 0; i < p->nk; ++i) p->dz[c][i] = dz;
  double zone_volume = dx * dy * dz;
  p->volume = alloc_chunks(p->nR, p->Zs);
  p->zone_to_mixelem = alloc_ichunks(p->nR, p->Zs);
  p->zone_to_num_mixelem = alloc_ichunks(p->nR, p->Zs);
  p->mixelem_to_zone = xcalloc((size_t)p->nR, sizeof(int *));
  p->mixelem_to_material = xcalloc((size_t)p->nR, sizeof(int *));
  p->mixelem_to_fraction = xcalloc((size_t)p->nR, sizeof(double *));
  p->num_mixelem = xcalloc((size_t)p->nR, sizeof(int));
  int ns = in->num_material_subsamples;
  double sample_vol_frac = 1.0 / (double)(ns * ns * ns);
  double total_volume[3] = {0, 0, 0};
  for (int r = 0; r < p->nR; ++r) {
    int rz = r % p->Sz, ry = (r / p->Sz) % p->Sy, rx = r / (p->Sz * p->Sy);
    double x0 = x_min + dx * (rx * p->ni), y0 = y_min + dy * (ry * p->nj), z0 = z_min + dz * (rz * p->nk);
    double(*mix)[3] = xcalloc((size_t)p->Zs, sizeof(double[3]));
    for (int k = 0; k < p->nk; k++)
      for (int j = 0; j < p->nj; j++)
        for (int i = 0; i < p->ni; i++) {
          int zone = (k * p->nj + j) * p->ni + i; /* ProductSet<3>(k,j,i) :92-93 */
          double xi = x0 + dx * i, yi = y0 + dy * j, zi = z0 + dz * k;
          for (int si = 0; si < ns; ++si)
            for (int sj = 0; sj < ns; ++sj)
              for (int sk = 0; sk < ns; ++sk) {
                double x = xi + dx * (si + 1) / (ns + 1);
                double y = yi + dy * (sj + 1) / (ns + 1);
                double z = zi + dz * (sk + 1) / (ns + 1);
                mix[zone][material_fcn(x, y, z)] += sample_vol_frac;
              }
        }
    int n = 0;
    for (int z = 0; z < p->Zs; ++z) n += (mix[z][0] > 0.0) + (mix[z][1] > 0.0) + (mix[z][2] > 0.0);
    p->num_mixelem[r] = n;
    p->mixelem_to_zone[r] = xcalloc((size_t)n, sizeof(int));
    p->mixelem_to_material[r] = xcalloc((size_t)n, sizeof(int));
    p->mixelem_to_fraction[r] = xcalloc((size_t)n, sizeof(double));
    int me = 0;
    for (int z = 0; z < p->Zs; ++z) {
      p->volume[r][z] = zone_volume;
      p->zone_to_mixelem[r][z] = me;
      int cnt = 0;
      for (int m = 0; m < 3; ++m)
        if (mix[z][m] > 0.0) {
          p->mixelem_to_zone[r][me] = z;
          p->mixelem_to_material[r][me] = m;
          p->mixelem_to_fraction[r][me] = mix[z][m];
          total_volume[m] += mix[z][m] * zone_volume;
          ++cnt; ++me;
        }
      p->zone_to_num_mixelem[r][z] = cnt;
    }
    free(mix);
  }
  memcpy(p->material_volume, total_volume, sizeof(total_volume));
  /* sigt_zonal :336-363 */
  p->sigt_zonal = alloc_chunks(p->nPR, (size_t)p->Gs * p->Zs);
  for (int pr = 0; pr < p->nPR; ++pr) {
    int r = pr % p->nR;
    for (int g = 0; g < p->Gs; ++g)
      for (int me = 0; me < p->num_mixelem[r]; ++me) {
        int z = p->mixelem_to_zone[r][me], mat = p->mixelem_to_material[r][me];
        p->sigt_zonal[pr][g * p->st_g + z * p->st_z] += p->mixelem_to_fraction[r][me] * in->sigt[mat];
      }
  }

  /* ---- fields: Generate/Data.cpp:22-105 ---- */
  size_t nflux = (size_t)p->Ds * p->Gs * p->Zs, nmom = (size_t)p->M * p->Gs * p->Zs;
  p->psi = alloc_chunks(p->nPQR, nflux);
  p->rhs = alloc_chunks(p->nPQR, nflux);
  p->phi = alloc_chunks(p->nPR, nmom);
  p->phi_out = alloc_chunks(p->nPR, nmom);
  p->i_plane = alloc_chunks(p->nPQR, (size_t)p->Ds * p->Gs * p->nj * p->nk);
  p->j_plane = alloc_chunks(p->nPQR, (size_t)p->Ds * p->Gs * p->ni * p->nk);
  p->k_plane = alloc_chunks(p->nPQR, (size_t)p->Ds * p->Gs * p->ni * p->nj);
  p->sigs = xcalloc((size_t)3 * p->L1 * p->G * p->G, sizeof(double));
  for (int mat = 0; mat < 3; ++mat)
    for (int g = 0; g < p->G; ++g)
      p->sigs[mat * p->ss_mat + 0 * p->ss_n + g * p->ss_g + g * p->ss_gp] = in->sigs[mat];
  return p;
}

static void free_chunks(void **c, int n) {
  if (!c) return;
  for (int i = 0; i < n; ++i) free(c[i]);
  free(c);
}
void ko_destroy(ko_problem *p) {
  if (!p) return;
  free_chunks((void **)p->psi, p->nPQR); free_chunks((void **)p->rhs, p->nPQR);
  free_chunks((void **)p->i_plane, p->nPQR); free_chunks((void **)p->j_plane, p->nPQR);
  free_chunks((void **)p->k_plane, p->nPQR);
  free_chunks((void **)p->phi, p->nPR); free_chunks((void **)p->phi_out, p->nPR);
  free_chunks((void **)p->sigt_zonal, p->nPR);
  free_chunks((void **)p->ell, p->nQ); free_chunks((void **)p->ell_plus, p->nQ);
  free_chunks((void **)p->xcos, p->nQ); free_chunks((void **)p->ycos, p->nQ);
  free_chunks((void **)p->zcos, p->nQ); free_chunks((void **)p->w, p->nQ);
  free_chunks((void **)p->id, p->nQ); free_chunks((void **)p->jd, p->nQ); free_chunks((void **)p->kd, p->nQ);
  free_chunks((void **)p->dx, p->Sx); free_chunks((void **)p->dy, p->Sy); free_chunks((void **)p->dz, p->Sz);
  free_chunks((void **)p->volume, p->nR);
  free_chunks((void **)p->zone_to_mixelem, p->nR); free_chunks((void **)p->zone_to_num_mixelem, p->nR);
  free_chunks((void **)p->mixelem_to_zone, p->nR); free_chunks((void **)p->mixelem_to_material, p->nR);
  free_chunks((void **)p->mixelem_to_fraction, p->nR);
  free(p->num_mixelem); free(p->sigs); free(p->moment_to_legendre); free(p->upwind); free(p->downwind);
  free(p);
}

/* ---------------- kernels ---------------- */
void ko_ltimes(ko_problem *p) { /* Kernel/LTimes.cpp:54-65,95-103: ascending sdom, then ascending d */
  for (int s = 0; s < p->nPQR; ++s) {
    const double *psi = p->psi[s], *ell = p->ell[chunk_q(p, s)];
    double *phi = p->phi[chunk_pr(p, s)];
    for (int nm = 0; nm < p->M; ++nm)
      for (int d = 0; d < p->Ds; ++d) {
        double e = ell[(size_t)d * p->M + nm];
        for (int g = 0; g < p->Gs; ++g)
          for (int z = 0; z < p->Zs; ++z)
            phi[nm * p->ms_m + g * p->ms_g + z * p->ms_z] += e * psi[d * p->fs_d + g * p->fs_g + z * p->fs_z];
      }
  }
}

void ko_lplustimes(ko_problem *p) { /* Kernel/LPlusTimes.cpp:49-60,84-90: ascending nm */
  for (int s = 0; s < p->nPQR; ++s) {
    double *rhs = p->rhs[s];
    const double *ellp = p->ell_plus[chunk_q(p, s)], *phi_out = p->phi_out[chunk_pr(p, s)];
    for (int d = 0; d < p->Ds; ++d)
      for (int nm = 0; nm < p->M; ++nm) {
        double e = ellp[(size_t)d * p->M + nm];
        for (int g = 0; g < p->Gs; ++g)
          for (int z = 0; z < p->Zs; ++z)
            rhs[d * p->fs_d + g * p->fs_g + z * p->fs_z] += e * phi_out[nm * p->ms_m + g * p->ms_g + z * p->ms_z];
      }
  }
}

void ko_scattering(ko_problem *p) { /* Kernel/Scattering.cpp:73-99,137-161 */
  for (int src = 0; src < p->nPR; ++src)
    for (int dst = 0; dst < p->nPR; ++dst) {
      if (src % p->nR != dst % p->nR) continue; /* same R subdomain only :141-145 */
      int r = src % p->nR;
      int glower_src = (src / p->nR) * p->Gs, glower_dst = (dst / p->nR) * p->Gs;
      const double *phi = p->phi[src];
      double *phi_out = p->phi_out[dst];
      for (int nm = 0; nm < p->M; ++nm) {
        int n = p->moment_to_legendre[nm];
        for (int g = 0; g < p->Gs; ++g)
          for (int gp = 0; gp < p->Gs; ++gp)
            for (int z = 0; z < p->Zs; ++z) {
              int m0 = p->zone_to_mixelem[r][z], m1 = m0 + p->zone_to_num_mixelem[r][z];
              double sigs_z = 0.0;
              for (int mix = m0; mix < m1; ++mix)
                sigs_z += p->sigs[p->mixelem_to_material[r][mix] * p->ss_mat + n * p->ss_n +
                                  (g + glower_dst) * p->ss_g + (gp + glower_src) * p->ss_gp] *
                          p->mixelem_to_fraction[r][mix];
              phi_out[nm * p->ms_m + g * p->ms_g + z * p->ms_z] += sigs_z * phi[nm * p->ms_m + gp * p->ms_g + z * p->ms_z];
            }
      }
    }
}

void ko_source(ko_problem *p) { /* Kernel/Source.cpp:59-75,98-112 ; strength 1.0, material 0 */
  double source_strength = 1.0;
  for (int pr = 0; pr < p->nPR; ++pr) {
    int r = pr % p->nR;
    for (int g = 0; g < p->Gs; ++g)
      for (int mix = 0; mix < p->num_mixelem[r]; ++mix)
        if (p->mixelem_to_material[r][mix] == 0) {
          int z = p->mixelem_to_zone[r][mix];
          p->phi_out[pr][0 * p->ms_m + g * p->ms_g + z * p->ms_z] += source_strength * p->mixelem_to_fraction[r][mix];
        }
  }
}

void ko_sweep_subdomain(ko_problem *p, int s) { /* Kernel/SweepSubdomain.cpp:34-109 */
  int c[5]; sdom_coord(p, s, c);
  int q = c[1], pr = chunk_pr(p, s);
  const double *xcos = p->xcos[q], *ycos = p->ycos[q], *zcos = p->zcos[q];
  const double *dx = p->dx[c[2]], *dy = p->dy[c[3]], *dz = p->dz[c[4]];
  const double *sigt = p->sigt_zonal[pr], *rhs = p->rhs[s];
  double *psi = p->psi[s], *lf = p->i_plane[s], *fr = p->j_plane[s], *bo = p->k_plane[s];
  int id = p->id[q][0], jd = p->jd[q][0], kd = p->kd[q][0]; /* direction 0 of the set :61-72 */
  int si = (id > 0) ? 0 : p->ni - 1, sj = (jd > 0) ? 0 : p->nj - 1, sk = (kd > 0) ? 0 : p->nk - 1;
  int ei = (id > 0) ? p->ni : -1, ej = (jd > 0) ? p->nj : -1, ek = (kd > 0) ? p->nk : -1;
  for (int d = 0; d < p->Ds; ++d)
    for (int g = 0; g < p->Gs; ++g)
      for (int k = sk; k != ek; k += kd)
        for (int j = sj; j != ej; j += jd)
          for (int i = si; i != ei; i += id) {
            double xcos_dxi = 2.0 * xcos[d] / dx[i];
            double ycos_dyj = 2.0 * ycos[d] / dy[j];
            double zcos_dzk = 2.0 * zcos[d] / dz[k];
            size_t z = (size_t)(k * p->nj + j) * p->ni + i;
            size_t ilf = d * p->ip_d + g * p->ip_g + j * p->ip_j + k * p->ip_k;
            size_t ifr = d * p->jp_d + g * p->jp_g + i * p->jp_i + k * p->jp_k;
            size_t ibo = d * p->kp_d + g * p->kp_g + i * p->kp_i + j * p->kp_j;
            size_t iz = d * p->fs_d + g * p->fs_g + z * p->fs_z;
            double psi_d_g_z = (rhs[iz] + lf[ilf] * xcos_dxi + fr[ifr] * ycos_dyj + bo[ibo] * zcos_dzk) /
                               (xcos_dxi + ycos_dyj + zcos_dzk + sigt[g * p->st_g + z * p->st_z]);
            psi[iz] = psi_d_g_z;
            lf[ilf] = 2.0 * psi_d_g_z - lf[ilf];
            fr[ifr] = 2.0 * psi_d_g_z - fr[ifr];
            bo[ibo] = 2.0 * psi_d_g_z - bo[ibo];
          }
}

static size_t plane_size(const ko_problem *p, int dim) {
  size_t e = (size_t)p->Ds * p->Gs;
  return dim == 0 ? e * p->nj * p->nk : dim == 1 ? e * p->ni * p->nk : e * p->ni * p->nj;
}
static double **plane_field(const ko_problem *p, int dim) {
  return dim == 0 ? p->i_plane : dim == 1 ? p->j_plane : p->k_plane;
}

/* SweepSolver.cpp:23-87 with SweepComm (ParallelComm.cpp:61-185, ParallelComm/SweepComm.cpp) or the
 * de-facto BlockJacobiComm semantics (SURVEY section 5 defects 1-3: zero inflow on every face). */
static void sweep_solver_impl(ko_problem *p, int bj, int *order_out, int do_compute) {
  int n = p->nPQR;
  if (do_compute)
    for (int dim = 0; dim < 3; ++dim)
      for (int s = 0; s < n; ++s) memset(plane_field(p, dim)[s], 0, plane_size(p, dim) * sizeof(double));
  int *queue = malloc(sizeof(int) * n), *depends = malloc(sizeof(int) * n);
  int qn = 0, nvisited = 0;
  for (int s = 0; s < n; ++s) { /* addSubdomain -> postRecvs: count on-rank upwind deps */
    int nd = 0;
    for (int dim = 0; dim < 3; ++dim)
      if (p->upwind[s * 3 + dim] >= 0) nd++;
    queue[qn] = s;
    depends[qn] = bj ? 0 : nd; /* BJ: all sends (of zeros) are posted before any sweep */
    qn++;
  }
  while (qn > 0) {
    int idx = -1;
    for (int i = 0; i < qn; ++i)
      if (depends[i] == 0) { idx = i; break; } /* first ready in queue order :58-59 */
    if (idx < 0) { fprintf(stderr, "oracle: sweep deadlock\n"); abort(); }
    int s = queue[idx];
    if (order_out) order_out[nvisited] = s;
    nvisited++;
    if (do_compute) {
      for (int dim = 0; dim < 3; ++dim) /* vacuum boundary :64-72 */
        if (p->upwind[s * 3 + dim] == -1) memset(plane_field(p, dim)[s], 0, plane_size(p, dim) * sizeof(double));
      ko_sweep_subdomain(p, s);
    }
    /* markComplete: dequeue, then postSends (not for BJ: BlockJacobiComm::markComplete only dequeues) */
    memmove(queue + idx, queue + idx + 1, sizeof(int) * (qn - idx - 1));
    memmove(depends + idx, depends + idx + 1, sizeof(int) * (qn - idx - 1));
    qn--;
    if (!bj)
      for (int dim = 0; dim < 3; ++dim) {
        int dw = p->downwind[s * 3 + dim];
        if (dw < 0) continue;
        for (int i = 0; i < qn; ++i)
          if (queue[i] == dw) { depends[i]--; break; }
        if (do_compute) memcpy(plane_field(p, dim)[dw], plane_field(p, dim)[s], plane_size(p, dim) * sizeof(double));
      }
  }
  free(queue); free(depends);
}

void ko_sweep_solver(ko_problem *p, int bj) { sweep_solver_impl(p, bj, NULL, 1); }
void ko_sweep_order(const ko_problem *p, int *order) { sweep_solver_impl((ko_problem *)p, 0, order, 0); }

double ko_population(ko_problem *p) { /* Kernel/Population.cpp:49-63,90-100 ; Sequential = memory order */
  double part = 0.0;
  for (int s = 0; s < p->nPQR; ++s) {
    const double *psi = p->psi[s], *w = p->w[chunk_q(p, s)], *vol = p->volume[chunk_r(p, s)];
    /* loop nest in storage order (Arch/Population.h Sequential policies) */
    int ext[3] = {p->Ds, p->Gs, p->Zs};
    size_t str[3] = {p->fs_d, p->fs_g, p->fs_z};
    int o[3] = {0, 1, 2};
    for (int a = 0; a < 3; ++a)
      for (int b = a + 1; b < 3; ++b)
        if (str[o[b]] > str[o[a]]) { int t = o[a]; o[a] = o[b]; o[b] = t; }
    double part_red = 0.0;
    int idx[3];
    for (idx[o[0]] = 0; idx[o[0]] < ext[o[0]]; ++idx[o[0]])
      for (idx[o[1]] = 0; idx[o[1]] < ext[o[1]]; ++idx[o[1]])
        for (idx[o[2]] = 0; idx[o[2]] < ext[o[2]]; ++idx[o[2]])
          part_red += w[idx[0]] * psi[idx[0] * p->fs_d + idx[1] * p->fs_g + idx[2] * p->fs_z] * vol[idx[2]];
    part += part_red;
  }
  return part;
}

static void zero_chunks(double **c, int n, size_t len) {
  for (int i = 0; i < n; ++i) memset(c[i], 0, len * sizeof(double));
}
void ko_zero(ko_problem *p, const char *f) {
  size_t nflux = (size_t)p->Ds * p->Gs * p->Zs, nmom = (size_t)p->M * p->Gs * p->Zs;
  if (!strcmp(f, "psi")) zero_chunks(p->psi, p->nPQR, nflux);
  else if (!strcmp(f, "rhs")) zero_chunks(p->rhs, p->nPQR, nflux);
  else if (!strcmp(f, "phi")) zero_chunks(p->phi, p->nPR, nmom);
  else if (!strcmp(f, "phi_out")) zero_chunks(p->phi_out, p->nPR, nmom);
  else if (!strcmp(f, "i_plane")) zero_chunks(p->i_plane, p->nPQR, plane_size(p, 0));
  else if (!strcmp(f, "j_plane")) zero_chunks(p->j_plane, p->nPQR, plane_size(p, 1));
  else if (!strcmp(f, "k_plane")) zero_chunks(p->k_plane, p->nPQR, plane_size(p, 2));
}

void ko_solve(ko_problem *p, int niter, int bj, double *particles) { /* SteadyStateSolver.cpp:39-104 */
  ko_zero(p, "psi");
  for (int it = 0; it < niter; ++it) {
    ko_zero(p, "phi");
    ko_ltimes(p);
    ko_zero(p, "phi_out");
    ko_scattering(p);
    ko_source(p);
    ko_zero(p, "rhs");
    ko_lplustimes(p);
    ko_sweep_solver(p, bj);
    double part = ko_population(p);
    if (particles) particles[it] = part;
  }
}

/* ---------------- introspection ---------------- */
int ko_num_subdomains(const ko_problem *p) { return p->nPQR; }
void ko_adjacency(const ko_problem *p, int s, int up[3], int dn[3]) {
  for (int d = 0; d < 3; ++d) { up[d] = p->upwind[s * 3 + d]; dn[d] = p->downwind[s * 3 + d]; }
}
int ko_dim(const ko_problem *p, const char *w) {
  if (!strcmp(w, "Gs")) return p->Gs; if (!strcmp(w, "Ds")) return p->Ds;
  if (!strcmp(w, "ni")) return p->ni; if (!strcmp(w, "nj")) return p->nj; if (!strcmp(w, "nk")) return p->nk;
  if (!strcmp(w, "M")) return p->M; if (!strcmp(w, "L1")) return p->L1;
  if (!strcmp(w, "G")) return p->G; if (!strcmp(w, "D")) return p->D; if (!strcmp(w, "Z")) return p->Z;
  if (!strcmp(w, "SP")) return p->SP; if (!strcmp(w, "SQ")) return p->SQ;
  if (!strcmp(w, "Sx")) return p->Sx; if (!strcmp(w, "Sy")) return p->Sy; if (!strcmp(w, "Sz")) return p->Sz;
  return -1;
}
size_t ko_offset_flux(const ko_problem *p, int d, int g, int z) { return d * p->fs_d + g * p->fs_g + z * p->fs_z; }
size_t ko_offset_moment(const ko_problem *p, int nm, int g, int z) { return nm * p->ms_m + g * p->ms_g + z * p->ms_z; }

typedef struct { const char *name; void **chunks; int nchunks; size_t n; int elsize; const int *var_n; } fdesc;

static int lookup(const ko_problem *p, const char *f, fdesc *o) {
  size_t nflux = (size_t)p->Ds * p->Gs * p->Zs, nmom = (size_t)p->M * p->Gs * p->Zs;
  o->var_n = NULL;
#define F(NAME, PTR, NCH, N, ES) if (!strcmp(f, NAME)) { o->name = NAME; o->chunks = (void **)(PTR); o->nchunks = NCH; o->n = N; o->elsize = ES; return 0; }
  F("psi", p->psi, p->nPQR, nflux, 8) F("rhs", p->rhs, p->nPQR, nflux, 8)
  F("phi", p->phi, p->nPR, nmom, 8) F("phi_out", p->phi_out, p->nPR, nmom, 8)
  F("i_plane", p->i_plane, p->nPQR, plane_size(p, 0), 8) F("j_plane", p->j_plane, p->nPQR, plane_size(p, 1), 8)
  F("k_plane", p->k_plane, p->nPQR, plane_size(p, 2), 8)
  F("sigt_zonal", p->sigt_zonal, p->nPR, (size_t)p->Gs * p->Zs, 8)
  F("ell", p->ell, p->nQ, (size_t)p->Ds * p->M, 8) F("ell_plus", p->ell_plus, p->nQ, (size_t)p->Ds * p->M, 8)
  F("quadrature/xcos", p->xcos, p->nQ, p->Ds, 8) F("quadrature/ycos", p->ycos, p->nQ, p->Ds, 8)
  F("quadrature/zcos", p->zcos, p->nQ, p->Ds, 8) F("quadrature/w", p->w, p->nQ, p->Ds, 8)
  F("quadrature/id", p->id, p->nQ, p->Ds, 4) F("quadrature/jd", p->jd, p->nQ, p->Ds, 4) F("quadrature/kd", p->kd, p->nQ, p->Ds, 4)
  F("dx", p->dx, p->Sx, p->ni, 8) F("dy", p->dy, p->Sy, p->nj, 8) F("dz", p->dz, p->Sz, p->nk, 8)
  F("volume", p->volume, p->nR, p->Zs, 8)
  F("zone_to_mixelem", p->zone_to_mixelem, p->nR, p->Zs, 4) F("zone_to_num_mixelem", p->zone_to_num_mixelem, p->nR, p->Zs, 4)
#undef F
  if (!strcmp(f, "mixelem_to_zone")) { o->chunks = (void **)p->mixelem_to_zone; o->nchunks = p->nR; o->elsize = 4; o->var_n = p->num_mixelem; return 0; }
  if (!strcmp(f, "mixelem_to_material")) { o->chunks = (void **)p->mixelem_to_material; o->nchunks = p->nR; o->elsize = 4; o->var_n = p->num_mixelem; return 0; }
  if (!strcmp(f, "mixelem_to_fraction")) { o->chunks = (void **)p->mixelem_to_fraction; o->nchunks = p->nR; o->elsize = 8; o->var_n = p->num_mixelem; return 0; }
  return 1;
}

int ko_field_chunks(const ko_problem *p, const char *f) {
  if (!strcmp(f, "data/sigs") || !strcmp(f, "moment_to_legendre")) return 1;
  if (!strcmp(f, "upwind") || !strcmp(f, "downwind")) return p->nPQR;
  fdesc d;
  return lookup(p, f, &d) ? -1 : d.nchunks;
}

int ko_field_chunk(const ko_problem *p, const char *f, int chunk, void **ptr, size_t *n, int *elsize) {
  if (!strcmp(f, "data/sigs")) { *ptr = p->sigs; *n = (size_t)3 * p->L1 * p->G * p->G; *elsize = 8; return chunk != 0; }
  if (!strcmp(f, "moment_to_legendre")) { *ptr = p->moment_to_legendre; *n = p->M; *elsize = 4; return chunk != 0; }
  if (!strcmp(f, "upwind")) { *ptr = p->upwind + 3 * chunk; *n = 3; *elsize = 4; return 0; }
  if (!strcmp(f, "downwind")) { *ptr = p->downwind + 3 * chunk; *n = 3; *elsize = 4; return 0; }
  fdesc d;
  if (lookup(p, f, &d) || chunk < 0 || chunk >= d.nchunks) return 1;
  *ptr = d.chunks[chunk];
  *n = d.var_n ? (size_t)d.var_n[chunk] : d.n;
  *elsize = d.elsize;
  return 0;
}
