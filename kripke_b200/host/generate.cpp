// Problem generation on the host (runs once): decomposition, energy groups, quadrature with the
// L / L+ matrices and sweep adjacency, the Kobayashi-3i mesh with sub-sampled material mixing,
// zonal total cross sections, and the flux / moment / plane / scattering-matrix fields.
//
// This defines the synthetic inputs of the benchmark, so every number must come out bit-identical
// to the reference's generator (src/Kripke/Generate.cpp, Generate/{Decomp,Energy,Quadrature,Space,
// Data}.cpp): the same libm calls in the same order, and std::sort from the same libstdc++ for the
// Gauss-Legendre direction ordering.
#include "Kripke.h"

#include <algorithm>
#include <cfloat>
#include <cmath>

using namespace Kripke;
using namespace Kripke::Core;

namespace {

struct QuadraturePoint {
  double xcos, ycos, zcos, w;
  int id, jd, kd, octant;
};

// n-point Gauss-Legendre rule on [x1,x2] by Newton iteration on P_n (Generate/Quadrature.cpp:39-71)
void gaussLegendre(double x1, double x2, std::vector<double> &x, std::vector<double> &w, double eps) {
  const int n = (int)x.size();
  const int m = (n + 1) / 2;
  const double xm = 0.5 * (x2 + x1), xl = 0.5 * (x2 - x1);
  for (int i = 1; i <= m; i++) {
    double z = cos(M_PI * (i - 0.25) / (n + 0.5));
    double z1, pp, p1, p2, p3;
    do {
      p1 = 1.0;
      p2 = 0.0;
      for (int j = 1; j <= n; j++) {
        p3 = p2;
        p2 = p1;
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

bool byDescendingOctant(QuadraturePoint const &a, QuadraturePoint const &b) { return b.octant < a.octant; }

double factorial(int n) {
  double f = 1.0;
  for (int i = n; i > 0; --i) f *= (double)i;
  return f;
}

// associated Legendre function P_n^m(x), 0 <= m <= n (Generate/Quadrature.cpp:87-138)
double legendrePnm(int n, int m, double x) {
  if (std::abs(x) > 1.0) KRIPKE_ABORT("Bad input to PnmFcn: abs(x) > 1.0, x = %e\n", x);
  double pmm = 1.0;
  if (m > 0) {
    const double somx2 = sqrt((1.0 - x) * (1.0 + x));
    double fact = 1.0;
    for (int i = 1; i <= m; i++) {
      pmm *= -fact * somx2;
      fact += 2.0;
    }
  }
  if (n == m) return pmm;
  double pmmp1 = x * (2 * m + 1) * pmm;
  if (n == m + 1) return pmmp1;
  double pnn = 0;
  for (int nn = m + 2; nn <= n; nn++) {
    pnn = (x * (2 * nn - 1) * pmmp1 - (nn + m - 1) * pmm) / (nn - m);
    pmm = pmmp1;
    pmmp1 = pnn;
  }
  return pnn;
}

// real spherical harmonic Y_nm(omega) (Generate/Quadrature.cpp:140-183)
double sphericalYnm(int n, int m, double mu, double eta, double xi) {
  const double tiny = 1.e-20;
  double tmp = fabs(eta / (mu + tiny));
  const double phi_tmp = atan(tmp);
  double phi;
  if ((mu > 0) && (eta > 0)) phi = phi_tmp;
  else if ((mu < 0) && (eta > 0)) phi = M_PI - fabs(phi_tmp);
  else if ((mu < 0) && (eta < 0)) phi = M_PI + fabs(phi_tmp);
  else phi = 2.0 * M_PI - fabs(phi_tmp);
  const double fac1 = factorial(n - std::abs(m));
  const double fac2 = factorial(n + std::abs(m));
  const double pnm = legendrePnm(n, std::abs(m), xi);
  tmp = ((double)m) * phi;
  const double taum = (m >= 0) ? cos(tmp) : sin(-tmp);
  const double dm0 = (m == 0) ? 1.0 : 0.0;
  tmp = ((2 * n + 1) * fac1) / (2.0 * (1.0 + dm0) * M_PI * fac2);
  const double anm = sqrt(tmp);
  return anm * pnm * taum;
}

// Generate/Quadrature.cpp:192-309
std::vector<QuadraturePoint> createQuadratureSet(InputVariables const &in) {
  const int num_directions = in.num_directions;
  std::vector<QuadraturePoint> dirs(num_directions);  // value-initialised: octant = 0 for the dummy set
  const int num_polar = in.quad_num_polar, num_azimuth = in.quad_num_azimuthal;
  if (num_polar > 0) {
    KRIPKE_ASSERT(num_polar % 4 == 0, "Must have number of polar angles be a multiple of 4\n");
    KRIPKE_ASSERT(num_azimuth % 2 == 0, "Must have number of azimuthal angles be a multiple of 2\n");
    KRIPKE_ASSERT(num_polar * num_azimuth == num_directions, "You need to specify %d total directions, not %d\n",
                  num_polar * num_azimuth, num_directions);
    std::vector<double> polar_cos(num_polar), polar_weight(num_polar);
    gaussLegendre(-1.0, 1.0, polar_cos, polar_weight, DBL_EPSILON);
    std::vector<double> az_angle(num_azimuth), az_weight(num_azimuth);
    const double dangle = 2.0 * M_PI / ((double)num_azimuth);
    for (int i = 0; i < num_azimuth; i++) {
      az_angle[i] = (i == 0) ? dangle / 2.0 : az_angle[i - 1] + dangle;
      az_weight[i] = dangle;
    }
    int d = 0;
    for (int i = 0; i < num_polar; i++)
      for (int j = 0; j < num_azimuth; j++, d++) {
        const double xcos = sqrt(1.0 - polar_cos[i] * polar_cos[i]) * cos(az_angle[j]);
        const double ycos = sqrt(1.0 - polar_cos[i] * polar_cos[i]) * sin(az_angle[j]);
        const double zcos = polar_cos[i];
        QuadraturePoint &q = dirs[d];
        q.id = (xcos > 0.) ? 1 : -1;
        q.jd = (ycos > 0.) ? 1 : -1;
        q.kd = (zcos > 0.) ? 1 : -1;
        q.octant = (q.id == -1 ? 1 : 0) + (q.jd == -1 ? 2 : 0) + (q.kd == -1 ? 4 : 0);
        q.xcos = std::abs(xcos);
        q.ycos = std::abs(ycos);
        q.zcos = std::abs(zcos);
        q.w = polar_weight[i] * az_weight[j];
      }
    std::sort(dirs.begin(), dirs.end(), byDescendingOctant);  // same libstdc++ => same order as the reference
  } else {
    // "dummy S2": one direction per octant, repeated (Generate/Quadrature.cpp:280-306)
    const double mu = cos(M_PI / 4);
    const double eta = sqrt(1 - mu * mu) * cos(M_PI / 4);
    const double xi = sqrt(1 - mu * mu) * sin(M_PI / 4);
    const int per_octant = num_directions / 8;
    int d = 0;
    for (int octant = 0; octant < 8; ++octant)
      for (int sd = 0; sd < per_octant; sd++, d++) {
        QuadraturePoint &q = dirs[d];
        q.id = (octant & 0x1) ? 1 : -1;
        q.jd = ((octant >> 1) & 0x1) ? 1 : -1;
        q.kd = ((octant >> 2) & 0x1) ? 1 : -1;
        q.w = 4.0 * M_PI / (double)num_directions;
        q.xcos = mu;
        q.ycos = eta;
        q.zcos = xi;
      }
  }
  return dirs;
}

// Kobayashi-3i regions (Generate/Space.cpp:142-177)
int materialAt(double x, double y, double z) {
  x = std::abs(x);
  y = std::abs(y);
  z = std::abs(z);
  if (x <= 10.0 && y <= 10.0 && z <= 10.0) return 0;
  if (x <= 10.0 && y <= 60.0 && z <= 10.0) return 1;
  if (x <= 40.0 && y >= 50.0 && y <= 60.0 && z <= 10.0) return 1;
  if (x >= 30.0 && x <= 40.0 && y >= 50.0 && y <= 60.0 && z <= 40.0) return 1;
  if (x >= 30.0 && x <= 40.0 && y >= 50.0 && z >= 30.0 && z <= 40.0) return 1;
  return 2;
}

template <typename F>
void fillField(F &field, typename F::ElementType v) {
  for (SdomId s : field.getWorkList()) {
    auto *p = field.getData(s);
    for (size_t i = 0; i < field.size(s); ++i) p[i] = v;
  }
}

}  // namespace

// ---- Generate/Decomp.cpp:23-56 -----------------------------------------------------------------
void Kripke::Generate::generateDecomp(DataStore &data_store, InputVariables const &in) {
  auto &comm = data_store.newVariable<Comm>("comm");
  auto &al_var = data_store.newVariable<ArchLayout>("al");
  al_var.al_v = in.al_v;
  auto &pspace = data_store.newVariable<PartitionSpace>("pspace", comm, 1, 1, in.npx, in.npy, in.npz);
  pspace.setup_createSubdomains(in.num_groupsets, in.num_dirsets, in.num_zonesets_dim[0], in.num_zonesets_dim[1],
                                in.num_zonesets_dim[2]);
  pspace.createSubdomainData(data_store);
  pspace.print();
}

// ---- Generate/Energy.cpp:22-42 -----------------------------------------------------------------
void Kripke::Generate::generateEnergy(DataStore &data_store, InputVariables const &in) {
  PartitionSpace &pspace = data_store.getVariable<PartitionSpace>("pspace");
  size_t ngrp_per_sdom = in.num_groups / pspace.getGlobalNumSubdomains(SPACE_P);
  std::vector<size_t> local_grps(pspace.getNumSubdomains(SPACE_P), ngrp_per_sdom);
  RangeSet *grp_set = new RangeSet(pspace, SPACE_P, local_grps);
  data_store.addVariable("Set/Group", grp_set);
  data_store.addVariable("Set/GlobalGroup", new GlobalRangeSet(pspace, *grp_set));
}

// ---- Generate/Quadrature.cpp:315-519 -------------------------------------------------------------
void Kripke::Generate::generateQuadrature(DataStore &data_store, InputVariables const &in) {
  PartitionSpace &pspace = data_store.getVariable<PartitionSpace>("pspace");
  ArchLayoutV al_v = data_store.getVariable<ArchLayout>("al").al_v;

  size_t ndir_per_sdom = in.num_directions / pspace.getGlobalNumSubdomains(SPACE_Q);
  std::vector<size_t> local_dirs(pspace.getNumSubdomains(SPACE_Q), ndir_per_sdom);
  RangeSet *dir_set = new RangeSet(pspace, SPACE_Q, local_dirs);
  data_store.addVariable("Set/Direction", dir_set);

  const int L1 = in.legendre_order + 1;
  const size_t num_moments = (size_t)L1 * L1;
  GlobalRangeSet *moment_set = new GlobalRangeSet(pspace, num_moments);
  data_store.addVariable("Set/Moment", moment_set);
  data_store.addVariable("Set/Legendre", new GlobalRangeSet(pspace, (size_t)L1));

  auto &f_m2l = createField<Field_Moment2Legendre>(data_store, "moment_to_legendre", al_v, *moment_set);
  for (SdomId s : f_m2l.getWorkList()) {
    int *m2l = f_m2l.getData(s);
    int nm = 0;
    for (int n = 0; n < L1; ++n)
      for (int m = -n; m <= n; ++m) m2l[nm++] = n;
    KRIPKE_ASSERT(nm == (int)moment_set->globalSize(), "moment count mismatch\n");
  }

  auto points = createQuadratureSet(in);
  auto &f_xcos = createField<Field_Direction2Double>(data_store, "quadrature/xcos", al_v, *dir_set);
  auto &f_ycos = createField<Field_Direction2Double>(data_store, "quadrature/ycos", al_v, *dir_set);
  auto &f_zcos = createField<Field_Direction2Double>(data_store, "quadrature/zcos", al_v, *dir_set);
  auto &f_w = createField<Field_Direction2Double>(data_store, "quadrature/w", al_v, *dir_set);
  auto &f_id = createField<Field_Direction2Int>(data_store, "quadrature/id", al_v, *dir_set);
  auto &f_jd = createField<Field_Direction2Int>(data_store, "quadrature/jd", al_v, *dir_set);
  auto &f_kd = createField<Field_Direction2Int>(data_store, "quadrature/kd", al_v, *dir_set);
  auto &f_oct = createField<Field_Direction2Int>(data_store, "quadrature/octant", al_v, *dir_set);
  for (SdomId s : f_xcos.getWorkList()) {
    const int nd = (int)dir_set->size(s), lower = (int)dir_set->lower(s);
    double *xc = f_xcos.getData(s), *yc = f_ycos.getData(s), *zc = f_zcos.getData(s), *w = f_w.getData(s);
    int *id = f_id.getData(s), *jd = f_jd.getData(s), *kd = f_kd.getData(s), *oc = f_oct.getData(s);
    for (int d = 0; d < nd; ++d) {
      QuadraturePoint const &q = points[d + lower];
      xc[d] = q.xcos; yc[d] = q.ycos; zc[d] = q.zcos; w[d] = q.w;
      id[d] = q.id; jd[d] = q.jd; kd[d] = q.kd; oc[d] = q.octant;
    }
  }

  auto &set_ell = data_store.newVariable<ProductSet<2>>("Set/Ell", pspace, SPACE_Q, *moment_set, *dir_set);
  auto &set_ell_plus = data_store.newVariable<ProductSet<2>>("Set/EllPlus", pspace, SPACE_Q, *dir_set, *moment_set);
  auto &f_ell = createField<Field_Ell>(data_store, "ell", al_v, set_ell);
  auto &f_ell_plus = createField<Field_EllPlus>(data_store, "ell_plus", al_v, set_ell_plus);
  const double SQRT4PI = std::sqrt(4 * M_PI);
  for (SdomId s : f_xcos.getWorkList()) {
    auto ell = f_ell.getView(s);
    auto ell_plus = f_ell_plus.getView(s);
    const int nd = (int)dir_set->size(s), lower = (int)dir_set->lower(s);
    int nm = 0;
    for (int n = 0; n < L1; n++)
      for (int m = -n; m <= n; m++, nm++)
        for (int d = 0; d < nd; ++d) {
          QuadraturePoint const &q = points[d + lower];
          const double xcos = (q.id) * (q.xcos), ycos = (q.jd) * (q.ycos), zcos = (q.kd) * (q.zcos);
          const double ynm = sphericalYnm(n, m, xcos, ycos, zcos);
          ell(nm, d) = q.w * ynm / SQRT4PI;
          ell_plus(d, nm) = ynm * SQRT4PI;
        }
  }

  // upwind / downwind subdomain adjacency (Generate/Quadrature.cpp:457-518)
  auto &set_dimension = data_store.newVariable<GlobalRangeSet>("Set/Dimension", pspace, 3);
  auto &set_adjacency = data_store.newVariable<ProductSet<1>>("Set/Adjacency", pspace, SPACE_PQR, set_dimension);
  auto &f_up = createField<Field_Adjacency>(data_store, "upwind", al_v, set_adjacency);
  auto &f_down = createField<Field_Adjacency>(data_store, "downwind", al_v, set_adjacency);
  for (SdomId s : f_up.getWorkList()) {
    auto global_coord = pspace.coordToGlobalCoord(pspace.sdomIdToCoord(s));
    const int sweep_dir[3] = {f_id.getDataConst(s)[0], f_jd.getDataConst(s)[0], f_kd.getDataConst(s)[0]};
    long *up = f_up.getData(s), *down = f_down.getData(s);
    for (int dim = 0; dim < 3; ++dim) {
      auto gu = global_coord, gd = global_coord;
      gu[dim + SPACE_RX] -= sweep_dir[dim];
      gd[dim + SPACE_RX] += sweep_dir[dim];
      const ptrdiff_t n = (ptrdiff_t)pspace.getGlobalNumSubdomains((SPACE)(dim + SPACE_RX));
      up[dim] = (gu[dim + SPACE_RX] < 0 || gu[dim + SPACE_RX] >= n) ? -1 : *pspace.coordToGlobalSdomId(gu);
      down[dim] = (gd[dim + SPACE_RX] < 0 || gd[dim + SPACE_RX] >= n) ? -1 : *pspace.coordToGlobalSdomId(gd);
    }
  }
}

// ---- Generate/Space.cpp:38-365 ---------------------------------------------------------------------
void Kripke::Generate::generateSpace(DataStore &data_store, InputVariables const &in) {
  PartitionSpace &pspace = data_store.getVariable<PartitionSpace>("pspace");
  ArchLayoutV al_v = data_store.getVariable<ArchLayout>("al").al_v;

  auto make_axis = [&](const char *name, SPACE space, int n, const char *axis) -> RangeSet & {
    size_t per = n / pspace.getGlobalNumSubdomains(space);
    KRIPKE_ASSERT(per * pspace.getGlobalNumSubdomains(space) == (size_t)n,
                  "Number of zones in %s must evenly divide into the number of subdomains\n", axis);
    std::vector<size_t> local(pspace.getNumSubdomains(space), per);
    return data_store.newVariable<RangeSet>(name, pspace, space, local);
  };
  auto &set_zonei = make_axis("Set/ZoneI", SPACE_RX, in.nx, "X");
  auto &set_zonej = make_axis("Set/ZoneJ", SPACE_RY, in.ny, "Y");
  auto &set_zonek = make_axis("Set/ZoneK", SPACE_RZ, in.nz, "Z");
  auto &set_zone = data_store.newVariable<ProductSet<3>>("Set/Zone", pspace, SPACE_R, set_zonek, set_zonej, set_zonei);
  auto &set_zone_linear = data_store.newVariable<ProductSet<1>>("Set/ZoneLinear", pspace, SPACE_R, set_zone);
  data_store.newVariable<GlobalRangeSet>("Set/Material", pspace, 3);

  const double x_min = -60.0, x_max = 60.0, y_min = -100.0, y_max = 100.0, z_min = -60.0, z_max = 60.0;
  auto &f_dx = createField<Field_ZoneI2Double>(data_store, "dx", al_v, set_zonei);
  const double dx = (x_max - x_min) / set_zonei.globalSize();
  fillField(f_dx, dx);
  auto &f_dy = createField<Field_ZoneJ2Double>(data_store, "dy", al_v, set_zonej);
  const double dy = (y_max - y_min) / set_zonej.globalSize();
  fillField(f_dy, dy);
  auto &f_dz = createField<Field_ZoneK2Double>(data_store, "dz", al_v, set_zonek);
  const double dz = (z_max - z_min) / set_zonek.globalSize();
  fillField(f_dz, dz);
  const double zone_volume = dx * dy * dz;
  auto &f_volume = createField<Field_Zone2Double>(data_store, "volume", al_v, set_zone_linear);
  fillField(f_volume, zone_volume);

  double total_volume[3] = {0.0, 0.0, 0.0};
  const int ns = in.num_material_subsamples;
  auto sdom_list = set_zone.getWorkList();
  // With a GPU bound to the process the zone loops run there (csrc/kb200_generate.cu, bit-identical
  // tables); without one -- planning runs, the CPU tests against the oracle -- the host loops below do the same work.
  const bool on_device = kb200_device_bound() != 0 && !(getenv("KB200_HOST_GENERATOR") && getenv("KB200_HOST_GENERATOR")[0] == '1');
  if (on_device) {
    std::vector<size_t> sdom_to_num_mixed;
    std::vector<double *> fracs;
    auto &f_z2n = createField<Field_Zone2Int>(data_store, "zone_to_num_mixelem", al_v, set_zone_linear);
    auto &f_z2m = createField<Field_Zone2MixElem>(data_store, "zone_to_mixelem", al_v, set_zone_linear);
    for (SdomId s : sdom_list) {
      kb200_zoneset_desc zs;
      zs.ni = (int)set_zonei.size(s); zs.nj = (int)set_zonej.size(s); zs.nk = (int)set_zonek.size(s);
      zs.i0 = (int)set_zonei.lower(s); zs.j0 = (int)set_zonej.lower(s); zs.k0 = (int)set_zonek.lower(s);
      zs.x_min = x_min; zs.y_min = y_min; zs.z_min = z_min; zs.dx = dx; zs.dy = dy; zs.dz = dz; zs.ns = ns;
      double *d_frac = nullptr;
      KB200_CALL(kb200_alloc(3 * set_zone.size(s) * sizeof(double), (void **)&d_frac));
      int nmix = 0;
      double vol[3];
      KB200_CALL(kb200_generate_mix_count(&zs, d_frac, f_z2n.devicePtrOverwrite(s), f_z2m.devicePtrOverwrite(s), &nmix, vol, nullptr));
      for (int m = 0; m < 3; ++m) total_volume[m] += vol[m];
      sdom_to_num_mixed.push_back((size_t)nmix);
      fracs.push_back(d_frac);
    }
    auto &set_mixelem = data_store.newVariable<RangeSet>("Set/MixElem", pspace, SPACE_R, sdom_to_num_mixed);
    auto &f_m2z = createField<Field_MixElem2Zone>(data_store, "mixelem_to_zone", al_v, set_mixelem);
    auto &f_m2m = createField<Field_MixElem2Material>(data_store, "mixelem_to_material", al_v, set_mixelem);
    auto &f_m2f = createField<Field_MixElem2Double>(data_store, "mixelem_to_fraction", al_v, set_mixelem);
    for (size_t idx = 0; idx < sdom_list.size(); ++idx) {
      SdomId s = sdom_list[idx];
      KB200_CALL(kb200_generate_mix_fill((int)set_zone.size(s), fracs[idx], f_z2m.devicePtrConst(s), f_m2z.devicePtrOverwrite(s),
                                         f_m2m.devicePtrOverwrite(s), f_m2f.devicePtrOverwrite(s), nullptr));
    }
    KB200_CALL(kb200_stream_sync(nullptr));
    for (double *p : fracs) KB200_CALL(kb200_free(p));
    Comm default_comm;
    pspace.getComm(SPACE_R).allReduceSumDouble(total_volume, 3);
    if (default_comm.rank() == 0)
      printf("\n  Material Volumes=[%e, %e, %e]\n", total_volume[0], total_volume[1], total_volume[2]);
    auto &set_group = data_store.getVariable<Set>("Set/Group");
    auto &set_sigt = data_store.newVariable<ProductSet<2>>("Set/SigmaTZonal", pspace, SPACE_PR, set_group, set_zone);
    auto &f_sigt = createField<Field_SigmaTZonal>(data_store, "sigt_zonal", al_v, set_sigt);
    for (SdomId s : f_sigt.getWorkList())
      KB200_CALL(kb200_generate_sigt((int)al_v.layout_v, (int)set_group.size(s), (int)set_zone.size(s), f_z2m.devicePtrConst(s),
                                     f_z2n.devicePtrConst(s), f_m2m.devicePtrConst(s), f_m2f.devicePtrConst(s), in.sigt,
                                     f_sigt.devicePtrOverwrite(s), nullptr));
    return;
  }

  // sub-sample every zone to get its material volume fractions
  const double sample_vol_frac = 1.0 / (double)(ns * ns * ns);
  struct Mix { double fraction[3]; };
  std::vector<std::vector<Mix>> mix;
  std::vector<size_t> sdom_to_num_mixed;
  for (SdomId s : sdom_list) {
    const double x0 = x_min + dx * set_zonei.lower(s), y0 = y_min + dy * set_zonej.lower(s), z0 = z_min + dz * set_zonek.lower(s);
    std::vector<Mix> sdom_mix(set_zone.size(s));
    auto zone_layout = set_zone.getLayout(s);
    const int ni = (int)set_zonei.size(s), nj = (int)set_zonej.size(s), nk = (int)set_zonek.size(s);
    for (int k = 0; k < nk; k++)
      for (int j = 0; j < nj; j++)
        for (int i = 0; i < ni; i++) {
          Mix &zm = sdom_mix[zone_layout(k, j, i)];
          zm = Mix{{0.0, 0.0, 0.0}};
          const double xi = x0 + dx * i, yi = y0 + dy * j, zi = z0 + dz * k;
          for (int si = 0; si < ns; ++si)
            for (int sj = 0; sj < ns; ++sj)
              for (int sk = 0; sk < ns; ++sk) {
                const double x = xi + dx * (si + 1) / (ns + 1);
                const double y = yi + dy * (sj + 1) / (ns + 1);
                const double z = zi + dz * (sk + 1) / (ns + 1);
                zm.fraction[materialAt(x, y, z)] += sample_vol_frac;
              }
        }
    size_t n = 0;
    for (auto &z : sdom_mix) n += (z.fraction[0] > 0.0) + (z.fraction[1] > 0.0) + (z.fraction[2] > 0.0);
    sdom_to_num_mixed.push_back(n);
    mix.push_back(std::move(sdom_mix));
  }

  auto &set_mixelem = data_store.newVariable<RangeSet>("Set/MixElem", pspace, SPACE_R, sdom_to_num_mixed);
  auto &f_m2z = createField<Field_MixElem2Zone>(data_store, "mixelem_to_zone", al_v, set_mixelem);
  auto &f_m2m = createField<Field_MixElem2Material>(data_store, "mixelem_to_material", al_v, set_mixelem);
  auto &f_m2f = createField<Field_MixElem2Double>(data_store, "mixelem_to_fraction", al_v, set_mixelem);
  auto &f_z2n = createField<Field_Zone2Int>(data_store, "zone_to_num_mixelem", al_v, set_zone_linear);
  auto &f_z2m = createField<Field_Zone2MixElem>(data_store, "zone_to_mixelem", al_v, set_zone_linear);

  for (size_t idx = 0; idx < sdom_list.size(); ++idx) {
    SdomId s = sdom_list[idx];
    const int num_zones = (int)set_zone.size(s);
    int *m2z = f_m2z.getData(s), *m2m = f_m2m.getData(s), *z2n = f_z2n.getData(s), *z2m = f_z2m.getData(s);
    double *m2f = f_m2f.getData(s);
    int me = 0;
    for (int z = 0; z < num_zones; ++z) {
      Mix const &zm = mix[idx][z];
      z2m[z] = me;
      int cnt = 0;
      double zone_frac = 0.0;
      for (int m = 0; m < 3; ++m)
        if (zm.fraction[m] > 0.0) {
          m2z[me] = z;
          m2m[me] = m;
          m2f[me] = zm.fraction[m];
          zone_frac += zm.fraction[m];
          total_volume[m] += zm.fraction[m] * zone_volume;
          ++cnt;
          ++me;
        }
      KRIPKE_ASSERT(zone_frac == 1.0, "Zone fraction wrong: %e", zone_frac);
      z2n[z] = cnt;
    }
    KRIPKE_ASSERT(me == (int)set_mixelem.size(s), "Mismatch in mixture info");
  }
  Comm default_comm;
  pspace.getComm(SPACE_R).allReduceSumDouble(total_volume, 3);
  if (default_comm.rank() == 0)
    printf("\n  Material Volumes=[%e, %e, %e]\n", total_volume[0], total_volume[1], total_volume[2]);

  // zonal total cross section (Generate/Space.cpp:331-363)
  auto &set_group = data_store.getVariable<Set>("Set/Group");
  auto &set_sigt = data_store.newVariable<ProductSet<2>>("Set/SigmaTZonal", pspace, SPACE_PR, set_group, set_zone);
  auto &f_sigt = createField<Field_SigmaTZonal>(data_store, "sigt_zonal", al_v, set_sigt);
  for (SdomId s : f_sigt.getWorkList()) {
    auto sigt = f_sigt.getView(s);
    const int *m2z = f_m2z.getDataConst(s), *m2m = f_m2m.getDataConst(s);
    const double *m2f = f_m2f.getDataConst(s);
    const int num_groups = (int)set_group.size(s), num_mixelem = (int)set_mixelem.size(s);
    for (size_t i = 0; i < f_sigt.size(s); ++i) sigt.ptr[i] = 0.0;
    for (int g = 0; g < num_groups; ++g)
      for (int me = 0; me < num_mixelem; ++me) sigt(g, m2z[me]) += m2f[me] * in.sigt[m2m[me]];
  }
}

// ---- Generate/Data.cpp:22-105 ------------------------------------------------------------------------
void Kripke::Generate::generateData(DataStore &data_store, InputVariables const &in) {
  PartitionSpace &pspace = data_store.getVariable<PartitionSpace>("pspace");
  ArchLayoutV al_v = data_store.getVariable<ArchLayout>("al").al_v;
  Set const &dir_set = data_store.getVariable<Set>("Set/Direction");
  Set const &group_set = data_store.getVariable<Set>("Set/Group");
  Set const &zone_set = data_store.getVariable<Set>("Set/Zone");
  ProductSet<3> *flux_set = new ProductSet<3>(pspace, SPACE_PQR, dir_set, group_set, zone_set);
  data_store.addVariable("Set/Flux", flux_set);
  createField<Field_Flux>(data_store, "psi", al_v, *flux_set);
  createField<Field_Flux>(data_store, "rhs", al_v, *flux_set);

  Set const &moment_set = data_store.getVariable<Set>("Set/Moment");
  ProductSet<3> *fluxmoment_set = new ProductSet<3>(pspace, SPACE_PR, moment_set, group_set, zone_set);
  data_store.addVariable("Set/FluxMoment", fluxmoment_set);
  createField<Field_Moments>(data_store, "phi", al_v, *fluxmoment_set);
  createField<Field_Moments>(data_store, "phi_out", al_v, *fluxmoment_set);

  Set const &zonei_set = data_store.getVariable<Set>("Set/ZoneI");
  Set const &zonej_set = data_store.getVariable<Set>("Set/ZoneJ");
  Set const &zonek_set = data_store.getVariable<Set>("Set/ZoneK");
  Set const &iplane_set = data_store.newVariable<ProductSet<4>>("Set/IPlane", pspace, SPACE_PQR, dir_set, group_set, zonej_set, zonek_set);
  Set const &jplane_set = data_store.newVariable<ProductSet<4>>("Set/JPlane", pspace, SPACE_PQR, dir_set, group_set, zonei_set, zonek_set);
  Set const &kplane_set = data_store.newVariable<ProductSet<4>>("Set/KPlane", pspace, SPACE_PQR, dir_set, group_set, zonei_set, zonej_set);
  createField<Field_IPlane>(data_store, "i_plane", al_v, iplane_set);
  createField<Field_JPlane>(data_store, "j_plane", al_v, jplane_set);
  createField<Field_KPlane>(data_store, "k_plane", al_v, kplane_set);

  Set const &material_set = data_store.getVariable<Set>("Set/Material");
  Set const &legendre_set = data_store.getVariable<Set>("Set/Legendre");
  Set const &global_group_set = data_store.getVariable<Set>("Set/GlobalGroup");
  ProductSet<4> *sigs_set = new ProductSet<4>(pspace, SPACE_NULL, material_set, legendre_set, global_group_set, global_group_set);
  data_store.addVariable("Set/SigmaS", sigs_set);
  auto &f_sigs = createField<Field_SigmaS>(data_store, "data/sigs", al_v, *sigs_set);
  for (SdomId s : f_sigs.getWorkList()) {
    auto sigs = f_sigs.getView(s);
    for (size_t i = 0; i < f_sigs.size(s); ++i) sigs.ptr[i] = 0.0;
    const int G = (int)global_group_set.size(s);
    for (int mat = 0; mat < 3; ++mat)
      for (int g = 0; g < G; ++g) sigs(mat, 0, g, g) = in.sigs[mat];
  }
}

// ---- Generate.cpp:24-94 ------------------------------------------------------------------------------
void Kripke::generateProblem(DataStore &data_store, InputVariables const &in) {
  Comm default_comm;
  if (default_comm.rank() == 0) {
    printf("\nGenerating Problem\n");
    printf("==================\n\n");
  }
  data_store.addVariable("timing", new Kripke::Timing());
  KRIPKE_TIMER(data_store, Generate);
  Generate::generateDecomp(data_store, in);
  Generate::generateEnergy(data_store, in);
  Generate::generateQuadrature(data_store, in);
  Generate::generateSpace(data_store, in);
  Generate::generateData(data_store, in);

  if (default_comm.rank() == 0) {
    std::vector<std::string> names;
    for (auto const &name : data_store.getVariableList())
      if (data_store.isVariableType<FieldStorage<double>>(name)) names.push_back(name);
    std::sort(names.begin(), names.end());
    printf("\n  Memory breakdown of Field variables:\n");
    printf("  Field Variable            Num Elements    Megabytes\n");
    printf("  --------------            ------------    ---------\n");
    unsigned long total = 0;
    for (auto const &name : names) {
      unsigned long n = data_store.getVariable<FieldStorage<double>>(name).getSet().globalSize();
      total += n;
      printf("  %-24s  %12lu %12.3lf\n", name.c_str(), n, (double)n * 8.0 / 1024.0 / 1024.0);
    }
    printf("  --------                  ------------    ---------\n");
    printf("  TOTAL                     %12lu %12.3lf\n\n  Generation Complete!\n", total, (double)total * 8.0 / 1024.0 / 1024.0);
  }
}
