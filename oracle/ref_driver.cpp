// TEST INFRASTRUCTURE (oracle) -- not product code.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference leg may execute this.
//
// High-precision driver around the UNMODIFIED reference (compiled from /root/reference by
// oracle/Makefile `make ref`).  It re-states the source-iteration loop of
// src/Kripke/SteadyStateSolver.cpp:39-104 through the reference's public entry points
// (Kripke::Kernel::*, Kripke::SweepSolver, Kripke::generateProblem) and prints what the
// reference's own stdout rounds away: %.17g particle counts per iteration, L2 norms / sums
// of phi, phi_out, rhs, psi, and (optionally) raw dumps of every field chunk so that the
// C restatement (oracle/kripke_oracle.c) and the CUDA path can be compared element-wise.
//
//   kripke_ref [kripke flags] [--dump DIR] [--adjacency] [--time]
//
#include <Kripke.h>
#include <Kripke/Core/DataStore.h>
#include <Kripke/Core/PartitionSpace.h>
#include <Kripke/Generate.h>
#include <Kripke/InputVariables.h>
#include <Kripke/Kernel.h>
#include <Kripke/ParallelComm.h>
#include <Kripke/SweepSolver.h>
#include <Kripke/Timing.h>
#include <Kripke/VarTypes.h>
#include <chrono>
#include <cmath>
#include <cstring>
#include <string>
#include <sys/stat.h>

using namespace Kripke;
using namespace Kripke::Core;

template <typename F>
static void fnorm(Kripke::Core::DataStore &ds, const char *name) {
  auto &f = ds.getVariable<F>(name);
  long double s2 = 0, s1 = 0;
  size_t n = 0;
  for (SdomId s : f.getWorkList()) {
    double *p = f.getData(s);
    size_t m = f.size(s);
    n += m;
    for (size_t i = 0; i < m; ++i) {
      s2 += (long double)p[i] * p[i];
      s1 += p[i];
    }
  }
  printf("NORM %s n=%zu l2=%.17g sum=%.17g\n", name, n, (double)sqrtl(s2), (double)s1);
}

template <typename F>
static void fdump(Kripke::Core::DataStore &ds, const char *name, std::string const &dir) {
  auto &f = ds.getVariable<F>(name);
  std::string fn(name);
  for (auto &c : fn)
    if (c == '/') c = '_';
  FILE *fp = fopen((dir + "/" + fn + ".bin").c_str(), "wb");
  if (!fp) {
    printf("cannot open dump file for %s\n", name);
    exit(1);
  }
  size_t nchunks = 0, total = 0;
  for (SdomId s : f.getWorkList()) {
    auto *p = f.getData(s);
    size_t m = f.size(s);
    fwrite((void *)p, sizeof(*p), m, fp);
    ++nchunks;
    total += m;
  }
  fclose(fp);
  printf("DUMP %s chunks=%zu elems=%zu elsize=%zu\n", name, nchunks, total,
         sizeof(typename F::ElementType));
}

int main(int argc, char **argv) {
  InputVariables v;
  bool bj = false, adjacency = false, timeit = false;
  std::string dumpdir;
  for (int i = 1; i < argc; ++i) {
    std::string o = argv[i];
    auto nx = [&]() { return std::string(argv[++i]); };
    auto three = [&](int *a) { sscanf(argv[++i], "%d,%d,%d", a, a + 1, a + 2); };
    if (o == "--zones") { int a[3]; three(a); v.nx = a[0]; v.ny = a[1]; v.nz = a[2]; }
    else if (o == "--procs") { int a[3]; three(a); v.npx = a[0]; v.npy = a[1]; v.npz = a[2]; }
    else if (o == "--zset") { three(v.num_zonesets_dim); }
    else if (o == "--groups") v.num_groups = atoi(nx().c_str());
    else if (o == "--quad") {
      std::string q = nx(); int p, a;
      if (sscanf(q.c_str(), "%d:%d", &p, &a) == 2) { v.quad_num_polar = p; v.quad_num_azimuthal = a; v.num_directions = p * a; }
      else v.num_directions = atoi(q.c_str());
    }
    else if (o == "--legendre") v.legendre_order = atoi(nx().c_str());
    else if (o == "--dset") v.num_dirsets = atoi(nx().c_str());
    else if (o == "--gset") v.num_groupsets = atoi(nx().c_str());
    else if (o == "--niter") v.niter = atoi(nx().c_str());
    else if (o == "--arch") v.al_v.arch_v = stringToArch(nx());
    else if (o == "--layout") v.al_v.layout_v = stringToLayout(nx());
    else if (o == "--pmethod") bj = !strcasecmp(nx().c_str(), "bj");
    else if (o == "--sigt") { double a[3]; sscanf(argv[++i], "%lf,%lf,%lf", a, a + 1, a + 2); for (int m = 0; m < 3; ++m) v.sigt[m] = a[m]; }
    else if (o == "--sigs") { double a[3]; sscanf(argv[++i], "%lf,%lf,%lf", a, a + 1, a + 2); for (int m = 0; m < 3; ++m) v.sigs[m] = a[m]; }
    else if (o == "--dump") dumpdir = nx();
    else if (o == "--adjacency") adjacency = true;
    else if (o == "--time") timeit = true;
    else { printf("bad opt %s\n", o.c_str()); return 1; }
  }
  if (v.checkValues()) return 1;

  Kripke::Core::DataStore ds;
  generateProblem(ds, v);
  auto &pspace = ds.getVariable<PartitionSpace>("pspace");
  int ns = pspace.getNumSubdomains(SPACE_PQR);

  if (adjacency) {
    auto &up = ds.getVariable<Field_Adjacency>("upwind");
    auto &dn = ds.getVariable<Field_Adjacency>("downwind");
    for (SdomId s{0}; s < ns; ++s) {
      auto c = pspace.sdomIdToCoord(s);
      auto u = up.getView(s), d = dn.getView(s);
      printf("SDOM %d coord=%d,%d,%d,%d,%d up=%d,%d,%d down=%d,%d,%d\n", (int)*s, (int)c[0], (int)c[1],
             (int)c[2], (int)c[3], (int)c[4], (int)*u(Dimension{0}), (int)*u(Dimension{1}),
             (int)*u(Dimension{2}), (int)*d(Dimension{0}), (int)*d(Dimension{1}), (int)*d(Dimension{2}));
    }
  }

  auto t0 = std::chrono::steady_clock::now();
  auto t_prev = t0;
  Kernel::kConst(ds.getVariable<Field_Flux>("psi"), 0.0);
  for (int it = 0; it < v.niter; ++it) {
    Kernel::kConst(ds.getVariable<Field_Moments>("phi"), 0.0);
    Kernel::LTimes(ds);
    Kernel::kConst(ds.getVariable<Field_Moments>("phi_out"), 0.0);
    Kernel::scattering(ds);
    Kernel::source(ds);
    Kernel::kConst(ds.getVariable<Field_Flux>("rhs"), 0.0);
    Kernel::LPlusTimes(ds);
    std::vector<SdomId> l(ns);
    for (SdomId i{0}; i < ns; ++i) l[*i] = i;
    SweepSolver(ds, l, bj);
    double part = Kernel::population(ds);
    printf("ITER %d particles=%.17g\n", it, part);
    if (timeit) {
      auto tn = std::chrono::steady_clock::now();
      printf("ITER_TIME %d %.6f\n", it, std::chrono::duration<double>(tn - t_prev).count());
      t_prev = tn;
    }
  }
  auto t1 = std::chrono::steady_clock::now();

  fnorm<Field_Moments>(ds, "phi");
  fnorm<Field_Moments>(ds, "phi_out");
  fnorm<Field_Flux>(ds, "rhs");
  fnorm<Field_Flux>(ds, "psi");

  if (timeit) {
    double solve = std::chrono::duration<double>(t1 - t0).count();
    size_t unknowns = ds.getVariable<Set>("Set/Group").globalSize() *
                      ds.getVariable<Set>("Set/Direction").globalSize() *
                      ds.getVariable<Set>("Set/Zone").globalSize();
    printf("TIME solve_s=%.6f niter=%d unknowns=%zu grind_ns=%.6f\n", solve, v.niter, unknowns,
           1e9 * solve / v.niter / (double)unknowns);
    auto &timing = ds.getVariable<Kripke::Timing>("timing");
    const char *names[] = {"LTimes", "LPlusTimes", "Scattering", "Source", "SweepSolver", "SweepSubdomain", "Population"};
    for (auto n : names) printf("TIMER %s %.6f\n", n, timing.getTotal(n));
  }

  if (!dumpdir.empty()) {
    mkdir(dumpdir.c_str(), 0777);
    fdump<Field_Flux>(ds, "psi", dumpdir);
    fdump<Field_Flux>(ds, "rhs", dumpdir);
    fdump<Field_Moments>(ds, "phi", dumpdir);
    fdump<Field_Moments>(ds, "phi_out", dumpdir);
    fdump<Field_IPlane>(ds, "i_plane", dumpdir);
    fdump<Field_JPlane>(ds, "j_plane", dumpdir);
    fdump<Field_KPlane>(ds, "k_plane", dumpdir);
    fdump<Field_Ell>(ds, "ell", dumpdir);
    fdump<Field_EllPlus>(ds, "ell_plus", dumpdir);
    fdump<Field_SigmaS>(ds, "data/sigs", dumpdir);
    fdump<Field_SigmaTZonal>(ds, "sigt_zonal", dumpdir);
    fdump<Field_Direction2Double>(ds, "quadrature/xcos", dumpdir);
    fdump<Field_Direction2Double>(ds, "quadrature/ycos", dumpdir);
    fdump<Field_Direction2Double>(ds, "quadrature/zcos", dumpdir);
    fdump<Field_Direction2Double>(ds, "quadrature/w", dumpdir);
    fdump<Field_Direction2Int>(ds, "quadrature/id", dumpdir);
    fdump<Field_Direction2Int>(ds, "quadrature/jd", dumpdir);
    fdump<Field_Direction2Int>(ds, "quadrature/kd", dumpdir);
    fdump<Field_ZoneI2Double>(ds, "dx", dumpdir);
    fdump<Field_ZoneJ2Double>(ds, "dy", dumpdir);
    fdump<Field_ZoneK2Double>(ds, "dz", dumpdir);
    fdump<Field_Zone2Double>(ds, "volume", dumpdir);
    fdump<Field_Zone2Int>(ds, "zone_to_num_mixelem", dumpdir);
    fdump<Field_Zone2MixElem>(ds, "zone_to_mixelem", dumpdir);
    fdump<Field_MixElem2Zone>(ds, "mixelem_to_zone", dumpdir);
    fdump<Field_MixElem2Material>(ds, "mixelem_to_material", dumpdir);
    fdump<Field_MixElem2Double>(ds, "mixelem_to_fraction", dumpdir);
    fdump<Field_Moment2Legendre>(ds, "moment_to_legendre", dumpdir);
    fdump<Field_Adjacency>(ds, "upwind", dumpdir);
    fdump<Field_Adjacency>(ds, "downwind", dumpdir);
  }
  return 0;
}
