// SteadyStateSolver, Timing, InputVariables and command-line parsing of the host layer
// (src/Kripke/SteadyStateSolver.cpp, Timing.cpp, InputVariables.cpp, src/kripke.cpp:297-390).
#include "Kripke.h"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <sstream>

using namespace Kripke;
using namespace Kripke::Core;

// ---- SteadyStateSolver (src/Kripke/SteadyStateSolver.cpp:24-115) ------------------------------------
static std::vector<double> g_last_particles;
std::vector<double> const &Kripke::lastParticleCounts() { return g_last_particles; }

int Kripke::SteadyStateSolver(DataStore &data_store, size_t max_iter, bool block_jacobi) {
  KRIPKE_TIMER(data_store, Solve);
  PartitionSpace &pspace = data_store.getVariable<PartitionSpace>("pspace");
  Comm const &comm = data_store.getVariable<Comm>("comm");
  if (comm.rank() == 0) {
    printf("\n");
    printf("Steady State Solve\n");
    printf("==================\n\n");
  }
  g_last_particles.clear();

  // Initialize unknowns
  Kernel::kConst(data_store.getVariable<Field_Flux>("psi"), 0.0);

  double part_last = 0.0;
  for (size_t iter = 0; iter < max_iter; ++iter) {
    // rhs = LPlus*S*L*psi + Q
    Kernel::kConst(data_store.getVariable<Field_Moments>("phi"), 0.0);
    Kernel::LTimes(data_store);

    Kernel::kConst(data_store.getVariable<Field_Moments>("phi_out"), 0.0);
    Kernel::scattering(data_store, true);  // Kernel::source below is folded into its epilogue

    Kernel::source(data_store);

    Kernel::kConst(data_store.getVariable<Field_Flux>("rhs"), 0.0);
    Kernel::LPlusTimes(data_store);

    // psi = Hinv*rhs
    {
      int num_subdomains = (int)pspace.getNumSubdomains(SPACE_PQR);
      std::vector<SdomId> sdom_list(num_subdomains);
      for (int i = 0; i < num_subdomains; ++i) sdom_list[i] = SdomId(i);
      SweepSolver(data_store, sdom_list, block_jacobi);
    }

    double part = Kernel::population(data_store);
    g_last_particles.push_back(part);
    if (comm.rank() == 0) {
      printf("  iter %d: particle count=%e, change=%e\n", (int)iter, part, (part - part_last) / part);
      fflush(stdout);
    }
    part_last = part;
  }
  if (comm.rank() == 0) printf("  Solver terminated\n");
  return 0;
}

// ---- Timing (src/Kripke/Timing.cpp) --------------------------------------------------------------------
bool Timing::s_sync = true;
static double wallSeconds() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
Timing::~Timing() { stopAll(); }
void Timing::start(std::string const &name) {
  Timer &t = timers[name];
  if (s_sync) kb200_device_sync();  // no-op before the device is initialised
  t.t0 = wallSeconds();
  t.started = true;
  ++t.count;
  m_stack.push_back(name);
  std::string path;
  for (auto &n : m_stack) path += (path.empty() ? "" : "/") + n;
  if (!m_region_seconds.count(path)) { m_region_seconds[path] = 0.0; m_region_order.push_back(path); }
  m_open.emplace_back(path, t.t0);
}
void Timing::stop(std::string const &name) {
  Timer &t = timers[name];
  if (!t.started) return;
  if (s_sync) kb200_device_sync();
  const double now = wallSeconds();
  t.elapsed += now - t.t0;
  t.started = false;
  // close the innermost open region of that name (and anything left open inside it)
  for (size_t k = m_stack.size(); k-- > 0;)
    if (m_stack[k] == name) {
      for (size_t q = m_stack.size(); q-- > k;) m_region_seconds[m_open[q].first] += now - m_open[q].second;
      m_stack.resize(k);
      m_open.resize(k);
      break;
    }
}
void Timing::stopAll(void) {
  for (auto &kv : timers) stop(kv.first);
}
double Timing::getTotal(std::string const &name) const {
  auto i = timers.find(name);
  return i == timers.end() ? 0.0 : i->second.elapsed;
}
size_t Timing::getCount(std::string const &name) const {
  auto i = timers.find(name);
  return i == timers.end() ? 0 : i->second.count;
}
void Timing::print(void) const {
  Comm default_comm;
  if (default_comm.rank() != 0) return;
  std::vector<std::string> names;
  for (auto &kv : timers) names.push_back(kv.first);
  std::sort(names.begin(), names.end());
  printf("\nTimers\n======\n\n");
  printf("  %-16s  %12s  %12s\n", "Timer", "Count", "Seconds");
  printf("  ----------------  ------------  ------------\n");
  for (auto &n : names) printf("  %-16s  %12d  %12.5lf\n", n.c_str(), (int)getCount(n), getTotal(n));
  printf("\nTIMER_NAMES:");
  for (size_t i = 0; i < names.size(); ++i) printf("%s%s", i ? "," : "", names[i].c_str());
  printf("\nTIMER_DATA:");
  for (size_t i = 0; i < names.size(); ++i) printf("%s%lf", i ? "," : "", getTotal(names[i]));
  printf("\n");
  const char *c1 = getenv("CALI_CONFIG"), *c2 = getenv("CALI_CONFIG_PROFILE");
  if ((c1 && strstr(c1, "runtime-report")) || (c2 && strstr(c2, "runtime-report"))) printRegions();
}
// the table Caliper's runtime-report prints for the reference's nested regions: one row per call path, indented by depth
void Timing::printRegions(void) const {
  double total = 0.0;
  for (auto &p : m_region_order)
    if (p.find('/') == std::string::npos) total += m_region_seconds.at(p);
  printf("\n%-36s %13s %13s %13s %7s\n", "Path", "Min time/rank", "Max time/rank", "Avg time/rank", "Time %");
  std::vector<std::string> order = m_region_order;
  std::sort(order.begin(), order.end());  // parents before children, siblings by name
  for (auto &p : order) {
    const size_t depth = std::count(p.begin(), p.end(), '/');
    const std::string leaf = p.substr(p.rfind('/') == std::string::npos ? 0 : p.rfind('/') + 1);
    const double s = m_region_seconds.at(p);
    // exclusive share of the run, as runtime-report shows it
    double child = 0.0;
    for (auto &q : m_region_order)
      if (q.size() > p.size() && q.compare(0, p.size(), p) == 0 && q[p.size()] == '/' && q.find('/', p.size() + 1) == std::string::npos)
        child += m_region_seconds.at(q);
    const std::string label = std::string(2 * depth, ' ') + leaf;
    printf("%-36s %13.6f %13.6f %13.6f %7.3f\n", label.c_str(), s, s, s, total > 0.0 ? 100.0 * (s - child) / total : 0.0);
  }
}

// ---- InputVariables (src/Kripke/InputVariables.cpp) -------------------------------------------------------
InputVariables::InputVariables()
    : nx(16), ny(16), nz(16), num_directions(96), num_groups(32), legendre_order(4), quad_num_polar(0), quad_num_azimuthal(0),
      al_v(ArchLayoutV{ArchV_B200, LayoutV_DGZ}), npx(1), npy(1), npz(1), num_dirsets(8), num_groupsets(2), niter(10),
      parallel_method(PMETHOD_SWEEP), num_material_subsamples(4), run_name("kripke") {
  num_zonesets_dim[0] = num_zonesets_dim[1] = num_zonesets_dim[2] = 1;
  sigt[0] = 0.1; sigt[1] = 0.0001; sigt[2] = 0.1;
  sigs[0] = 0.05; sigs[1] = 0.00005; sigs[2] = 0.05;
}

bool InputVariables::checkValues(void) const {
  Comm comm;
  const bool root = comm.rank() == 0;
  auto fail = [&](const char *fmt, int a = 0, int b = 0) {
    if (root) printf(fmt, a, b);
    return true;
  };
  if (num_zonesets_dim[0] <= 0 || num_zonesets_dim[1] <= 0 || num_zonesets_dim[2] <= 0)
    return fail("Number of zone-sets in each dim need to be greater than or equal to 1\n");
  if (num_groups < 1) return fail("Number of groups (%d) needs to be at least 1\n", num_groups);
  if (num_groups % num_groupsets)
    return fail("Number of groups (%d) must be evenly divided by number of groupsets (%d)\n", num_groups, num_groupsets);
  if (num_directions < 8) return fail("Number of directions (%d) needs to be at least 8\n", num_directions);
  if (num_dirsets % 8 && num_dirsets < 8) return fail("Number of direction sets (%d) must be a multiple of 8\n", num_dirsets);
  if (num_directions % num_dirsets)
    return fail("Number of directions (%d) must be evenly divided by number of directionsets(%d)\n", num_directions, num_dirsets);
  if (legendre_order < 0) return fail("Legendre scattering order (%d) must be >= 0\n", legendre_order);
  if (niter < 1) return fail("You must run at least one iteration (%d)\n", niter);
  // additional checks the reference leaves to later asserts / undefined behaviour
  if (al_v.arch_v == ArchV_Unknown) return fail("Unknown --arch: this build runs on the B200 only (accepted: B200, CUDA); there is no CPU path\n");
  if (al_v.layout_v == LayoutV_Unknown) return fail("Unknown --layout (available: DGZ,DZG,GDZ,GZD,ZDG,ZGD)\n");
  if (num_groupsets > KB200_MAX_DIRSETS) return fail("Number of groupsets (%d) is limited to %d\n", num_groupsets, KB200_MAX_DIRSETS);
  return false;
}

// ---- command line (src/kripke.cpp:43-162,297-390) ---------------------------------------------------------------
void printUsage() {
  Comm comm;
  if (comm.rank() != 0) return;
  InputVariables def;
  printf("Usage:  [torchrun ...] kripke [options...]\n\n");
  printf("Problem Size Options:\n---------------------\n");
  printf("  --groups <ngroups>     Number of energy groups\n                         Default:  --groups %d\n\n", def.num_groups);
  printf("  --legendre <lorder>    Scattering Legendre Expansion Order (0, 1, ...)\n                         Default:  --legendre %d\n\n", def.legendre_order);
  printf("  --quad [<ndirs>|<polar>:<azim>]\n                         Define the quadrature set to use\n"
         "                         Either a fake S2 with <ndirs> points,\n                         OR Gauss-Legendre with <polar> by <azim> points\n"
         "                         Default:  --quad %d\n\n", def.num_directions);
  printf("  --zones <x,y,z>        Number of zones in x,y,z\n                         Default:  --zones %d,%d,%d\n\n", def.nx, def.ny, def.nz);
  printf("\nPhysics Parameters:\n-------------------\n");
  printf("  --sigt <st0,st1,st2>   Total material cross-sections\n                         Default:   --sigt %lf,%lf,%lf\n\n", def.sigt[0], def.sigt[1], def.sigt[2]);
  printf("  --sigs <ss0,ss1,ss2>   Scattering material cross-sections\n                         Default:   --sigs %lf,%lf,%lf\n\n", def.sigs[0], def.sigs[1], def.sigs[2]);
  printf("\nOn-Node Options:\n----------------\n");
  printf("  --arch <ARCH>          Architecture selection\n                         Available: B200 (alias: CUDA)\n                         Default:   --arch %s\n\n", archToString(def.al_v.arch_v).c_str());
  printf("  --layout <LAYOUT>      Data layout and loop nesting order\n                         Available: DGZ,DZG,GDZ,GZD,ZDG,ZGD\n                         Default:   --layout %s\n\n", layoutToString(def.al_v.layout_v).c_str());
  printf("\nParallel Decomposition Options:\n-------------------------------\n");
  printf("  --procs <npx,npy,npz>  Number of ranks (one GPU each) in each spatial dimension\n                         Default:  --procs %d,%d,%d\n\n", def.npx, def.npy, def.npz);
  printf("  --dset <ds>            Number of direction-sets\n                         Must be a factor of 8, and divide evenly the number\n"
         "                         of quadrature points\n                         Default:  --dset %d\n\n", def.num_dirsets);
  printf("  --gset <gs>            Number of energy group-sets\n                         Must divide evenly the number energy groups\n                         Default:  --gset %d\n\n", def.num_groupsets);
  printf("  --zset <zx>,<zy>,<zz>  Number of zone-sets in x,y, and z\n                         Default:  --zset %d,%d,%d\n\n", def.num_zonesets_dim[0], def.num_zonesets_dim[1], def.num_zonesets_dim[2]);
  printf("\nSolver Options:\n---------------\n");
  printf("  --niter <NITER>        Number of solver iterations to run\n                         Default:  --niter %d\n\n", def.niter);
  printf("  --pmethod <method>     Parallel solver method\n                         sweep: Full up-wind sweep (wavefront algorithm)\n"
         "                         bj: Block Jacobi\n                         Default: --pmethod sweep\n\n\n");
}

static std::vector<std::string> splitString(std::string const &str, char delim) {
  std::vector<std::string> elem;
  std::stringstream ss(str);
  std::string e;
  while (std::getline(ss, e, delim)) elem.push_back(e);
  return elem;
}

int parseCommandLine(int argc, char **argv, InputVariables &vars) {
  std::vector<std::string> args(argv + 1, argv + argc);
  size_t cur = 0;
  bool bad = false;
  auto pop = [&]() -> std::string {
    if (cur >= args.size()) { bad = true; return std::string(); }
    return args[cur++];
  };
  auto triple = [&](int *out) {
    auto v = splitString(pop(), ',');
    if (v.size() != 3) { bad = true; return; }
    for (int i = 0; i < 3; ++i) out[i] = std::atoi(v[i].c_str());
  };
  auto dtriple = [&](double *out) {
    auto v = splitString(pop(), ',');
    if (v.size() != 3) { bad = true; return; }
    for (int i = 0; i < 3; ++i) out[i] = std::atof(v[i].c_str());
  };
  while (cur < args.size() && !bad) {
    std::string opt = pop();
    if (opt == "-h" || opt == "--help") bad = true;
    else if (opt == "--name") vars.run_name = pop();
    else if (opt == "--dset") vars.num_dirsets = std::atoi(pop().c_str());
    else if (opt == "--gset") vars.num_groupsets = std::atoi(pop().c_str());
    else if (opt == "--zset") triple(vars.num_zonesets_dim);
    else if (opt == "--zones") { int z[3] = {vars.nx, vars.ny, vars.nz}; triple(z); vars.nx = z[0]; vars.ny = z[1]; vars.nz = z[2]; }
    else if (opt == "--procs") { int p[3] = {vars.npx, vars.npy, vars.npz}; triple(p); vars.npx = p[0]; vars.npy = p[1]; vars.npz = p[2]; }
    else if (opt == "--pmethod") {
      std::string m = pop();
      if (!strcasecmp(m.c_str(), "sweep")) vars.parallel_method = PMETHOD_SWEEP;
      else if (!strcasecmp(m.c_str(), "bj")) vars.parallel_method = PMETHOD_BJ;
      else bad = true;
    }
    else if (opt == "--groups") vars.num_groups = std::atoi(pop().c_str());
    else if (opt == "--quad") {
      auto p = splitString(pop(), ':');
      if (p.size() == 1) { vars.num_directions = std::atoi(p[0].c_str()); vars.quad_num_polar = vars.quad_num_azimuthal = 0; }
      else if (p.size() == 2) {
        vars.quad_num_polar = std::atoi(p[0].c_str());
        vars.quad_num_azimuthal = std::atoi(p[1].c_str());
        vars.num_directions = vars.quad_num_polar * vars.quad_num_azimuthal;
      } else bad = true;
    }
    else if (opt == "--legendre") vars.legendre_order = std::atoi(pop().c_str());
    else if (opt == "--sigs") dtriple(vars.sigs);
    else if (opt == "--sigt") dtriple(vars.sigt);
    else if (opt == "--niter") vars.niter = std::atoi(pop().c_str());
    else if (opt == "--arch") vars.al_v.arch_v = stringToArch(pop());
    else if (opt == "--layout") vars.al_v.layout_v = stringToLayout(pop());
    else { printf("Unknwon options %s\n", opt.c_str()); bad = true; }
  }
  if (bad) { printUsage(); return 1; }
  return 0;
}
