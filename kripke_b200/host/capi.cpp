// Flat C entry points over the host layer so that Python (ctypes: tests/, bench.py, smoke) can
// drive the same Kripke:: objects the CLI uses.  No compute happens here.
#include "Kripke.h"

#include <cstring>

using namespace Kripke;
using namespace Kripke::Core;

namespace {
struct Handle {
  InputVariables vars;
  DataStore *ds = nullptr;
};
FieldStorageBase &fieldOf(Handle *h, const char *name) { return h->ds->getVariable<FieldStorageBase>(name); }
}  // namespace

extern "C" {

void kripke_b200_set_world(int rank, int size) { Comm::setWorld(rank, size); }
void kripke_b200_timer_sync(int on) { Timing::s_sync = on != 0; }

// parses a kripke command line (argv[0] ignored) and generates the problem (host only)
void *kripke_b200_create(int argc, char **argv) {
  Handle *h = new Handle();
  if (parseCommandLine(argc, argv, h->vars) || h->vars.checkValues()) {
    delete h;
    return nullptr;
  }
  h->ds = new DataStore();
  generateProblem(*h->ds, h->vars);
  return h;
}
void kripke_b200_destroy(void *hv) {
  Handle *h = (Handle *)hv;
  if (!h) return;
  kb200_device_sync();
  delete h->ds;
  delete h;
}
int kripke_b200_niter(void *hv) { return ((Handle *)hv)->vars.niter; }
int kripke_b200_is_bj(void *hv) { return ((Handle *)hv)->vars.parallel_method == PMETHOD_BJ; }

// SteadyStateSolver(data_store, niter, bj); particles[niter] receives the full-precision counts
int kripke_b200_solve(void *hv, int niter, double *particles) {
  Handle *h = (Handle *)hv;
  int rc = SteadyStateSolver(*h->ds, (size_t)niter, h->vars.parallel_method == PMETHOD_BJ);
  auto const &p = lastParticleCounts();
  if (particles)
    for (int i = 0; i < niter && i < (int)p.size(); ++i) particles[i] = p[i];
  return rc;
}

// single entry points by name: LTimes LPlusTimes scattering source SweepSolver population
// zero:<field> (kConst(field,0))
int kripke_b200_call(void *hv, const char *what, double *result) {
  Handle *h = (Handle *)hv;
  DataStore &ds = *h->ds;
  std::string w(what);
  if (w == "LTimes") Kernel::LTimes(ds);
  else if (w == "LPlusTimes") Kernel::LPlusTimes(ds);
  else if (w == "scattering") Kernel::scattering(ds);
  else if (w == "source") Kernel::source(ds);
  else if (w == "population") { double p = Kernel::population(ds); if (result) *result = p; }
  else if (w == "SweepSolver") {
    auto &pspace = ds.getVariable<PartitionSpace>("pspace");
    std::vector<SdomId> l;
    for (size_t i = 0; i < pspace.getNumSubdomains(SPACE_PQR); ++i) l.push_back(SdomId((long)i));
    SweepSolver(ds, l, h->vars.parallel_method == PMETHOD_BJ);
  }
  else if (w.rfind("sweepSubdomain:", 0) == 0) Kernel::sweepSubdomain(ds, SdomId(atol(w.c_str() + 15)));
  else if (w.rfind("zero:", 0) == 0) {
    FieldStorageBase &f = fieldOf(h, w.c_str() + 5);
    for (SdomId s : f.getWorkList()) f.setZeroPending(s);
  }
  else return 1;
  return 0;
}

int kripke_b200_field_chunks(void *hv, const char *name) {
  try { return (int)fieldOf((Handle *)hv, name).getNumChunks(); } catch (std::exception &) { return -1; }
}
long kripke_b200_field_chunk_size(void *hv, const char *name, int chunk, int *elem_size) {
  FieldStorageBase &f = fieldOf((Handle *)hv, name);
  if (elem_size) *elem_size = (int)f.elemSize();
  return (long)f.size(f.getWorkList()[chunk]);
}
// copies a chunk to / from caller memory (device -> host download if the device copy is newer)
int kripke_b200_field_get(void *hv, const char *name, int chunk, void *out) {
  FieldStorageBase &f = fieldOf((Handle *)hv, name);
  SdomId s = f.getWorkList()[chunk];
  memcpy(out, f.hostPtr(s, false), f.size(s) * f.elemSize());
  return 0;
}
int kripke_b200_field_set(void *hv, const char *name, int chunk, const void *in) {
  FieldStorageBase &f = fieldOf((Handle *)hv, name);
  SdomId s = f.getWorkList()[chunk];
  memcpy(f.hostPtr(s, true), in, f.size(s) * f.elemSize());
  return 0;
}
// device pointer of a chunk (uploads pending host data); for e2e staging and external kernels
void *kripke_b200_field_device_ptr(void *hv, const char *name, int chunk, int will_write) {
  FieldStorageBase &f = fieldOf((Handle *)hv, name);
  return f.devPtr(f.getWorkList()[chunk], will_write != 0);
}
void kripke_b200_release_host_mirrors(void *hv, const char *name) { fieldOf((Handle *)hv, name).releaseHostMirrors(); }

double kripke_b200_timer_total(void *hv, const char *name) { return ((Handle *)hv)->ds->getVariable<Timing>("timing").getTotal(name); }
long kripke_b200_timer_count(void *hv, const char *name) { return (long)((Handle *)hv)->ds->getVariable<Timing>("timing").getCount(name); }

long kripke_b200_num_unknowns(void *hv) {
  DataStore &ds = *((Handle *)hv)->ds;
  return (long)(ds.getVariable<Set>("Set/Group").globalSize() * ds.getVariable<Set>("Set/Direction").globalSize() *
                ds.getVariable<Set>("Set/Zone").globalSize());
}
int kripke_b200_num_subdomains(void *hv) { return (int)((Handle *)hv)->ds->getVariable<PartitionSpace>("pspace").getNumSubdomains(SPACE_PQR); }

// host-only schedule of one SweepSolver call for this rank: visit order of the local subdomains
// (order[n]), the wavefront stage of each visit (stage[n]) and, per visited subdomain and
// dimension, the peer rank a face is received from / sent to (-1 = none or on-rank):
// recv_from[3n], send_to[3n].  Needs no GPU; used by the world_size>1 CPU tests.
int kripke_b200_sweep_schedule(void *hv, int *order, int *stage, int *recv_from, int *send_to) {
  Handle *h = (Handle *)hv;
  DataStore &ds = *h->ds;
  auto &pspace = ds.getVariable<PartitionSpace>("pspace");
  Comm comm;
  const int n = (int)pspace.getNumSubdomains(SPACE_PQR);
  auto &f_up = ds.getVariable<Field_Adjacency>("upwind");
  auto &f_down = ds.getVariable<Field_Adjacency>("downwind");
  const int nstages = numStages(pspace);  // the planner of host/sweep_solver.cpp, not a copy of it
  int k = 0;
  for (int st = 0; st < nstages; ++st)
    for (int s = 0; s < n; ++s) {
      SdomId sdom(s);
      int depth = sweepDepth(ds, pspace, sdom);
      if (h->vars.parallel_method == PMETHOD_BJ) depth = 0;
      if (depth != st) continue;
      order[k] = s;
      stage[k] = st;
      const long *up = f_up.getDataConst(sdom), *down = f_down.getDataConst(sdom);
      for (int dim = 0; dim < 3; ++dim) {
        int pr = up[dim] >= 0 ? pspace.globalSdomIdToRank(GlobalSdomId(up[dim])) : -1;
        int ps = down[dim] >= 0 ? pspace.globalSdomIdToRank(GlobalSdomId(down[dim])) : -1;
        recv_from[3 * k + dim] = (pr == (int)comm.rank()) ? -1 : pr;
        send_to[3 * k + dim] = (ps == (int)comm.rank()) ? -1 : ps;
      }
      ++k;
    }
  return k;
}

}  // extern "C"
