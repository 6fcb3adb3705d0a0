// Kripke::Kernel entry points (src/Kripke/Kernel.h:19-34): same signatures, same DataStore names,
// but each one gathers the chunk pointers of ALL subdomains it covers into C-ABI descriptors and
// issues one batched launch of the sm_100a kernels (the reference launches one RAJA kernel per
// subdomain).  Results are visible to later kernels through stream order; entry points that
// return a value (population) synchronise.
#include "Kripke.h"

#include <algorithm>
#include <cstring>
#include <map>

using namespace Kripke;
using namespace Kripke::Core;

namespace {
int layoutOf(DataStore &ds) { return (int)ds.getVariable<ArchLayout>("al").al_v.layout_v; }
}  // namespace

// ---- LTimes: src/Kripke/Kernel/LTimes.cpp:79-105 -----------------------------------------------------
void Kripke::Kernel::LTimes(DataStore &data_store) {
  KRIPKE_TIMER(data_store, LTimes);
  auto &pspace = data_store.getVariable<PartitionSpace>("pspace");
  Set const &set_dir = data_store.getVariable<Set>("Set/Direction");
  Set const &set_group = data_store.getVariable<Set>("Set/Group");
  Set const &set_zone = data_store.getVariable<Set>("Set/Zone");
  Set const &set_moment = data_store.getVariable<Set>("Set/Moment");
  auto &field_psi = data_store.getVariable<Field_Flux>("psi");
  auto &field_phi = data_store.getVariable<Field_Moments>("phi");
  auto &field_ell = data_store.getVariable<Field_Ell>("ell");
  const int nq = (int)pspace.getNumSubdomains(SPACE_Q);

  // one descriptor per phi chunk; its direction-set subdomains in ascending id order, KB200_MAX_DIRSETS at a time
  for (int q0 = 0; q0 < nq; q0 += KB200_MAX_DIRSETS) {
    std::vector<kb200_ltimes_desc> descs;
    for (SdomId sdom_phi : field_phi.getWorkList()) {
      auto c = pspace.sdomIdToCoord(sdom_phi);
      kb200_ltimes_desc d;
      memset(&d, 0, sizeof(d));
      d.layout = layoutOf(data_store);
      d.M = (int)set_moment.size(sdom_phi);
      d.Gs = (int)set_group.size(sdom_phi);
      d.Zs = (int)set_zone.size(sdom_phi);
      d.nsets = std::min(KB200_MAX_DIRSETS, nq - q0);
      for (int q = 0; q < d.nsets; ++q) {
        auto cq = c;
        cq[1] = q0 + q;
        SdomId s = pspace.coordToSdomId(cq);
        d.Ds = (int)set_dir.size(s);
        d.ell[q] = field_ell.devicePtrConst(s);
        d.psi[q] = field_psi.devicePtrConst(s);
      }
      // the caller's kConst(phi, 0) (SteadyStateSolver.cpp:53) is folded into the first batch
      d.accumulate = field_phi.consumeZeroPending(sdom_phi) ? 0 : 1;
      d.phi = d.accumulate ? field_phi.devicePtr(sdom_phi) : field_phi.devicePtrOverwrite(sdom_phi);
      descs.push_back(d);
    }
    // descriptors of one call must agree on `accumulate`
    std::vector<kb200_ltimes_desc> a, b;
    for (auto &d : descs) (d.accumulate ? a : b).push_back(d);
    if (!b.empty()) KB200_CALL(kb200_ltimes(b.data(), (int)b.size(), nullptr));
    if (!a.empty()) KB200_CALL(kb200_ltimes(a.data(), (int)a.size(), nullptr));
  }
}

// ---- LPlusTimes: src/Kripke/Kernel/LPlusTimes.cpp:68-92 ------------------------------------------------
void Kripke::Kernel::LPlusTimes(DataStore &data_store) {
  KRIPKE_TIMER(data_store, LPlusTimes);
  auto &pspace = data_store.getVariable<PartitionSpace>("pspace");
  Set const &set_dir = data_store.getVariable<Set>("Set/Direction");
  Set const &set_group = data_store.getVariable<Set>("Set/Group");
  Set const &set_zone = data_store.getVariable<Set>("Set/Zone");
  Set const &set_moment = data_store.getVariable<Set>("Set/Moment");
  auto &field_phi_out = data_store.getVariable<Field_Moments>("phi_out");
  auto &field_rhs = data_store.getVariable<Field_Flux>("rhs");
  auto &field_ell_plus = data_store.getVariable<Field_EllPlus>("ell_plus");
  const int nq = (int)pspace.getNumSubdomains(SPACE_Q);

  for (int q0 = 0; q0 < nq; q0 += KB200_MAX_DIRSETS) {
    std::vector<kb200_lplustimes_desc> a, b;
    for (SdomId sdom_phi : field_phi_out.getWorkList()) {
      auto c = pspace.sdomIdToCoord(sdom_phi);
      kb200_lplustimes_desc d;
      memset(&d, 0, sizeof(d));
      d.layout = layoutOf(data_store);
      d.M = (int)set_moment.size(sdom_phi);
      d.Gs = (int)set_group.size(sdom_phi);
      d.Zs = (int)set_zone.size(sdom_phi);
      d.nsets = std::min(KB200_MAX_DIRSETS, nq - q0);
      d.phi_out = field_phi_out.devicePtrConst(sdom_phi);
      bool all_zero = true;
      for (int q = 0; q < d.nsets; ++q) {
        auto cq = c;
        cq[1] = q0 + q;
        all_zero = all_zero && field_rhs.isZeroPending(pspace.coordToSdomId(cq));
      }
      d.accumulate = all_zero ? 0 : 1;  // kConst(rhs, 0) (SteadyStateSolver.cpp:70) folded when possible
      for (int q = 0; q < d.nsets; ++q) {
        auto cq = c;
        cq[1] = q0 + q;
        SdomId s = pspace.coordToSdomId(cq);
        d.Ds = (int)set_dir.size(s);
        d.ell_plus[q] = field_ell_plus.devicePtrConst(s);
        d.rhs[q] = all_zero ? field_rhs.devicePtrOverwrite(s) : field_rhs.devicePtr(s);
      }
      (d.accumulate ? a : b).push_back(d);
    }
    if (!b.empty()) KB200_CALL(kb200_lplustimes(b.data(), (int)b.size(), nullptr));
    if (!a.empty()) KB200_CALL(kb200_lplustimes(a.data(), (int)a.size(), nullptr));
  }
}

// ---- scattering: src/Kripke/Kernel/Scattering.cpp:112-164 -------------------------------------------------
namespace {
// Kernel::source folded into the scattering epilogue: the write epoch every phi_out chunk had right after a
// scattering call that also added the source.  Kernel::source is a no-op for exactly that state of phi_out.
struct SourceFold {
  const void *field = nullptr;
  std::vector<unsigned long> epochs;
  bool valid = false;
};
SourceFold g_src_fold;
}  // namespace

void Kripke::Kernel::scattering(DataStore &data_store, bool fold_source) {
  KRIPKE_TIMER(data_store, Scattering);
  g_src_fold.valid = false;
  auto &pspace = data_store.getVariable<PartitionSpace>("pspace");
  auto &set_group = data_store.getVariable<Set>("Set/Group");
  auto &set_moment = data_store.getVariable<Set>("Set/Moment");
  auto &set_zone = data_store.getVariable<Set>("Set/Zone");
  auto &set_legendre = data_store.getVariable<Set>("Set/Legendre");
  auto &set_gg = data_store.getVariable<Set>("Set/GlobalGroup");
  auto &field_phi = data_store.getVariable<Field_Moments>("phi");
  auto &field_phi_out = data_store.getVariable<Field_Moments>("phi_out");
  auto &field_sigs = data_store.getVariable<Field_SigmaS>("data/sigs");
  auto &f_z2m = data_store.getVariable<Field_Zone2MixElem>("zone_to_mixelem");
  auto &f_z2n = data_store.getVariable<Field_Zone2Int>("zone_to_num_mixelem");
  auto &f_m2m = data_store.getVariable<Field_MixElem2Material>("mixelem_to_material");
  auto &f_m2f = data_store.getVariable<Field_MixElem2Double>("mixelem_to_fraction");
  auto &f_m2l = data_store.getVariable<Field_Moment2Legendre>("moment_to_legendre");

  std::vector<kb200_scattering_desc> a, b;
  for (auto sdom_dst : field_phi_out.getWorkList()) {
    kb200_scattering_desc d;
    memset(&d, 0, sizeof(d));
    d.layout = layoutOf(data_store);
    d.M = (int)set_moment.size(sdom_dst);
    d.L1 = (int)set_legendre.size(sdom_dst);
    d.G = (int)set_gg.size(sdom_dst);
    d.Gs = (int)set_group.size(sdom_dst);
    d.Zs = (int)set_zone.size(sdom_dst);
    d.glower_dst = (int)set_group.lower(sdom_dst);
    size_t r_dst = pspace.subdomainToSpace(SPACE_R, sdom_dst);
    for (auto sdom_src : field_phi.getWorkList()) {  // ascending source group set (:137-145)
      if (pspace.subdomainToSpace(SPACE_R, sdom_src) != r_dst) continue;
      KRIPKE_ASSERT(d.nsrc < KB200_MAX_DIRSETS, "too many group sets (%d) for one scattering launch\n", d.nsrc + 1);
      d.glower_src[d.nsrc] = (int)set_group.lower(sdom_src);
      d.phi_src[d.nsrc] = field_phi.devicePtrConst(sdom_src);
      d.nsrc++;
    }
    d.sigs = field_sigs.devicePtrConst(sdom_dst);
    d.moment_to_legendre = f_m2l.devicePtrConst(sdom_dst);
    d.zone_to_mixelem = f_z2m.devicePtrConst(sdom_dst);
    d.zone_to_num_mixelem = f_z2n.devicePtrConst(sdom_dst);
    d.mixelem_to_material = f_m2m.devicePtrConst(sdom_dst);
    d.mixelem_to_fraction = f_m2f.devicePtrConst(sdom_dst);
    d.accumulate = field_phi_out.consumeZeroPending(sdom_dst) ? 0 : 1;  // kConst(phi_out,0), SteadyStateSolver.cpp:59
    d.phi_out = d.accumulate ? field_phi_out.devicePtr(sdom_dst) : field_phi_out.devicePtrOverwrite(sdom_dst);
    (d.accumulate ? a : b).push_back(d);
  }
  const char *nf = getenv("KB200_FOLD_SOURCE");
  if (fold_source && !(nf && nf[0] == '0')) {
    int fb = -1, fa = -1;  // -1: no such batch
    if (!b.empty()) KB200_CALL(kb200_scattering_source(b.data(), (int)b.size(), 1.0 /* Kernel/Source.cpp:98 */, &fb, nullptr));
    if (!a.empty()) KB200_CALL(kb200_scattering_source(a.data(), (int)a.size(), 1.0, &fa, nullptr));
    if (fb == -1) fb = fa;
    if (fa == -1) fa = fb;
    if (fb == 1 && fa == 1) {
      g_src_fold.field = &field_phi_out;
      g_src_fold.epochs.clear();
      for (auto sdom : field_phi_out.getWorkList()) g_src_fold.epochs.push_back(field_phi_out.writeEpoch(sdom));
      g_src_fold.valid = true;
    } else if (fb != fa) {
      KRIPKE_ABORT("scattering: the source was folded into one batch of phi_out chunks but not the other\n");
    }
    return;
  }
  if (!b.empty()) KB200_CALL(kb200_scattering(b.data(), (int)b.size(), nullptr));
  if (!a.empty()) KB200_CALL(kb200_scattering(a.data(), (int)a.size(), nullptr));
}

// ---- source: src/Kripke/Kernel/Source.cpp:83-115 ---------------------------------------------------------------
void Kripke::Kernel::source(DataStore &data_store) {
  KRIPKE_TIMER(data_store, Source);
  if (g_src_fold.valid && g_src_fold.field == &data_store.getVariable<Field_Moments>("phi_out")) {
    auto &f = data_store.getVariable<Field_Moments>("phi_out");
    bool same = true;
    size_t k = 0;
    for (auto sdom : f.getWorkList()) same = same && k < g_src_fold.epochs.size() && f.writeEpoch(sdom) == g_src_fold.epochs[k++];
    g_src_fold.valid = false;
    if (same) return;  // the scattering call that produced this phi_out has already added the source
    KRIPKE_ABORT("source: phi_out changed between the scattering that folded the source in and this call\n");
  }
  auto &set_group = data_store.getVariable<Set>("Set/Group");
  auto &set_mixelem = data_store.getVariable<Set>("Set/MixElem");
  auto &set_moment = data_store.getVariable<Set>("Set/Moment");
  auto &set_zone = data_store.getVariable<Set>("Set/Zone");
  auto &field_phi_out = data_store.getVariable<Field_Moments>("phi_out");
  auto &f_m2z = data_store.getVariable<Field_MixElem2Zone>("mixelem_to_zone");
  auto &f_m2m = data_store.getVariable<Field_MixElem2Material>("mixelem_to_material");
  auto &f_m2f = data_store.getVariable<Field_MixElem2Double>("mixelem_to_fraction");
  std::vector<kb200_source_desc> descs;
  for (auto sdom_id : field_phi_out.getWorkList()) {
    kb200_source_desc d;
    memset(&d, 0, sizeof(d));
    d.layout = layoutOf(data_store);
    d.M = (int)set_moment.size(sdom_id);
    d.Gs = (int)set_group.size(sdom_id);
    d.Zs = (int)set_zone.size(sdom_id);
    d.num_mixelem = (int)set_mixelem.size(sdom_id);
    d.strength = 1.0;  // Kernel/Source.cpp:98
    d.mixelem_to_zone = f_m2z.devicePtrConst(sdom_id);
    d.mixelem_to_material = f_m2m.devicePtrConst(sdom_id);
    d.mixelem_to_fraction = f_m2f.devicePtrConst(sdom_id);
    d.phi_out = field_phi_out.devicePtr(sdom_id);
    descs.push_back(d);
  }
  KB200_CALL(kb200_source(descs.data(), (int)descs.size(), nullptr));
}

// ---- population: src/Kripke/Kernel/Population.cpp:74-101 ------------------------------------------------------------
namespace {
// State of the population sum accumulated by the sweep kernels (one per process: one DataStore is solved at a time).
struct PopFusion {
  double *d_partials = nullptr;  // device scratch the sweep launches of one SweepSolver pass append to
  int capacity = 0, used = 0;
  bool collecting = false, complete = false;
  const void *psi = nullptr;               // the field the partials belong to
  std::vector<unsigned long> epochs;       // write epoch of every psi chunk right after it was swept
  std::vector<char> fused;                 // chunk was swept by a kernel that produced partials
};
PopFusion g_pop;
bool popFusionEnabled() {
  const char *e = getenv("KB200_FUSE_POP");
  return !(e && e[0] == '0');
}
}  // namespace

void Kripke::Kernel::populationFusionBegin(DataStore &data_store) {
  g_pop.collecting = false;
  g_pop.complete = false;
  if (!popFusionEnabled()) return;
  auto &field_psi = data_store.getVariable<Field_Flux>("psi");
  if (!g_pop.d_partials) {
    g_pop.capacity = 1 << 16;
    KB200_CALL(kb200_alloc((size_t)g_pop.capacity * sizeof(double), (void **)&g_pop.d_partials));
  }
  const size_t n = data_store.getVariable<PartitionSpace>("pspace").getNumSubdomains(SPACE_PQR);
  g_pop.used = 0;
  g_pop.psi = &field_psi;
  g_pop.epochs.assign(n, 0);
  g_pop.fused.assign(n, 0);
  g_pop.collecting = true;
}

void Kripke::Kernel::populationFusionEnd(DataStore &data_store, std::vector<SdomId> const &swept) {
  if (!g_pop.collecting) return;
  g_pop.collecting = false;
  auto &field_psi = data_store.getVariable<Field_Flux>("psi");
  // complete only if every chunk of psi was swept with partials in this pass, exactly once
  size_t n_fused = 0;
  for (char f : g_pop.fused) n_fused += f ? 1 : 0;
  g_pop.complete = (&field_psi == g_pop.psi) && n_fused == g_pop.fused.size() && swept.size() == g_pop.fused.size();
}

double Kripke::Kernel::population(DataStore &data_store) {
  KRIPKE_TIMER(data_store, Population);
  if (g_pop.complete && g_pop.psi == &data_store.getVariable<Field_Flux>("psi")) {
    auto &psi = data_store.getVariable<Field_Flux>("psi");
    bool valid = true;
    for (SdomId s : psi.getWorkList())
      if ((size_t)*s >= g_pop.epochs.size() || psi.writeEpoch(s) != g_pop.epochs[*s]) { valid = false; break; }
    if (valid) {  // psi is exactly what the sweeps left: finish the sum they accumulated
      static double *d_res = nullptr;
      if (!d_res) KB200_CALL(kb200_alloc(sizeof(double), (void **)&d_res));
      KB200_CALL(kb200_population_reduce(g_pop.d_partials, g_pop.used, d_res, nullptr));
      KB200_CALL(kb200_comm_allreduce_sum_f64(d_res, 1, nullptr));
      double part = 0.0;
      KB200_CALL(kb200_download(&part, d_res, sizeof(double), nullptr));
      KB200_CALL(kb200_stream_sync(nullptr));
      return part;
    }
    g_pop.complete = false;
  }
  Set const &set_dir = data_store.getVariable<Set>("Set/Direction");
  Set const &set_group = data_store.getVariable<Set>("Set/Group");
  Set const &set_zone = data_store.getVariable<Set>("Set/Zone");
  auto &field_psi = data_store.getVariable<Field_Flux>("psi");
  auto &field_w = data_store.getVariable<Field_Direction2Double>("quadrature/w");
  auto &field_volume = data_store.getVariable<Field_Zone2Double>("volume");
  std::vector<kb200_population_desc> descs;
  for (SdomId sdom_id : field_psi.getWorkList()) {
    kb200_population_desc d;
    memset(&d, 0, sizeof(d));
    d.layout = layoutOf(data_store);
    d.Ds = (int)set_dir.size(sdom_id);
    d.Gs = (int)set_group.size(sdom_id);
    d.Zs = (int)set_zone.size(sdom_id);
    d.w = field_w.devicePtrConst(sdom_id);
    d.volume = field_volume.devicePtrConst(sdom_id);
    d.psi = field_psi.devicePtrConst(sdom_id);
    descs.push_back(d);
  }
  static double *d_scratch = nullptr, *d_result = nullptr;
  if (!d_scratch) {
    KB200_CALL(kb200_alloc((kb200_population_scratch_doubles() + 1) * sizeof(double), (void **)&d_scratch));
    d_result = d_scratch + kb200_population_scratch_doubles();
  }
  KB200_CALL(kb200_population(descs.data(), (int)descs.size(), d_scratch, d_result, nullptr));
  // reduce over ranks (Core/Comm.h:161-166) on the device, then one 8-byte read-back
  KB200_CALL(kb200_comm_allreduce_sum_f64(d_result, 1, nullptr));
  double part = 0.0;
  KB200_CALL(kb200_download(&part, d_result, sizeof(double), nullptr));
  KB200_CALL(kb200_stream_sync(nullptr));
  return part;
}

// ---- sweepSubdomain(s): src/Kripke/Kernel/SweepSubdomain.cpp:115-124 ---------------------------------------------------
void Kripke::Kernel::sweepSubdomains(DataStore &data_store, std::vector<SdomId> const &sdom_ids, bool deliver_downwind,
                                     bool fuse_population) {
  if (sdom_ids.empty()) return;
  KRIPKE_TIMER(data_store, SweepSubdomain);
  auto &pspace = data_store.getVariable<PartitionSpace>("pspace");
  Comm comm;
  auto &set_dir = data_store.getVariable<Set>("Set/Direction");
  auto &set_group = data_store.getVariable<Set>("Set/Group");
  auto &set_zi = data_store.getVariable<Set>("Set/ZoneI");
  auto &set_zj = data_store.getVariable<Set>("Set/ZoneJ");
  auto &set_zk = data_store.getVariable<Set>("Set/ZoneK");
  auto &f_xcos = data_store.getVariable<Field_Direction2Double>("quadrature/xcos");
  auto &f_ycos = data_store.getVariable<Field_Direction2Double>("quadrature/ycos");
  auto &f_zcos = data_store.getVariable<Field_Direction2Double>("quadrature/zcos");
  auto &f_id = data_store.getVariable<Field_Direction2Int>("quadrature/id");
  auto &f_jd = data_store.getVariable<Field_Direction2Int>("quadrature/jd");
  auto &f_kd = data_store.getVariable<Field_Direction2Int>("quadrature/kd");
  auto &f_dx = data_store.getVariable<Field_ZoneI2Double>("dx");
  auto &f_dy = data_store.getVariable<Field_ZoneJ2Double>("dy");
  auto &f_dz = data_store.getVariable<Field_ZoneK2Double>("dz");
  auto &f_sigt = data_store.getVariable<Field_SigmaTZonal>("sigt_zonal");
  auto &f_psi = data_store.getVariable<Field_Flux>("psi");
  auto &f_rhs = data_store.getVariable<Field_Flux>("rhs");
  FieldStorage<double> *planes[3] = {&data_store.getVariable<Field_IPlane>("i_plane"), &data_store.getVariable<Field_JPlane>("j_plane"),
                                     &data_store.getVariable<Field_KPlane>("k_plane")};
  auto &f_down = data_store.getVariable<Field_Adjacency>("downwind");

  PeerExchange *px = (deliver_downwind && comm.size() > 1) ? PeerExchange::get(data_store) : nullptr;
  if (px && !px->usable()) px = nullptr;
  std::vector<kb200_sweep_desc> descs;
  for (SdomId s : sdom_ids) {
    kb200_sweep_desc d;
    memset(&d, 0, sizeof(d));
    d.layout = layoutOf(data_store);
    d.Ds = (int)set_dir.size(s);
    d.Gs = (int)set_group.size(s);
    d.ni = (int)set_zi.size(s);
    d.nj = (int)set_zj.size(s);
    d.nk = (int)set_zk.size(s);
    d.id = f_id.getDataConst(s)[0];  // direction 0 of the set (SweepSubdomain.cpp:61-72)
    d.jd = f_jd.getDataConst(s)[0];
    d.kd = f_kd.getDataConst(s)[0];
    d.xcos = f_xcos.devicePtrConst(s);
    d.ycos = f_ycos.devicePtrConst(s);
    d.zcos = f_zcos.devicePtrConst(s);
    d.dx = f_dx.devicePtrConst(s);
    d.dy = f_dy.devicePtrConst(s);
    d.dz = f_dz.devicePtrConst(s);
    d.sigt = f_sigt.devicePtrConst(s);
    d.rhs = f_rhs.devicePtrConst(s);
    d.psi = f_psi.devicePtrOverwrite(s);  // psi is fully overwritten (SweepSubdomain.cpp:101)
    const long *down = f_down.getDataConst(s);
    for (int dim = 0; dim < 3; ++dim) {
      // a pending zero-fill of the plane chunk (SweepSolver.cpp:28-30,64-72) becomes "vacuum inflow"
      d.inflow_zero[dim] = planes[dim]->consumeZeroPending(s) ? 1 : 0;
      double *p = d.inflow_zero[dim] ? planes[dim]->devicePtrOverwrite(s) : planes[dim]->devicePtr(s);
      (dim == 0 ? d.i_plane : dim == 1 ? d.j_plane : d.k_plane) = p;
      d.out_plane[dim] = nullptr;
      if (deliver_downwind && down[dim] >= 0 && pspace.globalSdomIdToRank(GlobalSdomId(down[dim])) == (int)comm.rank()) {
        SdomId sd = pspace.globalSdomIdToSdomId(GlobalSdomId(down[dim]));
        d.out_plane[dim] = planes[dim]->devicePtrOverwrite(sd);  // folds ParallelComm::postSends' on-rank copy
      } else if (px && down[dim] >= 0) {
        d.out_plane[dim] = px->outPlane(s, dim);  // the downwind chunk on another GPU, mapped over NVLink
      }
    }
    descs.push_back(d);
  }
  if (fuse_population && g_pop.collecting && g_pop.psi == &f_psi) {
    auto &f_w = data_store.getVariable<Field_Direction2Double>("quadrature/w");
    auto &f_vol = data_store.getVariable<Field_Zone2Double>("volume");
    std::vector<const double *> pw, pv;
    std::vector<double> vc;
    for (SdomId s : sdom_ids) {
      pw.push_back(f_w.devicePtrConst(s));
      // > 0: every zone of the chunk has this volume (all the generator produces, Generate/Space.cpp:110-136)
      vc.push_back(f_vol.uniformPositiveValue(s));
      pv.push_back(f_vol.devicePtrConst(s));
    }
    int count = 0;
    KB200_CALL(kb200_sweep_population_uniform(descs.data(), (int)descs.size(), pw.data(), pv.data(), vc.data(),
                                              g_pop.d_partials + g_pop.used, g_pop.capacity - g_pop.used, &count, nullptr));
    if (count > 0) {
      g_pop.used += count;
      for (SdomId s : sdom_ids) {
        g_pop.fused[*s] = g_pop.fused[*s] ? 2 : 1;  // 2 = swept twice: the partials would double count
        g_pop.epochs[*s] = f_psi.writeEpoch(s);
      }
      for (char &f : g_pop.fused) if (f == 2) { g_pop.collecting = false; }
    } else {
      g_pop.collecting = false;  // some subdomains went through a kernel without the fused sum
    }
    return;
  }
  KB200_CALL(kb200_sweep(descs.data(), (int)descs.size(), nullptr));
}

void Kripke::Kernel::sweepSubdomain(DataStore &data_store, SdomId sdom_id) {
  sweepSubdomains(data_store, std::vector<SdomId>{sdom_id}, false);
}
