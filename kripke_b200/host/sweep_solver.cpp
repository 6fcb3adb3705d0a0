// SweepSolver and the ParallelComm family (src/Kripke/SweepSolver.cpp, ParallelComm.cpp,
// ParallelComm/{SweepComm,BlockJacobiComm}.cpp) for one process per B200.
//
// What changed relative to the reference, behind the same interface:
//  * The reference sweeps ONE ready subdomain at a time (SweepSolver.cpp:58-59).  Here every
//    subdomain of the same wavefront stage is swept by one batched launch.  A subdomain's stage is
//    its depth in the global upwind DAG (sum over x,y,z of its zone-set distance from the inflow
//    boundary of its octant), so stage t only needs faces produced in stage t-1.  Any topological
//    order gives bit-identical psi, so results do not change.
//  * Faces to an on-rank downwind subdomain are written by the sweep kernel itself into the
//    neighbour's plane chunk (kb200_sweep_desc.out_plane), which replaces postSends' host copy loop
//    (ParallelComm.cpp:149-161).  Faces to another GPU go through NCCL send/recv over NVLink, one
//    grouped exchange per stage; message order inside a group is (receiver global subdomain id,
//    dimension) on both sides, the NCCL analogue of the reference's MPI tag (ParallelComm.cpp:106,176).
//  * KB200_SWEEP_ORDER=reference restores the reference's serial "first ready subdomain in queue
//    order" schedule (single rank only) so the visit order fixture of SURVEY 3.3 can be checked.
#include "Kripke.h"

#include <algorithm>
#include <cstring>

using namespace Kripke;
using namespace Kripke::Core;

namespace Kripke {
// (declared in Kripke.h: the C API's schedule query and the solver below share these)
// depth of a global subdomain in the sweep DAG of its octant
int sweepDepth(DataStore &ds, PartitionSpace const &pspace, SdomId local) {
  auto gc = pspace.coordToGlobalCoord(pspace.sdomIdToCoord(local));
  const int dir[3] = {ds.getVariable<Field_Direction2Int>("quadrature/id").getDataConst(local)[0],
                      ds.getVariable<Field_Direction2Int>("quadrature/jd").getDataConst(local)[0],
                      ds.getVariable<Field_Direction2Int>("quadrature/kd").getDataConst(local)[0]};
  int depth = 0;
  for (int dim = 0; dim < 3; ++dim) {
    const long n = (long)pspace.getGlobalNumSubdomains((SPACE)(SPACE_RX + dim));
    depth += (dir[dim] > 0) ? (int)gc[SPACE_RX + dim] : (int)(n - 1 - gc[SPACE_RX + dim]);
  }
  return depth;
}

int numStages(PartitionSpace const &pspace) {
  return (int)(pspace.getGlobalNumSubdomains(SPACE_RX) + pspace.getGlobalNumSubdomains(SPACE_RY) +
               pspace.getGlobalNumSubdomains(SPACE_RZ)) - 2;
}
}  // namespace Kripke

namespace {

bool referenceOrder() {
  const char *e = getenv("KB200_SWEEP_ORDER");
  return e && !strcasecmp(e, "reference");
}

struct Message {
  long receiver_global;  // global id of the receiving subdomain (the reference's MPI tag)
  int dim;
  int peer;
  bool is_send;
  double *buf;
  size_t count;
};

// one grouped NCCL exchange; both sides order their messages identically per peer
void runExchange(std::vector<Message> &msgs) {
  if (msgs.empty()) return;
  std::sort(msgs.begin(), msgs.end(), [](Message const &a, Message const &b) {
    if (a.receiver_global != b.receiver_global) return a.receiver_global < b.receiver_global;
    return a.dim < b.dim;
  });
  KB200_CALL(kb200_comm_group_start());
  for (auto &m : msgs) {
    if (m.is_send) KB200_CALL(kb200_comm_send(m.buf, m.count, m.peer, nullptr));
    else KB200_CALL(kb200_comm_recv(m.buf, m.count, m.peer, nullptr));
  }
  KB200_CALL(kb200_comm_group_end());
}

}  // namespace

// ---- PeerExchange ------------------------------------------------------------------------------------------------------
namespace {
struct ExportEntry {  // one plane chunk this rank lets a peer write
  long gid;           // global id of the receiving subdomain
  int dim, pad;
  unsigned char handle[64];
};
bool peerExchangeWanted() {
  const char *e = getenv("KB200_P2P");
  return !(e && e[0] == '0');
}
}  // namespace

PeerExchange *PeerExchange::get(DataStore &ds) {
  if (ds.isVariableType<PeerExchange>("kb200/p2p")) return &ds.getVariable<PeerExchange>("kb200/p2p");
  PeerExchange *px = new PeerExchange();
  ds.addVariable("kb200/p2p", px);
  Comm comm;
  if (comm.size() <= 1 || !peerExchangeWanted()) return px;
  auto &pspace = ds.getVariable<PartitionSpace>("pspace");
  FieldStorage<double> *planes[3] = {&ds.getVariable<Field_IPlane>("i_plane"), &ds.getVariable<Field_JPlane>("j_plane"),
                                     &ds.getVariable<Field_KPlane>("k_plane")};
  auto &f_up = ds.getVariable<Field_Adjacency>("upwind");
  auto &f_down = ds.getVariable<Field_Adjacency>("downwind");
  const long *l2g = ds.getVariable<Field_SdomId2GlobalSdomId>("SdomId2GlobalSdomId").getDataConst(SdomId(0));
  const size_t nlocal = pspace.getNumSubdomains(SPACE_PQR);
  const int me = (int)comm.rank(), nranks = (int)comm.size();
  px->m_out_ptr.assign(3 * nlocal, nullptr);
  px->m_out_flag.assign(3 * nlocal, nullptr);
  px->m_in_offrank.assign(3 * nlocal, 0);

  // what this rank exports: its flags, and every plane chunk with an off-rank upwind neighbour
  KB200_CALL(kb200_alloc(3 * nlocal * sizeof(unsigned), (void **)&px->m_flags));
  KB200_CALL(kb200_memset(px->m_flags, 0, 3 * nlocal * sizeof(unsigned), nullptr));
  std::vector<ExportEntry> mine;
  int ok = 1;
  for (size_t s = 0; s < nlocal; ++s) {
    const long *up = f_up.getDataConst(SdomId((long)s));
    for (int dim = 0; dim < 3; ++dim) {
      if (up[dim] < 0 || pspace.globalSdomIdToRank(GlobalSdomId(up[dim])) == me) continue;
      px->m_in_offrank[3 * s + dim] = 1;
      ExportEntry e;
      memset(&e, 0, sizeof(e));
      e.gid = l2g[s];
      e.dim = dim;
      if (kb200_ipc_export(planes[dim]->devicePtr(SdomId((long)s)), e.handle) != 0) ok = 0;
      mine.push_back(e);
    }
  }
  unsigned char flag_handle[64];
  memset(flag_handle, 0, sizeof(flag_handle));
  if (kb200_ipc_export(px->m_flags, flag_handle) != 0) ok = 0;
  KB200_CALL(kb200_stream_sync(nullptr));

  // gather: [ok, count] of every rank, then the flag handles, then the (padded) tables
  std::vector<int> head(2 * (size_t)nranks);
  const int myhead[2] = {ok, (int)mine.size()};
  KB200_CALL(kb200_comm_allgather(myhead, sizeof(myhead), head.data()));
  int maxn = 0;
  bool all_ok = true;
  for (int r = 0; r < nranks; ++r) { all_ok = all_ok && head[2 * r] != 0; maxn = std::max(maxn, head[2 * r + 1]); }
  std::vector<unsigned char> flag_handles(64 * (size_t)nranks);
  KB200_CALL(kb200_comm_allgather(flag_handle, 64, flag_handles.data()));
  mine.resize((size_t)std::max(maxn, 1));
  std::vector<ExportEntry> all((size_t)std::max(maxn, 1) * nranks);
  KB200_CALL(kb200_comm_allgather(mine.data(), mine.size() * sizeof(ExportEntry), all.data()));

  // map what this rank writes: the downwind chunk and its flag, for every off-rank downwind face
  std::vector<unsigned *> peer_flags((size_t)nranks, nullptr);
  int mapped = all_ok ? 1 : 0;
  for (size_t s = 0; s < nlocal && mapped; ++s) {
    const long *down = f_down.getDataConst(SdomId((long)s));
    for (int dim = 0; dim < 3 && mapped; ++dim) {
      if (down[dim] < 0) continue;
      const int r = pspace.globalSdomIdToRank(GlobalSdomId(down[dim]));
      if (r == me) continue;
      if (!peer_flags[r]) {
        void *p = nullptr;
        if (kb200_ipc_open(&flag_handles[64 * (size_t)r], &p) != 0) { mapped = 0; break; }
        peer_flags[r] = (unsigned *)p;
        px->m_opened.push_back(p);
      }
      const ExportEntry *found = nullptr;
      for (int k = 0; k < head[2 * r + 1]; ++k) {
        const ExportEntry &e = all[(size_t)r * mine.size() + k];
        if (e.gid == down[dim] && e.dim == dim) { found = &e; break; }
      }
      void *p = nullptr;
      if (!found || kb200_ipc_open(found->handle, &p) != 0) { mapped = 0; break; }
      px->m_opened.push_back(p);
      px->m_out_ptr[3 * s + dim] = (double *)p;
      const long peer_local = *pspace.globalSdomIdToSdomId(GlobalSdomId(down[dim]));
      px->m_out_flag[3 * s + dim] = peer_flags[r] + 3 * peer_local + dim;
    }
  }
  // every rank must take the same path
  std::vector<int> votes((size_t)nranks);
  KB200_CALL(kb200_comm_allgather(&mapped, sizeof(int), votes.data()));
  bool all_mapped = true;
  for (int v : votes) all_mapped = all_mapped && v != 0;
  px->m_usable = all_mapped;
  if (!all_mapped) {
    std::fill(px->m_out_ptr.begin(), px->m_out_ptr.end(), nullptr);
    if (me == 0) fprintf(stderr, "kripke_b200: peer mapping unavailable (%s), faces go through NCCL send/recv\n", kb200_last_error());
  }
  return px;
}

PeerExchange::~PeerExchange() {
  kb200_device_sync();
  for (void *p : m_opened) kb200_ipc_close(p);
  if (m_flags) kb200_free(m_flags);
}

void PeerExchange::signal(std::vector<SdomId> const &produced) {
  std::vector<unsigned *> flags;
  for (SdomId s : produced)
    for (int dim = 0; dim < 3; ++dim)
      if (m_out_flag[3 * (size_t)*s + dim]) flags.push_back(m_out_flag[3 * (size_t)*s + dim]);
  if (!flags.empty()) KB200_CALL(kb200_p2p_signal(flags.data(), (int)flags.size(), m_epoch, nullptr));
}

void PeerExchange::wait(std::vector<SdomId> const &needed) {
  std::vector<const unsigned *> flags;
  for (SdomId s : needed)
    for (int dim = 0; dim < 3; ++dim)
      if (m_in_offrank[3 * (size_t)*s + dim]) flags.push_back(m_flags + 3 * (size_t)*s + dim);
  if (!flags.empty()) KB200_CALL(kb200_p2p_wait(flags.data(), (int)flags.size(), m_epoch, nullptr));
}

// ---- ParallelComm base (src/Kripke/ParallelComm.cpp) -------------------------------------------------
ParallelComm::ParallelComm(DataStore &data_store) : m_data_store(&data_store) {
  m_plane_data[0] = &m_data_store->getVariable<Field_IPlane>("i_plane");
  m_plane_data[1] = &m_data_store->getVariable<Field_JPlane>("j_plane");
  m_plane_data[2] = &m_data_store->getVariable<Field_KPlane>("k_plane");
}

int ParallelComm::findSubdomain(SdomId sdom_id) {
  for (size_t i = 0; i < queue_sdom_ids.size(); ++i)
    if (queue_sdom_ids[i] == *sdom_id) return (int)i;
  KRIPKE_ABORT("Cannot find subdomain id %ld in work queue\n", (long)*sdom_id);
  return -1;
}

void ParallelComm::dequeueSubdomain(SdomId sdom_id) {
  int index = findSubdomain(sdom_id);
  queue_sdom_ids.erase(queue_sdom_ids.begin() + index);
  queue_depends.erase(queue_depends.begin() + index);
}

// counts the upwind dependencies of a subdomain and queues it (ParallelComm.cpp:61-120).  Off-rank
// receives are not posted here: NCCL has no unmatched-receive queue, they are issued stage by stage.
void ParallelComm::postRecvs(DataStore &data_store, SdomId sdom_id) {
  const long *upwind = data_store.getVariable<Field_Adjacency>("upwind").getDataConst(sdom_id);
  int num_depends = 0;
  for (int dim = 0; dim < 3; ++dim)
    if (upwind[dim] >= 0) num_depends++;
  queue_sdom_ids.push_back((int)*sdom_id);
  queue_depends.push_back(num_depends);
}

// on-rank downwind bookkeeping + copy (ParallelComm.cpp:122-185)
void ParallelComm::postSends(DataStore &data_store, SdomId sdom_id, FieldStorage<double> *src_fields[3]) {
  auto &pspace = data_store.getVariable<PartitionSpace>("pspace");
  Comm comm;
  const long *downwind = data_store.getVariable<Field_Adjacency>("downwind").getDataConst(sdom_id);
  for (int dim = 0; dim < 3; ++dim) {
    if (downwind[dim] < 0) continue;
    GlobalSdomId g(downwind[dim]);
    if (pspace.globalSdomIdToRank(g) != (int)comm.rank()) continue;  // off-rank: handled by the staged exchange
    SdomId sdom_down = pspace.globalSdomIdToSdomId(g);
    for (size_t i = 0; i < queue_sdom_ids.size(); ++i)
      if (queue_sdom_ids[i] == *sdom_down) { queue_depends[i]--; break; }
    if (!m_delivered) {
      auto *src = src_fields[dim];
      if (src->isZeroPending(sdom_id)) m_plane_data[dim]->setZeroPending(sdom_down);
      else KB200_CALL(kb200_copy(m_plane_data[dim]->devicePtrOverwrite(sdom_down), src->devicePtrConst(sdom_id),
                                 m_plane_data[dim]->size(sdom_down) * sizeof(double), nullptr));
    }
  }
}

bool ParallelComm::workRemaining(void) { return queue_sdom_ids.size() > 0; }
void ParallelComm::waitAllSends(void) {}  // stream-ordered: nothing outstanding once the stage's group has been issued
void ParallelComm::testRecieves(void) {}
std::vector<SdomId> ParallelComm::getReadyList(void) {
  std::vector<SdomId> ready;
  for (size_t i = 0; i < queue_depends.size(); ++i)
    if (queue_depends[i] == 0) ready.push_back(SdomId(queue_sdom_ids[i]));
  return ready;
}

// ---- SweepComm (src/Kripke/ParallelComm/SweepComm.cpp) ---------------------------------------------------
namespace {
struct StageState {
  int calls = 0;       // readySubdomains() calls so far == current stage + 1
  int nstages = 1;
  std::vector<int> depth;  // per local subdomain
};
std::map<ParallelComm *, StageState> g_stage;

// exchange of the faces produced in `stage` (sends) / needed in stage+1 (recvs)
void exchangeStage(DataStore &ds, FieldStorage<double> *planes[3], StageState const &st, int stage) {
  auto &pspace = ds.getVariable<PartitionSpace>("pspace");
  Comm comm;
  if (comm.size() <= 1) return;
  auto &f_up = ds.getVariable<Field_Adjacency>("upwind");
  auto &f_down = ds.getVariable<Field_Adjacency>("downwind");
  PeerExchange *px = PeerExchange::get(ds);
  if (px->usable()) {
    // the sweep kernels of `stage` have stored their off-rank faces straight into the peers' plane chunks: raise the
    // peers' flags behind them, and hold the stream until the faces stage+1 needs have been flagged by their producers
    std::vector<SdomId> produced, needed;
    for (size_t s = 0; s < st.depth.size(); ++s) {
      SdomId sdom((long)s);
      if (st.depth[s] == stage) produced.push_back(sdom);
      else if (st.depth[s] == stage + 1) {
        const long *up = f_up.getDataConst(sdom);
        for (int dim = 0; dim < 3; ++dim)
          if (up[dim] >= 0 && pspace.globalSdomIdToRank(GlobalSdomId(up[dim])) != (int)comm.rank())
            planes[dim]->devicePtrOverwrite(sdom);  // defined by the peer: neither a pending zero-fill nor stale
        needed.push_back(sdom);
      }
    }
    px->signal(produced);
    px->wait(needed);
    return;
  }
  auto &l2g = ds.getVariable<Field_SdomId2GlobalSdomId>("SdomId2GlobalSdomId");
  const long *local_to_global = l2g.getDataConst(SdomId(0));
  std::vector<Message> msgs;
  for (size_t s = 0; s < st.depth.size(); ++s) {
    SdomId sdom((long)s);
    if (st.depth[s] == stage) {
      const long *down = f_down.getDataConst(sdom);
      for (int dim = 0; dim < 3; ++dim) {
        if (down[dim] < 0) continue;
        int peer = pspace.globalSdomIdToRank(GlobalSdomId(down[dim]));
        if (peer == (int)comm.rank()) continue;
        msgs.push_back(Message{down[dim], dim, peer, true, planes[dim]->devicePtr(sdom), planes[dim]->size(sdom)});
      }
    } else if (st.depth[s] == stage + 1) {
      const long *up = f_up.getDataConst(sdom);
      for (int dim = 0; dim < 3; ++dim) {
        if (up[dim] < 0) continue;
        int peer = pspace.globalSdomIdToRank(GlobalSdomId(up[dim]));
        if (peer == (int)comm.rank()) continue;
        msgs.push_back(Message{local_to_global[s], dim, peer, false, planes[dim]->devicePtrOverwrite(sdom), planes[dim]->size(sdom)});
      }
    }
  }
  runExchange(msgs);
}
}  // namespace

SweepComm::SweepComm(DataStore &data_store) : ParallelComm(data_store) {
  auto &pspace = data_store.getVariable<PartitionSpace>("pspace");
  StageState st;
  st.nstages = numStages(pspace);
  size_t n = pspace.getNumSubdomains(SPACE_PQR);
  st.depth.resize(n);
  for (size_t s = 0; s < n; ++s) st.depth[s] = sweepDepth(data_store, pspace, SdomId((long)s));
  g_stage[this] = st;
  Comm comm;
  if (comm.size() > 1) PeerExchange::get(data_store)->beginSweep();  // collective on first use; a new epoch of flags
}
SweepComm::~SweepComm() {
  g_stage.erase(this);
  Comm comm;
  // a peer may start the next sweep (and overwrite a plane chunk on this GPU) only when every rank is done reading
  if (comm.size() > 1 && PeerExchange::get(*m_data_store)->usable()) KB200_CALL(kb200_comm_barrier(nullptr));
}

void SweepComm::addSubdomain(DataStore &data_store, SdomId sdom_id) { postRecvs(data_store, sdom_id); }

bool SweepComm::workRemaining(void) {
  if (referenceOrder()) return ParallelComm::workRemaining();
  StageState &st = g_stage[this];
  return st.calls < st.nstages || ParallelComm::workRemaining();
}

std::vector<SdomId> SweepComm::readySubdomains(void) {
  if (referenceOrder()) {
    Comm comm;
    KRIPKE_ASSERT(comm.size() == 1, "KB200_SWEEP_ORDER=reference supports a single rank only\n");
    std::vector<SdomId> ready = getReadyList();
    if (ready.size() > 1) ready.resize(1);  // SweepSolver.cpp:58-59
    return ready;
  }
  StageState &st = g_stage[this];
  if (st.calls > 0) exchangeStage(*m_data_store, m_plane_data, st, st.calls - 1);
  const int stage = st.calls++;
  std::vector<SdomId> ready;
  for (int id : queue_sdom_ids)
    if (st.depth[id] == stage) ready.push_back(SdomId(id));
  return ready;
}

void SweepComm::markComplete(SdomId sdom_id) {
  dequeueSubdomain(sdom_id);
  FieldStorage<double> *src[3] = {m_plane_data[0], m_plane_data[1], m_plane_data[2]};
  postSends(*m_data_store, sdom_id, src);
}

// ---- BlockJacobiComm (src/Kripke/ParallelComm/BlockJacobiComm.cpp) ------------------------------------------
// De-facto semantics of the reference (defects 1-3): the "old" planes that are
// exchanged are all zeros, so every subdomain is swept with zero inflow on all three faces; the
// exchange is still performed once per iteration (communication-pattern proxy).
BlockJacobiComm::BlockJacobiComm(DataStore &data_store) : ParallelComm(data_store), posted_sends(false) {
  {
    Comm comm;
    if (comm.size() > 1) PeerExchange::get(data_store)->beginSweep();  // collective on first use; a new epoch of flags
  }
  ArchLayoutV al_v = data_store.getVariable<ArchLayout>("al").al_v;
  createField<Field_IPlane>(data_store, "old_i_plane", al_v, data_store.getVariable<Set>("Set/IPlane"));
  createField<Field_JPlane>(data_store, "old_j_plane", al_v, data_store.getVariable<Set>("Set/JPlane"));
  createField<Field_KPlane>(data_store, "old_k_plane", al_v, data_store.getVariable<Set>("Set/KPlane"));
  // the reference leaves these uninitialised and happens to read zero pages; make that explicit
  Kernel::kConst(data_store.getVariable<Field_IPlane>("old_i_plane"), 0.0);
  Kernel::kConst(data_store.getVariable<Field_JPlane>("old_j_plane"), 0.0);
  Kernel::kConst(data_store.getVariable<Field_KPlane>("old_k_plane"), 0.0);
}
BlockJacobiComm::~BlockJacobiComm() {
  {
    Comm comm;  // a peer may overwrite a plane chunk on this GPU only when every rank is done reading
    if (comm.size() > 1 && PeerExchange::get(*m_data_store)->usable()) KB200_CALL(kb200_comm_barrier(nullptr));
  }
  KB200_CALL(kb200_stream_sync(nullptr));
  m_data_store->deleteVariable("old_i_plane");
  m_data_store->deleteVariable("old_j_plane");
  m_data_store->deleteVariable("old_k_plane");
}
void BlockJacobiComm::addSubdomain(DataStore &data_store, SdomId sdom_id) { postRecvs(data_store, sdom_id); }

bool BlockJacobiComm::workRemaining(void) {
  if (!posted_sends) {
    FieldStorage<double> *old_planes[3] = {&m_data_store->getVariable<Field_IPlane>("old_i_plane"),
                                           &m_data_store->getVariable<Field_JPlane>("old_j_plane"),
                                           &m_data_store->getVariable<Field_KPlane>("old_k_plane")};
    // on-rank neighbours: dependency bookkeeping + (zero) copy, exactly one pass over the queue
    std::vector<int> ids = queue_sdom_ids;
    for (int id : ids) postSends(*m_data_store, SdomId(id), old_planes);
    // off-rank neighbours: one exchange of the old planes per iteration -- peer-memory copies straight into the
    // neighbours' plane chunks with device-side flags, or (fallback) one grouped NCCL exchange
    auto &pspace = m_data_store->getVariable<PartitionSpace>("pspace");
    Comm comm;
    if (comm.size() > 1) {
      auto &f_up = m_data_store->getVariable<Field_Adjacency>("upwind");
      auto &f_down = m_data_store->getVariable<Field_Adjacency>("downwind");
      const long *l2g = m_data_store->getVariable<Field_SdomId2GlobalSdomId>("SdomId2GlobalSdomId").getDataConst(SdomId(0));
      PeerExchange *px = PeerExchange::get(*m_data_store);
      std::vector<Message> msgs;
      std::vector<SdomId> all;
      for (int id : ids) {
        SdomId sdom(id);
        all.push_back(sdom);
        const long *down = f_down.getDataConst(sdom), *up = f_up.getDataConst(sdom);
        for (int dim = 0; dim < 3; ++dim) {
          if (down[dim] >= 0) {
            int peer = pspace.globalSdomIdToRank(GlobalSdomId(down[dim]));
            if (peer != (int)comm.rank()) {
              if (px->usable())
                KB200_CALL(kb200_copy(px->outPlane(sdom, dim), old_planes[dim]->devicePtr(sdom), old_planes[dim]->size(sdom) * sizeof(double), nullptr));
              else
                msgs.push_back(Message{down[dim], dim, peer, true, old_planes[dim]->devicePtr(sdom), old_planes[dim]->size(sdom)});
            }
          }
          if (up[dim] >= 0) {
            int peer = pspace.globalSdomIdToRank(GlobalSdomId(up[dim]));
            if (peer != (int)comm.rank()) {
              if (px->usable()) m_plane_data[dim]->devicePtrOverwrite(sdom);  // defined by the peer's copy
              else msgs.push_back(Message{l2g[id], dim, peer, false, m_plane_data[dim]->devicePtrOverwrite(sdom), m_plane_data[dim]->size(sdom)});
              for (size_t i = 0; i < queue_sdom_ids.size(); ++i)
                if (queue_sdom_ids[i] == id) { queue_depends[i]--; break; }
            }
          }
        }
      }
      if (px->usable()) {
        px->signal(all);
        px->wait(all);
      } else {
        runExchange(msgs);
      }
    }
    posted_sends = true;
  }
  return ParallelComm::workRemaining();
}
std::vector<SdomId> BlockJacobiComm::readySubdomains(void) { return getReadyList(); }
void BlockJacobiComm::markComplete(SdomId sdom_id) { dequeueSubdomain(sdom_id); }

// ---- SweepSolver (src/Kripke/SweepSolver.cpp:23-87) -----------------------------------------------------------
void Kripke::SweepSolver(DataStore &data_store, std::vector<SdomId> subdomain_list, bool block_jacobi) {
  KRIPKE_TIMER(data_store, SweepSolver);
  auto &i_plane = data_store.getVariable<Field_IPlane>("i_plane");
  auto &j_plane = data_store.getVariable<Field_JPlane>("j_plane");
  auto &k_plane = data_store.getVariable<Field_KPlane>("k_plane");
  Kernel::kConst(i_plane, 0.0);  // recorded as pending zero-fills, consumed by the sweep as "vacuum inflow"
  Kernel::kConst(j_plane, 0.0);
  Kernel::kConst(k_plane, 0.0);

  ParallelComm *comm = block_jacobi ? (ParallelComm *)new BlockJacobiComm(data_store) : (ParallelComm *)new SweepComm(data_store);
  for (SdomId sdom_id : subdomain_list) comm->addSubdomain(data_store, sdom_id);
  auto &field_upwind = data_store.getVariable<Field_Adjacency>("upwind");
  const bool deliver = !block_jacobi;  // the kernel writes outgoing faces straight into on-rank downwind chunks
  comm->setDelivered(deliver);
  Kernel::populationFusionBegin(data_store);

  while (comm->workRemaining()) {
    std::vector<SdomId> ready = comm->readySubdomains();
    if (ready.empty()) continue;
    for (SdomId sdom_id : ready) {
      const long *upwind = field_upwind.getDataConst(sdom_id);  // vacuum boundary (SweepSolver.cpp:64-72)
      if (upwind[0] == -1) Kernel::kConst(i_plane, sdom_id, 0.0);
      if (upwind[1] == -1) Kernel::kConst(j_plane, sdom_id, 0.0);
      if (upwind[2] == -1) Kernel::kConst(k_plane, sdom_id, 0.0);
    }
    Kernel::sweepSubdomains(data_store, ready, deliver, true);
    for (SdomId sdom_id : ready) comm->markComplete(sdom_id);
  }
  Kernel::populationFusionEnd(data_store, subdomain_list);
  delete comm;
}
