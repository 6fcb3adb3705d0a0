// DataStore, Comm, PartitionSpace, DomainVar, Sets and Field storage of the host layer.
// Reference behaviour restated (not copied) from src/Kripke/Core/*.{h,cpp}; see Kripke.h for the
// per-class citations.
#include "Kripke.h"

#include <algorithm>
#include <cstring>
#include <ctype.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

using namespace Kripke;
using namespace Kripke::Core;

// ------------------------------------------------------------------------------------------------
// ArchLayout (src/Kripke/ArchLayout.h:52-126)
// ------------------------------------------------------------------------------------------------
std::string Kripke::archToString(ArchV av) { return av == ArchV_B200 ? "B200" : "unknown"; }
ArchV Kripke::stringToArch(std::string const &str) {
  if (!strcasecmp(str.c_str(), "B200") || !strcasecmp(str.c_str(), "CUDA")) return ArchV_B200;
  return ArchV_Unknown;
}
std::string Kripke::layoutToString(LayoutV lv) {
  static const char *names[] = {"DGZ", "DZG", "GDZ", "GZD", "ZDG", "ZGD"};
  return (lv >= 0 && lv < LayoutV_num_values) ? names[lv] : "unknown";
}
LayoutV Kripke::stringToLayout(std::string const &str) {
  for (int lv = 0; lv < (int)LayoutV_num_values; ++lv)
    if (!strcasecmp(layoutToString((LayoutV)lv).c_str(), str.c_str())) return (LayoutV)lv;
  return LayoutV_Unknown;
}

// position of an index kind in DefaultOrder<Layout> (src/Kripke/VarTypes.h:73-101)
int Kripke::layoutRank(LayoutV layout, int kind) {
  // blocks: 'D' = Direction,Legendre,Moment ; 'G' = GlobalGroup,Group ; 'Z' = Zone,ZoneK,ZoneJ,ZoneI(,MixElem)
  static const char *block_order[6] = {"DGZ", "DZG", "GDZ", "GZD", "ZDG", "ZGD"};
  if (kind == K_LONG || kind == K_SDOM || kind == K_GSDOM) return 0;
  if (kind == K_DIM) return 1;
  if (kind == K_MAT) return 2;
  char blk;
  int within;
  switch (kind) {
    case K_DIR: blk = 'D'; within = 0; break;
    case K_LEG: blk = 'D'; within = 1; break;
    case K_MOM: blk = 'D'; within = 2; break;
    case K_GG: blk = 'G'; within = 0; break;
    case K_GRP: blk = 'G'; within = 1; break;
    case K_ZONE: blk = 'Z'; within = 0; break;
    case K_ZK: blk = 'Z'; within = 1; break;
    case K_ZJ: blk = 'Z'; within = 2; break;
    case K_ZI: blk = 'Z'; within = 3; break;
    default: blk = 'Z'; within = 4; break;  // K_MIX
  }
  // MixElem sits at the very end when the zone block is not last (DGZ, DZG, GDZ: "... MixElem" last;
  // GZD, ZDG, ZGD: right after ZoneI) -- in both cases directly after the Z block or after everything,
  // which never changes the relative order of the index types any Field actually combines.
  int pos = (int)(strchr(block_order[layout], blk) - block_order[layout]);
  if (kind == K_MIX && (layout == LayoutV_DGZ || layout == LayoutV_DZG || layout == LayoutV_GDZ)) return 3 + 3 * 8 + 7;
  return 3 + pos * 8 + within;
}

// ------------------------------------------------------------------------------------------------
// BaseVar / DataStore (src/Kripke/Core/BaseVar.cpp, DataStore.cpp)
// ------------------------------------------------------------------------------------------------
std::string BaseVar::getName() const { return m_parent ? m_parent->getVariableName(*this) : "---"; }

DataStore::~DataStore() {
  // delete in reverse-dependency-safe order: Fields reference Sets by pointer, so Fields go first
  std::vector<std::string> sets, others;
  for (auto &kv : m_vars) (dynamic_cast<Set *>(kv.second) ? sets : others).push_back(kv.first);
  for (auto &n : others) deleteVariable(n);
  for (auto &n : sets) deleteVariable(n);
}
void DataStore::addVariable(std::string const &name, BaseVar *var) {
  if (m_vars.find(name) != m_vars.end()) throw std::domain_error("Variable '" + name + "' already exists");
  m_vars[name] = var;
  var->setParent(this);
}
void DataStore::deleteVariable(std::string const &name) {
  auto it = m_vars.find(name);
  if (it == m_vars.end()) throw std::domain_error("Variable '" + name + "' does not exist");
  delete it->second;
  m_vars.erase(it);
}
std::string DataStore::getVariableName(BaseVar const &var) const {
  for (auto &kv : m_vars)
    if (kv.second == &var) return kv.first;
  return "===";
}
std::vector<std::string> DataStore::getVariableList() const {
  std::vector<std::string> l;
  for (auto &kv : m_vars) l.push_back(kv.first);
  return l;
}

// ------------------------------------------------------------------------------------------------
// Comm (src/Kripke/Core/Comm.h).  The decomposition is regular (every rank holds the same number
// of subdomains, groups, directions and zones), so integer reductions/scans over the P,Q,R
// sub-communicators are evaluated in closed form: sum = value * size, scan = value * (rank+1).
// The only value that genuinely differs between ranks is the number of mixed elements, whose
// global size/offset no hot-path code reads.  Floating-point reductions over all ranks
// (material volumes, Kernel::population) go through NCCL.
// ------------------------------------------------------------------------------------------------
static int g_world_rank = 0, g_world_size = 1;
static bool g_world_set = false;

Comm::Comm() : m_rank(g_world_rank), m_size(g_world_size), m_world(true) {}
void Comm::setWorld(int rank, int size) {
  g_world_rank = rank;
  g_world_size = size;
  g_world_set = true;
}
// MPI_Init's role: learn rank/size from the launcher (torchrun-style RANK / WORLD_SIZE / LOCAL_RANK),
// bind the GPU, and wire up NCCL.  The 128-byte NCCL id travels through a file next to
// MASTER_PORT (single node), unless an embedding process (bench.py) already called
// kb200_comm_init itself.
void Comm::init(int *, char ***) {
  const char *r = getenv("RANK"), *s = getenv("WORLD_SIZE"), *lr = getenv("LOCAL_RANK");
  int rank = r ? atoi(r) : 0, size = s ? atoi(s) : 1, local = lr ? atoi(lr) : 0;
  if (!g_world_set) setWorld(rank, size);
  KB200_CALL(kb200_init(local));
  int cr = 0, cn = 1;
  kb200_comm_rank(&cr, &cn);
  if (g_world_size > 1 && cn != g_world_size) {
    // The file lives in a directory only this user can enter and is named after this launch (the launcher's run id and
    // pid: every rank of one torchrun has the same parent), so a leftover of a crashed run is never read; it is created
    // exclusively (no symlink following) under a temporary name and renamed when complete.
    const char *port = getenv("MASTER_PORT"), *run = getenv("TORCHELASTIC_RUN_ID");
    char nonce[96];
    snprintf(nonce, sizeof(nonce), "%s_%ld", (run && *run) ? run : "none", (long)getppid());
    for (char *c = nonce; *c; ++c)
      if (!isalnum((unsigned char)*c) && *c != '_' && *c != '-') *c = '_';
    std::string dir = std::string("/tmp/kripke_b200_") + std::to_string((long)getuid());
    mkdir(dir.c_str(), 0700);
    struct stat sb;
    KRIPKE_ASSERT(lstat(dir.c_str(), &sb) == 0 && S_ISDIR(sb.st_mode) && sb.st_uid == getuid() && (sb.st_mode & 077) == 0,
                  "%s is not a private directory of this user\n", dir.c_str());
    std::string path = dir + "/ncclid_" + (port ? port : "0") + "_" + nonce;
    unsigned char id[128];
    if (g_world_rank == 0) {
      KB200_CALL(kb200_comm_unique_id(id));
      std::string tmp = path + ".tmp";
      unlink(tmp.c_str());
      unlink(path.c_str());
      int fd = open(tmp.c_str(), O_WRONLY | O_CREAT | O_EXCL | O_NOFOLLOW, 0600);
      KRIPKE_ASSERT(fd >= 0, "cannot create %s\n", tmp.c_str());
      ssize_t w = write(fd, id, sizeof(id));
      close(fd);
      KRIPKE_ASSERT(w == (ssize_t)sizeof(id), "short write of %s\n", tmp.c_str());
      rename(tmp.c_str(), path.c_str());
    } else {
      int fd = -1;
      for (int tries = 0; tries < 6000 && fd < 0; ++tries) {
        fd = open(path.c_str(), O_RDONLY | O_NOFOLLOW);
        if (fd < 0) { struct timespec ts = {0, 10000000}; nanosleep(&ts, nullptr); }
      }
      KRIPKE_ASSERT(fd >= 0, "timed out waiting for %s\n", path.c_str());
      ssize_t got = read(fd, id, sizeof(id));
      close(fd);
      KRIPKE_ASSERT(got == (ssize_t)sizeof(id), "short read of %s\n", path.c_str());
    }
    KB200_CALL(kb200_comm_init(g_world_rank, g_world_size, id));
    if (g_world_rank == 0) {  // everyone has joined once init returns
      remove(path.c_str());
    }
  }
}
void Comm::finalize() {
  kb200_comm_destroy();
  kb200_finalize();
}
Comm Comm::split(int, int key) const {
  // key = rank inside the new group; PartitionSpace builds its sub-communicators directly with
  // the (rank, size) constructor because it knows the processor grid
  return Comm(key, m_size, false);
}
long Comm::allReduceSumLong(long value) const { return value * m_size; }
void Comm::allReduceSumLong(long *value, size_t len) const {
  for (size_t i = 0; i < len; ++i) value[i] *= m_size;
}
void Comm::allReduceSumInt(int *value, size_t len) const {
  for (size_t i = 0; i < len; ++i) value[i] *= m_size;
}
long Comm::scanSumLong(long value) const { return value * (m_rank + 1); }
void Comm::allReduceSumDouble(double *value, size_t len) const {
  if (m_size <= 1 || len == 0) return;
  int r = 0, n = 1;
  kb200_comm_rank(&r, &n);
  if (n <= 1) return;  // no device communicator (dry planning run): keep the local value
  double *d = nullptr;
  KB200_CALL(kb200_alloc(len * sizeof(double), (void **)&d));
  KB200_CALL(kb200_upload(d, value, len * sizeof(double), nullptr));
  KB200_CALL(kb200_comm_allreduce_sum_f64(d, len, nullptr));
  KB200_CALL(kb200_download(value, d, len * sizeof(double), nullptr));
  KB200_CALL(kb200_stream_sync(nullptr));
  KB200_CALL(kb200_free(d));
}
double Comm::allReduceSumDouble(double value) const {
  allReduceSumDouble(&value, 1);
  return value;
}

// ------------------------------------------------------------------------------------------------
// PartitionSpace (src/Kripke/Core/PartitionSpace.cpp)
// ------------------------------------------------------------------------------------------------
static const int kSpaceMask[NUM_SPACES][5] = {
    {1, 0, 0, 0, 0},  // P
    {0, 1, 0, 0, 0},  // Q
    {0, 0, 1, 0, 0},  // RX
    {0, 0, 0, 1, 0},  // RY
    {0, 0, 0, 0, 1},  // RZ
    {0, 0, 1, 1, 1},  // R
    {1, 0, 1, 1, 1},  // PR
    {1, 1, 1, 1, 1},  // PQR
    {0, 0, 0, 0, 0},  // NULL
};

static long project(const std::array<long, 5> &dims, const long idx[5], const int mask[5]) {
  long v = 0;
  for (int d = 0; d < 5; ++d)
    if (mask[d]) v = v * dims[d] + idx[d];
  return v;
}
static void unproject(const std::array<long, 5> &dims, long v, long idx[5], const int mask[5]) {
  for (int d = 4; d >= 0; --d) {
    if (mask[d]) { idx[d] = v % dims[d]; v /= dims[d]; }
    else idx[d] = 0;
  }
}

PartitionSpace::PartitionSpace(Comm &base_comm, size_t P, size_t Q, size_t Rx, size_t Ry, size_t Rz)
    : m_comm_all(base_comm) {
  m_local_num_sdom.fill(0);
  m_global_num_sdom.fill(0);
  m_global_sdom_lower.fill(0);
  m_procs = {{(long)P, (long)Q, (long)Rx, (long)Ry, (long)Rz}};
  m_local = {{0, 0, 0, 0, 0}};
  size_t num_ranks = P * Q * Rx * Ry * Rz;
  KRIPKE_ASSERT(num_ranks == base_comm.size(), "Number of MPI ranks must match decomposition, expected %lu ranks\n",
                (unsigned long)num_ranks);
  // rank -> (p,q,x,y,z), row-major over (P,Q,Rx,Ry,Rz): z is the fastest rank coordinate (:25,41-43)
  long idx[5];
  unproject(m_procs, (long)base_comm.rank(), idx, kSpaceMask[SPACE_PQR]);
  for (int d = 0; d < 5; ++d) m_rank_coord[d] = idx[d];
  for (int space = 0; space < NUM_SPACES; ++space) {
    // communicator of a space = ranks that differ only in that space's coordinates
    long size = 1;
    for (int d = 0; d < 5; ++d)
      if (kSpaceMask[space][d]) size *= m_procs[d];
    long rank_in = project(m_procs, idx, kSpaceMask[space]);
    if (space == SPACE_NULL) {  // left default-constructed (= world) by the reference
      size = (long)base_comm.size();
      rank_in = (long)base_comm.rank();
    }
    m_comm_space[space] = Comm((int)rank_in, (int)size, space == SPACE_PQR || space == SPACE_R || space == SPACE_NULL);
  }
}

void PartitionSpace::setup_createSubdomains(size_t SP, size_t SQ, size_t Sx, size_t Sy, size_t Sz) {
  m_local = {{(long)SP, (long)SQ, (long)Sx, (long)Sy, (long)Sz}};
  for (int space = 0; space < NUM_SPACES; ++space) {
    long n = 1;
    for (int d = 0; d < 5; ++d)
      if (kSpaceMask[space][d]) n *= m_local[d];
    m_local_num_sdom[space] = n;
    Comm const &comm = m_comm_space[space];
    m_global_num_sdom[space] = comm.allReduceSumLong(n);
    m_global_sdom_lower[space] = comm.scanSumLong(n) - n;
  }
}

PartitionSpace::SdomCoord PartitionSpace::sdomIdToCoord(SdomId sdom_id) const {
  long idx[5];
  unproject(m_local, *sdom_id, idx, kSpaceMask[SPACE_PQR]);
  return SdomCoord{{idx[0], idx[1], idx[2], idx[3], idx[4]}};
}
SdomId PartitionSpace::coordToSdomId(SdomCoord c) const {
  long idx[5] = {c[0], c[1], c[2], c[3], c[4]};
  return SdomId(project(m_local, idx, kSpaceMask[SPACE_PQR]));
}
PartitionSpace::SdomCoord PartitionSpace::coordToGlobalCoord(SdomCoord l) const {
  return SdomCoord{{(ptrdiff_t)(l[0] + m_global_sdom_lower[SPACE_P]), (ptrdiff_t)(l[1] + m_global_sdom_lower[SPACE_Q]),
                    (ptrdiff_t)(l[2] + m_global_sdom_lower[SPACE_RX]), (ptrdiff_t)(l[3] + m_global_sdom_lower[SPACE_RY]),
                    (ptrdiff_t)(l[4] + m_global_sdom_lower[SPACE_RZ])}};
}
static std::array<long, 5> globalDims(std::array<long, NUM_SPACES> const &g) {
  return {{g[SPACE_P], g[SPACE_Q], g[SPACE_RX], g[SPACE_RY], g[SPACE_RZ]}};
}
PartitionSpace::SdomCoord PartitionSpace::globalSdomIdToCoord(GlobalSdomId g) const {
  long idx[5];
  unproject(globalDims(m_global_num_sdom), *g, idx, kSpaceMask[SPACE_PQR]);
  return SdomCoord{{idx[0], idx[1], idx[2], idx[3], idx[4]}};
}
GlobalSdomId PartitionSpace::coordToGlobalSdomId(SdomCoord c) const {
  long idx[5] = {c[0], c[1], c[2], c[3], c[4]};
  return GlobalSdomId(project(globalDims(m_global_num_sdom), idx, kSpaceMask[SPACE_PQR]));
}
int PartitionSpace::globalSdomIdToRank(GlobalSdomId g) const {
  SdomCoord c = globalSdomIdToCoord(g);
  long pr[5];
  for (int d = 0; d < 5; ++d) pr[d] = c[d] / m_local[d];
  return (int)project(m_procs, pr, kSpaceMask[SPACE_PQR]);
}
SdomId PartitionSpace::globalSdomIdToSdomId(GlobalSdomId g) const {
  SdomCoord c = globalSdomIdToCoord(g);
  long l[5];
  for (int d = 0; d < 5; ++d) l[d] = c[d] % m_local[d];
  return SdomId(project(m_local, l, kSpaceMask[SPACE_PQR]));
}
size_t PartitionSpace::subdomainToSpace(SPACE space, SdomId sdom_id) const {
  long idx[5];
  unproject(m_local, *sdom_id, idx, kSpaceMask[SPACE_PQR]);
  return (size_t)project(m_local, idx, kSpaceMask[space]);
}
SdomId PartitionSpace::spaceToSubdomain(SPACE space, size_t space_id) const {
  long idx[5];
  unproject(m_local, (long)space_id, idx, kSpaceMask[space]);
  return SdomId(project(m_local, idx, kSpaceMask[SPACE_PQR]));
}

void PartitionSpace::createSubdomainData(DataStore &data_store) const {
  auto &pspace = data_store.getVariable<PartitionSpace>("pspace");
  auto &set_sdomid = data_store.newVariable<LocalRangeSet>("Set/SdomId", *this, getNumSubdomains(SPACE_PQR));
  auto &set_global = data_store.newVariable<GlobalRangeSet>("Set/GlobalSdomIdLinear", pspace, set_sdomid);
  auto &l2g = data_store.newVariable<Field_SdomId2GlobalSdomId>("SdomId2GlobalSdomId", set_sdomid, LayoutV_DGZ);
  auto &g2l = data_store.newVariable<Field_GlobalSdomId2SdomId>("GlobalSdomId2SdomId", set_global, LayoutV_DGZ);
  auto &g2r = data_store.newVariable<Field_GlobalSdomId2Rank>("GlobalSdomId2Rank", set_global, LayoutV_DGZ);
  // every rank can compute the tables in closed form (the reference all-reduces them, :184-188)
  for (SdomId sdom_id : set_sdomid.getWorkList()) {
    long *p_l2g = l2g.getData(sdom_id);
    long *p_g2l = g2l.getData(sdom_id), *p_g2r = g2r.getData(sdom_id);
    for (SdomId local{0}; local < (long)set_sdomid.size(sdom_id); ++local)
      p_l2g[*local] = *coordToGlobalSdomId(coordToGlobalCoord(sdomIdToCoord(local)));
    for (long g = 0; g < (long)set_global.size(sdom_id); ++g) {
      p_g2l[g] = *globalSdomIdToSdomId(GlobalSdomId(g));
      p_g2r[g] = globalSdomIdToRank(GlobalSdomId(g));
    }
  }
}

void PartitionSpace::print() const {
  if (m_comm_all.rank() != 0) return;
  printf("  Decomposition Space:   Procs:      Subdomains (local/global):\n");
  printf("  ---------------------  ----------  --------------------------\n");
  printf("  (P) Energy:            %-10d  %d / %d\n", (int)m_comm_space[SPACE_P].size(), (int)m_local_num_sdom[SPACE_P], (int)m_global_num_sdom[SPACE_P]);
  printf("  (Q) Direction:         %-10d  %d / %d\n", (int)m_comm_space[SPACE_Q].size(), (int)m_local_num_sdom[SPACE_Q], (int)m_global_num_sdom[SPACE_Q]);
  printf("  (R) Space:             %-10d  %d / %d\n", (int)m_comm_space[SPACE_R].size(), (int)m_local_num_sdom[SPACE_R], (int)m_global_num_sdom[SPACE_R]);
  printf("  (Rx,Ry,Rz) R in XYZ:   %dx%dx%d       %dx%dx%d / %dx%dx%d\n", (int)m_comm_space[SPACE_RX].size(),
         (int)m_comm_space[SPACE_RY].size(), (int)m_comm_space[SPACE_RZ].size(), (int)m_local_num_sdom[SPACE_RX],
         (int)m_local_num_sdom[SPACE_RY], (int)m_local_num_sdom[SPACE_RZ], (int)m_global_num_sdom[SPACE_RX],
         (int)m_global_num_sdom[SPACE_RY], (int)m_global_num_sdom[SPACE_RZ]);
  printf("  (PQR) TOTAL:           %-10d  %d / %d\n", (int)m_comm_all.size(), (int)getNumSubdomains(),
         (int)(m_global_num_sdom[SPACE_P] * m_global_num_sdom[SPACE_Q] * m_global_num_sdom[SPACE_RX] *
               m_global_num_sdom[SPACE_RY] * m_global_num_sdom[SPACE_RZ]));
}

// ------------------------------------------------------------------------------------------------
// DomainVar / Sets (src/Kripke/Core/DomainVar.cpp, Set.cpp)
// ------------------------------------------------------------------------------------------------
void DomainVar::setup_initChunks(PartitionSpace const &pspace, SPACE space) {
  size_t num_subdomains = pspace.getNumSubdomains();
  size_t num_chunks = pspace.getNumSubdomains(space);
  m_subdomain_to_chunk.resize(num_subdomains);
  for (size_t s = 0; s < num_subdomains; ++s) m_subdomain_to_chunk[s] = pspace.subdomainToSpace(space, SdomId((long)s));
  m_chunk_to_subdomain.resize(num_chunks);
  m_work_list.resize(num_chunks);
  for (size_t c = 0; c < num_chunks; ++c) {
    SdomId sdom_id = pspace.spaceToSubdomain(space, c);
    m_chunk_to_subdomain[c] = (size_t)*sdom_id;
    m_work_list[c] = sdom_id;
  }
}
void DomainVar::setup_initChunks(DomainVar const &o) {
  m_subdomain_to_chunk = o.m_subdomain_to_chunk;
  m_chunk_to_subdomain = o.m_chunk_to_subdomain;
  m_work_list = o.m_work_list;
}

RangeSet::RangeSet(PartitionSpace const &pspace, SPACE space, std::vector<size_t> const &local_sizes) : m_space(space) {
  Comm const &comm = pspace.getComm(m_space);
  setup_initChunks(pspace, m_space);
  size_t num_chunks = m_chunk_to_subdomain.size();
  KRIPKE_ASSERT(local_sizes.size() == num_chunks, "Space %d has %lu subdomains, but provided %lu subdomains", (int)m_space,
                (unsigned long)num_chunks, (unsigned long)local_sizes.size());
  long total_local = 0;
  for (size_t s : local_sizes) total_local += (long)s;
  m_global_size = (size_t)comm.allReduceSumLong(total_local);
  m_chunk_to_size = local_sizes;
  m_chunk_to_lower.resize(num_chunks);
  if (num_chunks) m_chunk_to_lower[0] = (size_t)(comm.scanSumLong(total_local) - total_local);
  for (size_t i = 1; i < num_chunks; ++i) m_chunk_to_lower[i] = m_chunk_to_lower[i - 1] + m_chunk_to_size[i - 1];
}
LocalRangeSet::LocalRangeSet(PartitionSpace const &pspace, size_t local_size) {
  Comm const &comm = pspace.getComm(SPACE_PQR);
  setup_initChunks(pspace, SPACE_NULL);
  KRIPKE_ASSERT(m_chunk_to_subdomain.size() == 1, "Something's wrong, SPACE_NULL should have 1");
  m_global_size = (size_t)comm.allReduceSumLong((long)local_size);
  m_chunk_to_size = {local_size};
  m_chunk_to_lower = {(size_t)(comm.scanSumLong((long)local_size) - (long)local_size)};
}
GlobalRangeSet::GlobalRangeSet(PartitionSpace const &pspace, size_t global_size) { setup_setGlobalSize(pspace, global_size); }
GlobalRangeSet::GlobalRangeSet(PartitionSpace const &pspace, Set &parent_set) { setup_setGlobalSize(pspace, parent_set.globalSize()); }
void GlobalRangeSet::setup_setGlobalSize(PartitionSpace const &pspace, size_t global_size) {
  setup_initChunks(pspace, SPACE_NULL);
  m_chunk_to_size.assign(1, global_size);
  m_chunk_to_lower.assign(1, 0);
  m_global_size = global_size;
}

// ------------------------------------------------------------------------------------------------
// Field storage (src/Kripke/Core/Field.h:61-148): device chunks + lazy host mirrors
// ------------------------------------------------------------------------------------------------
FieldStorageBase::FieldStorageBase(Set const &spanned_set, size_t elem_size) : m_set(&spanned_set), m_elem_size(elem_size) {
  setup_initChunks(spanned_set);
  size_t num_chunks = m_chunk_to_subdomain.size();
  m_chunk_to_size.resize(num_chunks, 0);
  m_chunks.resize(num_chunks);
  for (size_t c = 0; c < num_chunks; ++c) m_chunk_to_size[c] = spanned_set.size(SdomId((long)m_chunk_to_subdomain[c]));
  // device memory is allocated on first device use (devPtr), host mirrors on first host use
}
FieldStorageBase::~FieldStorageBase() {
  for (auto &c : m_chunks) {
    if (c.dev) kb200_free(c.dev);
    if (c.host) free(c.host);
  }
}
void FieldStorageBase::releaseHostMirrors() {
  for (size_t i = 0; i < m_chunks.size(); ++i) {
    Chunk &c = m_chunks[i];
    if (c.host && c.dev_valid) { free(c.host); c.host = nullptr; c.host_valid = false; }
  }
}
double FieldStorageBase::uniformPositiveValue(SdomId sdom_id) {
  size_t ci = m_subdomain_to_chunk[*sdom_id];
  if (m_elem_size != sizeof(double)) return 0.0;
  if (m_chunks[ci].uniform_epoch != m_chunks[ci].write_epoch) {
    Chunk &c = m_chunks[ci];
    const size_t n = m_chunk_to_size[ci];
    double u = 0.0;
    if (c.dev && c.dev_valid && !c.host_valid && !c.zero_pending) {
      // the current contents live on the device only (written there by a kernel or an upload): ask the device
      KB200_CALL(kb200_uniform_positive_value((const double *)c.dev, n, &u, nullptr));
    } else {
      const double *v = (const double *)hostPtr(sdom_id, false);
      u = n ? v[0] : 0.0;
      for (size_t i = 1; i < n && u > 0.0; ++i)
        if (v[i] != u) u = 0.0;
    }
    c.uniform_value = (u > 0.0) ? u : 0.0;
    c.uniform_epoch = c.write_epoch;
  }
  return m_chunks[ci].uniform_value;
}
void FieldStorageBase::materializeZero(Chunk &c, size_t bytes) {
  // the pending zero-fill becomes real, on whichever side is being touched
  if (c.host) { memset(c.host, 0, bytes); c.host_valid = true; }
  if (c.dev) { KB200_CALL(kb200_memset(c.dev, 0, bytes, nullptr)); c.dev_valid = true; }
  if (!c.host && !c.dev) { c.host_valid = c.dev_valid = false; }
  c.zero_pending = false;
}
void *FieldStorageBase::hostPtr(SdomId sdom_id, bool will_write) {
  size_t ci = m_subdomain_to_chunk[*sdom_id];
  Chunk &c = m_chunks[ci];
  size_t bytes = m_chunk_to_size[ci] * m_elem_size;
  if (!c.host) {
    c.host = malloc(bytes ? bytes : 8);
    KRIPKE_ASSERT(c.host != nullptr, "out of host memory for a %lu byte mirror\n", (unsigned long)bytes);
    c.host_valid = false;
  }
  if (c.zero_pending) {
    memset(c.host, 0, bytes);
    c.host_valid = true;
    c.dev_valid = false;
    c.zero_pending = false;
  } else if (!c.host_valid) {
    if (c.dev_valid) {
      KB200_CALL(kb200_download(c.host, c.dev, bytes, nullptr));
      KB200_CALL(kb200_stream_sync(nullptr));
    }
    // else: never written anywhere -- uninitialised, exactly like the reference's new ELEMENT[] (Field.h:69)
    c.host_valid = true;
  }
  if (will_write) { c.dev_valid = false; ++c.write_epoch; }
  return c.host;
}
void *FieldStorageBase::devPtr(SdomId sdom_id, bool will_write) {
  size_t ci = m_subdomain_to_chunk[*sdom_id];
  Chunk &c = m_chunks[ci];
  size_t bytes = m_chunk_to_size[ci] * m_elem_size;
  if (!c.dev) {
    KB200_CALL(kb200_alloc(bytes, &c.dev));
    c.dev_valid = false;
  }
  if (c.zero_pending) {
    KB200_CALL(kb200_memset(c.dev, 0, bytes, nullptr));  // exact byte count: 4-byte fields may have an odd length
    c.dev_valid = true;
    c.host_valid = false;
    c.zero_pending = false;
  } else if (!c.dev_valid) {
    if (c.host_valid) {
      KB200_CALL(kb200_upload(c.dev, c.host, bytes, nullptr));
      KB200_CALL(kb200_stream_sync(nullptr));  // the mirror is pageable memory
    }
    c.dev_valid = true;
  }
  if (will_write) { c.host_valid = false; ++c.write_epoch; }
  return c.dev;
}
void *FieldStorageBase::devPtrOverwrite(SdomId sdom_id) {
  size_t ci = m_subdomain_to_chunk[*sdom_id];
  Chunk &c = m_chunks[ci];
  if (!c.dev) KB200_CALL(kb200_alloc(m_chunk_to_size[ci] * m_elem_size, &c.dev));
  c.dev_valid = true;
  c.host_valid = false;
  c.zero_pending = false;
  ++c.write_epoch;
  return c.dev;
}
void FieldStorageBase::setZeroPending(SdomId sdom_id) {
  Chunk &c = m_chunks[m_subdomain_to_chunk[*sdom_id]];
  c.zero_pending = true;
  c.host_valid = false;
  c.dev_valid = false;
  ++c.write_epoch;
}
bool FieldStorageBase::consumeZeroPending(SdomId sdom_id) {
  Chunk &c = m_chunks[m_subdomain_to_chunk[*sdom_id]];
  bool z = c.zero_pending;
  c.zero_pending = false;
  return z;
}
