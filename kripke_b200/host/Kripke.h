// Host-side mirror of the reference's C++ interface for the source-iteration hot path.
//
// Same names, argument meaning and error behaviour as LLNL/Kripke's headers (cited per item,
// paths relative to the reference tree), re-implemented without RAJA/CHAI: every Field chunk lives
// in B200 HBM (allocated through the C ABI in include/kripke_b200.h) with an optional, lazily
// created host mirror, and the Kripke::Kernel entry points forward to hand-written sm_100a kernels.
// There is no CPU execution path.
#ifndef KRIPKE_B200_HOST_KRIPKE_H
#define KRIPKE_B200_HOST_KRIPKE_H

#include <array>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <string>
#include <strings.h>
#include <utility>
#include <vector>

#include "kripke_b200.h"

#define KRIPKE_VERSION "b200-0.1 (interface of Kripke 1.2.5-dev)"

// src/Kripke.h:39-53
#define KRIPKE_ABORT(...)       \
  do {                          \
    printf(__VA_ARGS__);        \
    fflush(stdout);             \
    fprintf(stderr, __VA_ARGS__); /* embedders (ctypes) may have redirected stdout */ \
    fflush(stderr);             \
    exit(1);                    \
  } while (0)
#define KRIPKE_ASSERT(EXPR, ...)                        \
  do {                                                  \
    if (!(EXPR)) {                                      \
      KRIPKE_ABORT("Assertion Failed: " __VA_ARGS__);   \
    }                                                   \
  } while (0)
// every C-ABI call goes through this: non-zero -> KRIPKE_ABORT
#define KB200_CALL(x)                                                            \
  do {                                                                           \
    int kb_rc_ = (x);                                                            \
    if (kb_rc_) KRIPKE_ABORT("kripke_b200: %s failed (%d): %s\n", #x, kb_rc_, kb200_last_error()); \
  } while (0)

enum ParallelMethod { PMETHOD_SWEEP, PMETHOD_BJ };  // src/Kripke.h:80-83

namespace Kripke {

// Strongly typed indices (the reference uses RAJA_INDEX_VALUE, src/Kripke.h:64-72, VarTypes.h:19-30)
template <typename TAG>
struct IndexValue {
  static constexpr int kind = TAG::kind;
  long value;
  IndexValue() : value(0) {}
  explicit IndexValue(long v) : value(v) {}
  long operator*() const { return value; }
  IndexValue &operator++() { ++value; return *this; }
  IndexValue operator++(int) { IndexValue t(*this); ++value; return t; }
  bool operator<(IndexValue o) const { return value < o.value; }
  bool operator<(long o) const { return value < o; }
  bool operator==(IndexValue o) const { return value == o.value; }
  bool operator!=(IndexValue o) const { return value != o.value; }
  IndexValue operator+(IndexValue o) const { return IndexValue(value + o.value); }
};
inline long idx_value(long v) { return v; }
template <typename TAG>
inline long idx_value(IndexValue<TAG> v) { return v.value; }

// kinds follow the position logic of DefaultOrder (VarTypes.h:73-101)
enum IndexKind { K_LONG = 0, K_DIM, K_MAT, K_DIR, K_LEG, K_MOM, K_GG, K_GRP, K_ZONE, K_ZK, K_ZJ, K_ZI, K_MIX, K_SDOM, K_GSDOM };
#define KRIPKE_INDEX(NAME, KIND)        \
  struct NAME##_tag { static constexpr int kind = KIND; }; \
  using NAME = IndexValue<NAME##_tag>
KRIPKE_INDEX(SdomId, K_SDOM);
KRIPKE_INDEX(GlobalSdomId, K_GSDOM);
KRIPKE_INDEX(Dimension, K_DIM);
KRIPKE_INDEX(Direction, K_DIR);
KRIPKE_INDEX(GlobalGroup, K_GG);
KRIPKE_INDEX(Group, K_GRP);
KRIPKE_INDEX(Legendre, K_LEG);
KRIPKE_INDEX(Material, K_MAT);
KRIPKE_INDEX(MixElem, K_MIX);
KRIPKE_INDEX(Moment, K_MOM);
KRIPKE_INDEX(Zone, K_ZONE);
KRIPKE_INDEX(ZoneI, K_ZI);
KRIPKE_INDEX(ZoneJ, K_ZJ);
KRIPKE_INDEX(ZoneK, K_ZK);

// ---- ArchLayout: src/Kripke/ArchLayout.h ---------------------------------------------------------
// Only one architecture exists here: the B200.  "CUDA" is accepted as an alias so existing
// command lines keep working; Sequential/OpenMP are rejected (no CPU fallback).
enum ArchV { ArchV_Unknown = -1, ArchV_B200 = 0, ArchV_num_values };
enum LayoutV { LayoutV_Unknown = -1, LayoutV_DGZ, LayoutV_DZG, LayoutV_GDZ, LayoutV_GZD, LayoutV_ZDG, LayoutV_ZGD, LayoutV_num_values };
std::string archToString(ArchV av);
ArchV stringToArch(std::string const &str);
std::string layoutToString(LayoutV lv);
LayoutV stringToLayout(std::string const &str);
struct ArchLayoutV { ArchV arch_v; LayoutV layout_v; };
// rank of an index kind in DefaultOrder<layout> (VarTypes.h:73-101); smaller = slower
int layoutRank(LayoutV layout, int kind);

namespace Core {

class DataStore;

// ---- BaseVar: src/Kripke/Core/BaseVar.h ----------------------------------------------------------
class BaseVar {
 public:
  BaseVar() : m_parent(nullptr) {}
  virtual ~BaseVar() = default;
  void setParent(DataStore *parent) { m_parent = parent; }
  std::string getName() const;
 private:
  DataStore *m_parent;
};

// ---- DataStore: src/Kripke/Core/DataStore.h:30-84, DataStore.cpp ---------------------------------
class DataStore {
 public:
  DataStore() = default;
  ~DataStore();
  DataStore(DataStore const &) = delete;
  DataStore &operator=(DataStore const &) = delete;

  void addVariable(std::string const &name, BaseVar *var);
  template <typename T, typename... CTOR_ARGS>
  T &newVariable(std::string const &name, CTOR_ARGS &&... ctor_args) {
    T *new_var = new T(ctor_args...);
    addVariable(name, new_var);
    return *new_var;
  }
  void deleteVariable(std::string const &name);
  template <typename T>
  T &getVariable(std::string const &name) {
    auto it = m_vars.find(name);
    if (it == m_vars.end()) throw std::domain_error("Cannot find '" + name + "' in DataStore");
    T *var_ptr = dynamic_cast<T *>(it->second);
    KRIPKE_ASSERT(var_ptr != nullptr, "Error casting '%s'", name.c_str());
    return *var_ptr;
  }
  template <typename T>
  T const &getVariable(std::string const &name) const {
    return const_cast<DataStore *>(this)->template getVariable<T>(name);
  }
  std::string getVariableName(BaseVar const &var) const;
  template <typename T>
  bool isVariableType(std::string const &name) const {
    auto it = m_vars.find(name);
    if (it == m_vars.end()) return false;
    return dynamic_cast<T *>(it->second) != nullptr;
  }
  std::vector<std::string> getVariableList() const;
 private:
  std::map<std::string, BaseVar *> m_vars;
};

// ---- Comm: src/Kripke/Core/Comm.h ------------------------------------------------------------------
// One process per GPU.  rank/size come from the launcher environment (RANK / WORLD_SIZE, as set by
// torchrun) or from Comm::setWorld(); reductions over ranks go through NCCL (kb200_comm_*).
class Comm : public BaseVar {
 public:
  Comm();
  Comm(int rank, int size, bool world) : m_rank(rank), m_size(size), m_world(world) {}
  static void init(int *argc, char ***argv);
  static void finalize();
  static void setWorld(int rank, int size);
  size_t size() const { return (size_t)m_size; }
  size_t rank() const { return (size_t)m_rank; }
  Comm split(int color, int key) const;
  long allReduceSumLong(long value) const;
  void allReduceSumLong(long *value, size_t len) const;
  void allReduceSumInt(int *value, size_t len) const;
  double allReduceSumDouble(double value) const;
  void allReduceSumDouble(double *value, size_t len) const;
  long scanSumLong(long value) const;
 private:
  int m_rank, m_size;
  bool m_world;  // collectives over all ranks (true) or a degenerate 1-rank communicator
};

// ---- PartitionSpace: src/Kripke/Core/PartitionSpace.h:19-104 ---------------------------------------
enum SPACE { SPACE_P = 0, SPACE_Q, SPACE_RX, SPACE_RY, SPACE_RZ, SPACE_R, SPACE_PR, SPACE_PQR, SPACE_NULL, NUM_SPACES };

class PartitionSpace : public BaseVar {
 public:
  using SdomCoord = std::array<ptrdiff_t, 5>;
  PartitionSpace(Comm &base_comm, size_t P, size_t Q, size_t Rx, size_t Ry, size_t Rz);
  virtual ~PartitionSpace() = default;
  void setup_createSubdomains(size_t SP, size_t SQ, size_t Sx, size_t Sy, size_t Sz);
  void createSubdomainData(DataStore &data_store) const;
  size_t getNumSubdomains(SPACE space = SPACE_PQR) const { return m_local_num_sdom[space]; }
  size_t getGlobalNumSubdomains(SPACE space = SPACE_PQR) const { return m_global_num_sdom[space]; }
  SdomCoord sdomIdToCoord(SdomId sdom_id) const;
  SdomId coordToSdomId(SdomCoord coord) const;
  SdomCoord coordToGlobalCoord(SdomCoord local_coord) const;
  SdomCoord globalSdomIdToCoord(GlobalSdomId global_sdom_id) const;
  GlobalSdomId coordToGlobalSdomId(SdomCoord global_coord) const;
  // rank that owns a global subdomain / its local id there (computed, no collective needed)
  int globalSdomIdToRank(GlobalSdomId g) const;
  SdomId globalSdomIdToSdomId(GlobalSdomId g) const;
  Comm const &getComm(SPACE space) const { return m_comm_space[space]; }
  size_t subdomainToSpace(SPACE space, SdomId sdom_id) const;
  SdomId spaceToSubdomain(SPACE space, size_t sdom_space) const;
  std::array<long, 5> const &procDims() const { return m_procs; }
  std::array<long, 5> const &procCoord() const { return m_rank_coord; }
  void print() const;
 private:
  Comm m_comm_all;
  Comm m_comm_space[NUM_SPACES];
  std::array<long, NUM_SPACES> m_local_num_sdom, m_global_num_sdom, m_global_sdom_lower;
  std::array<long, 5> m_procs;       // P, Q, Rx, Ry, Rz ranks
  std::array<long, 5> m_rank_coord;  // this rank's coordinate
  std::array<long, 5> m_local;       // SP, SQ, Sx, Sy, Sz
};

// ---- DomainVar: src/Kripke/Core/DomainVar.h --------------------------------------------------------
class DomainVar : public BaseVar {
 public:
  DomainVar() = default;
  virtual ~DomainVar() = default;
  DomainVar(DomainVar const &) = delete;
  DomainVar &operator=(DomainVar const &) = delete;
  size_t getNumSubdomains() const { return m_subdomain_to_chunk.size(); }
  std::vector<SdomId> const &getWorkList() const { return m_work_list; }
  size_t chunkOf(SdomId sdom_id) const { return m_subdomain_to_chunk[*sdom_id]; }
  size_t getNumChunks() const { return m_chunk_to_subdomain.size(); }
 protected:
  void setup_initChunks(PartitionSpace const &pspace, SPACE space);
  void setup_initChunks(DomainVar const &clone_from);
  std::vector<size_t> m_subdomain_to_chunk;
  std::vector<size_t> m_chunk_to_subdomain;
  std::vector<SdomId> m_work_list;
};

// ---- Set family: src/Kripke/Core/Set.h, Set.cpp ----------------------------------------------------
class Set : public DomainVar {
 public:
  Set() : m_global_size(0) {}
  virtual ~Set() = default;
  Set(Set const &) = delete;
  size_t size(SdomId sdom_id) const { return m_chunk_to_size[m_subdomain_to_chunk[*sdom_id]]; }
  size_t lower(SdomId sdom_id) const { return m_chunk_to_lower[m_subdomain_to_chunk[*sdom_id]]; }
  size_t globalSize() const { return m_global_size; }
  virtual size_t getNumDimensions() const = 0;
  virtual size_t dimSize(SdomId sdom_id, size_t) const { return size(sdom_id); }
 protected:
  std::vector<size_t> m_chunk_to_size, m_chunk_to_lower;
  size_t m_global_size;
};

class RangeSet : public Set {
 public:
  RangeSet(PartitionSpace const &pspace, SPACE space, std::vector<size_t> const &local_sizes);
  size_t getNumDimensions() const override { return 1; }
 private:
  SPACE m_space;
};
class LocalRangeSet : public Set {
 public:
  LocalRangeSet(PartitionSpace const &pspace, size_t local_size);
  size_t getNumDimensions() const override { return 1; }
};
class GlobalRangeSet : public Set {
 public:
  GlobalRangeSet(PartitionSpace const &pspace, size_t global_size);
  GlobalRangeSet(PartitionSpace const &pspace, Set &parent_set);
  size_t getNumDimensions() const override { return 1; }
 private:
  void setup_setGlobalSize(PartitionSpace const &pspace, size_t global_size);
};

// row-major layout over N extents (what ProductSet::getLayout returns, Set.h:158-169)
template <size_t N>
struct DenseLayout {
  std::array<long, N> sizes;
  template <typename... I>
  long operator()(I... idx) const {
    long v[N] = {(long)idx_value(idx)...};
    long off = 0;
    for (size_t d = 0; d < N; ++d) off = off * sizes[d] + v[d];
    return off;
  }
};

template <size_t NUM_SETS>
class ProductSet : public Set {
 public:
  using LayoutType = DenseLayout<NUM_SETS>;
  template <typename... SPAN>
  ProductSet(PartitionSpace &pspace, SPACE space, SPAN const &... spanned_sets) {
    static_assert(sizeof...(SPAN) == NUM_SETS, "Must provide same number of sets as dimensionality of ProductSet");
    setup_initChunks(pspace, space);
    m_spanned_sets = {{(&spanned_sets)...}};
    size_t num_chunks = m_chunk_to_subdomain.size();
    m_chunk_to_size.assign(num_chunks, 1);
    m_chunk_to_lower.assign(num_chunks, 0);
    for (size_t c = 0; c < num_chunks; ++c) {
      SdomId sdom_id(m_chunk_to_subdomain[c]);
      for (size_t s = 0; s < NUM_SETS; ++s) m_chunk_to_size[c] *= m_spanned_sets[s]->size(sdom_id);
    }
    m_global_size = 1;
    for (size_t s = 0; s < NUM_SETS; ++s) m_global_size *= m_spanned_sets[s]->globalSize();
  }
  size_t getNumDimensions() const override { return NUM_SETS; }
  size_t dimSize(SdomId sdom_id, size_t dim) const override { return m_spanned_sets[dim]->size(sdom_id); }
  LayoutType getLayout(SdomId sdom_id) const {
    LayoutType l;
    for (size_t d = 0; d < NUM_SETS; ++d) l.sizes[d] = (long)dimSize(sdom_id, d);
    return l;
  }
 private:
  std::array<Set const *, NUM_SETS> m_spanned_sets;
};

// ---- Field: src/Kripke/Core/Field.h -----------------------------------------------------------------
// Chunk storage in device memory + lazily created host mirror, CHAI-like coherence:
//   getData()/getView()      -> host pointer, data migrated to the host, device copy invalidated
//   getDataConst()           -> host pointer, both copies stay valid
//   devicePtr()/devicePtrConst() -> device pointer, host data (if newer) uploaded first
// kConst(field, 0.0) only records a pending zero-fill; kernels that overwrite or accumulate into
// the chunk consume the flag (fold the memset), any other access materialises it.
class FieldStorageBase : public DomainVar {
 public:
  explicit FieldStorageBase(Set const &spanned_set, size_t elem_size);
  virtual ~FieldStorageBase();
  FieldStorageBase(FieldStorageBase const &) = delete;
  size_t size(SdomId sdom_id) const { return m_chunk_to_size[m_subdomain_to_chunk[*sdom_id]]; }
  Set const &getSet() const { return *m_set; }
  size_t elemSize() const { return m_elem_size; }
  void *hostPtr(SdomId sdom_id, bool will_write);
  void *devPtr(SdomId sdom_id, bool will_write);
  // marks the chunk as "all zeros" without touching memory
  void setZeroPending(SdomId sdom_id);
  // returns true (and clears the flag) if the chunk is a pending zero-fill; the caller promises to
  // define every element of the chunk on the device
  bool consumeZeroPending(SdomId sdom_id);
  bool isZeroPending(SdomId sdom_id) const { return m_chunks[m_subdomain_to_chunk[*sdom_id]].zero_pending; }
  // bumped by every access that may change the chunk's contents (used to validate results cached from it)
  unsigned long writeEpoch(SdomId sdom_id) const { return m_chunks[m_subdomain_to_chunk[*sdom_id]].write_epoch; }
  // device pointer for a kernel that defines the whole chunk (no upload, no memset)
  void *devPtrOverwrite(SdomId sdom_id);
  void releaseHostMirrors();
  // fields of doubles: the value if every element of the chunk is the same positive number, else 0; looked at on the
  // host once per content (write epoch) of the chunk
  double uniformPositiveValue(SdomId sdom_id);
 protected:
  struct Chunk {
    void *dev = nullptr;
    void *host = nullptr;
    bool host_valid = false, dev_valid = false, zero_pending = false;
    unsigned long write_epoch = 0;
    unsigned long uniform_epoch = ~0ul;  // write epoch uniform_value was computed for
    double uniform_value = 0.0;
  };
  void materializeZero(Chunk &c, size_t bytes);
  Set const *m_set;
  size_t m_elem_size;
  std::vector<size_t> m_chunk_to_size;
  std::vector<Chunk> m_chunks;
};

template <typename ELEMENT>
class FieldStorage : public FieldStorageBase {
 public:
  using ElementType = ELEMENT;
  explicit FieldStorage(Set const &spanned_set) : FieldStorageBase(spanned_set, sizeof(ELEMENT)) {}
  ElementType *getData(SdomId sdom_id) { return (ElementType *)hostPtr(sdom_id, true); }
  ElementType const *getDataConst(SdomId sdom_id) { return (ElementType const *)hostPtr(sdom_id, false); }
  ElementType *devicePtr(SdomId sdom_id) { return (ElementType *)devPtr(sdom_id, true); }
  ElementType const *devicePtrConst(SdomId sdom_id) { return (ElementType const *)devPtr(sdom_id, false); }
  ElementType *devicePtrOverwrite(SdomId sdom_id) { return (ElementType *)devPtrOverwrite(sdom_id); }
};

// strided host view
template <typename ELEMENT, size_t N>
struct View {
  ELEMENT *ptr;
  std::array<long, N> strides;
  template <typename... I>
  ELEMENT &operator()(I... idx) const {
    static_assert(sizeof...(I) == N, "wrong number of indices");
    long v[N] = {(long)idx_value(idx)...};
    long off = 0;
    for (size_t d = 0; d < N; ++d) off += v[d] * strides[d];
    return ptr[off];
  }
};

template <typename ELEMENT, typename... IDX_TYPES>
class Field : public FieldStorage<ELEMENT> {
 public:
  using Parent = FieldStorage<ELEMENT>;
  using ElementType = ELEMENT;
  static constexpr size_t NumDims = sizeof...(IDX_TYPES);
  using DefaultViewType = View<ELEMENT, NumDims>;

  Field(Set const &spanned_set, LayoutV layout) : Parent(spanned_set) {
    KRIPKE_ASSERT(NumDims == spanned_set.getNumDimensions(), "Number of dimensions must match between Field<%d> and Set<%d>\n",
                  (int)NumDims, (int)spanned_set.getNumDimensions());
    const int kinds[NumDims] = {IDX_TYPES::kind...};
    // permutation: stable sort of the index types by their DefaultOrder rank (Core/VarLayout.h:98-118)
    std::array<size_t, NumDims> perm;
    for (size_t i = 0; i < NumDims; ++i) perm[i] = i;
    for (size_t i = 1; i < NumDims; ++i) {
      size_t v = perm[i];
      long j = (long)i - 1;
      while (j >= 0 && layoutRank(layout, kinds[perm[j]]) > layoutRank(layout, kinds[v])) { perm[j + 1] = perm[j]; --j; }
      perm[j + 1] = v;
    }
    size_t num_chunks = this->m_chunk_to_subdomain.size();
    m_chunk_to_strides.resize(num_chunks);
    for (size_t c = 0; c < num_chunks; ++c) {
      SdomId sdom_id(this->m_chunk_to_subdomain[c]);
      long s = 1;
      for (long d = (long)NumDims - 1; d >= 0; --d) {
        m_chunk_to_strides[c][perm[d]] = s;
        s *= (long)spanned_set.dimSize(sdom_id, perm[d]);
      }
    }
  }
  // host view for reading AND writing (migrates the chunk to the host)
  DefaultViewType getView(SdomId sdom_id) {
    return DefaultViewType{this->getData(sdom_id), m_chunk_to_strides[this->m_subdomain_to_chunk[*sdom_id]]};
  }
  std::array<long, NumDims> const &getStrides(SdomId sdom_id) const {
    return m_chunk_to_strides[this->m_subdomain_to_chunk[*sdom_id]];
  }
 protected:
  std::vector<std::array<long, NumDims>> m_chunk_to_strides;
};

}  // namespace Core

// ---- Field typedefs: src/Kripke/VarTypes.h:32-64, PartitionSpace.h:97-99 -----------------------------
using Field_Flux = Core::Field<double, Direction, Group, Zone>;
using Field_Moments = Core::Field<double, Moment, Group, Zone>;
using Field_IPlane = Core::Field<double, Direction, Group, ZoneJ, ZoneK>;
using Field_JPlane = Core::Field<double, Direction, Group, ZoneI, ZoneK>;
using Field_KPlane = Core::Field<double, Direction, Group, ZoneI, ZoneJ>;
using Field_Ell = Core::Field<double, Moment, Direction>;
using Field_EllPlus = Core::Field<double, Direction, Moment>;
using Field_SigmaS = Core::Field<double, Material, Legendre, GlobalGroup, GlobalGroup>;
using Field_Direction2Double = Core::Field<double, Direction>;
using Field_Direction2Int = Core::Field<int, Direction>;
using Field_Adjacency = Core::Field<long, Dimension>;  // holds GlobalSdomId values
using Field_Moment2Legendre = Core::Field<int, Moment>;
using Field_ZoneI2Double = Core::Field<double, ZoneI>;
using Field_ZoneJ2Double = Core::Field<double, ZoneJ>;
using Field_ZoneK2Double = Core::Field<double, ZoneK>;
using Field_Zone2Double = Core::Field<double, Zone>;
using Field_Zone2Int = Core::Field<int, Zone>;
using Field_Zone2MixElem = Core::Field<int, Zone>;
using Field_MixElem2Double = Core::Field<double, MixElem>;
using Field_MixElem2Material = Core::Field<int, MixElem>;
using Field_MixElem2Zone = Core::Field<int, MixElem>;
using Field_SigmaTZonal = Core::Field<double, Group, Zone>;
using Field_SdomId2GlobalSdomId = Core::Field<long, SdomId>;
using Field_GlobalSdomId2Rank = Core::Field<long, GlobalSdomId>;
using Field_GlobalSdomId2SdomId = Core::Field<long, GlobalSdomId>;

class ArchLayout : public Core::BaseVar {  // src/Kripke/ArchLayout.h:139-145
 public:
  ArchLayoutV al_v;
};

// src/Kripke/VarTypes.h:139-152
template <typename FieldType, typename SetType>
FieldType &createField(Core::DataStore &data_store, std::string const &name, ArchLayoutV al_v, SetType const &set) {
  FieldType *field = new FieldType(set, al_v.layout_v);
  data_store.addVariable(name, field);
  return *field;
}

}  // namespace Kripke

// ---- InputVariables: src/Kripke/InputVariables.h ------------------------------------------------------
struct InputVariables {
  InputVariables();
  bool checkValues(void) const;
  int nx, ny, nz;
  int num_directions;
  int num_groups;
  int legendre_order;
  int quad_num_polar;
  int quad_num_azimuthal;
  Kripke::ArchLayoutV al_v;
  int npx, npy, npz;
  int num_dirsets;
  int num_groupsets;
  int num_zonesets_dim[3];
  int niter;
  ParallelMethod parallel_method;
  double sigt[3];
  double sigs[3];
  int num_material_subsamples;
  std::string run_name;
};
// parses the reference's command line (src/kripke.cpp:297-390); returns 0, or 1 after printing usage
int parseCommandLine(int argc, char **argv, InputVariables &vars);
void printUsage();

namespace Kripke {

// ---- Timing: src/Kripke/Timing.h ------------------------------------------------------------------------
// Wall-clock timers bracketed by a device synchronisation (kernels are asynchronous here, the
// reference's are synchronous) plus, per timer, the device time measured with CUDA events.
class Timing : public Core::BaseVar {
 public:
  virtual ~Timing();
  void start(std::string const &name);
  void stop(std::string const &name);
  void stopAll(void);
  void print(void) const;
  double getTotal(std::string const &name) const;
  size_t getCount(std::string const &name) const;
  static bool s_sync;  // synchronise the device at start/stop (default true)
 private:
  struct Timer {
    bool started = false;
    double elapsed = 0.0, t0 = 0.0;
    size_t count = 0;
  };
  std::map<std::string, Timer> timers;
  // nested regions, as the reference's Caliper annotations record them (Timing.h:93-109: every KRIPKE_TIMER is also a
  // nested "kripke" region): inclusive seconds per call path.  Printed as a runtime-report style tree when CALI_CONFIG or
  // CALI_CONFIG_PROFILE asks for "runtime-report" (README.md:182-186 of the reference).
  std::vector<std::string> m_stack;                 // currently open regions, outermost first
  std::vector<std::pair<std::string, double>> m_open;  // their paths and start times
  std::map<std::string, double> m_region_seconds;   // "Solve/SweepSolver/SweepSubdomain" -> inclusive seconds
  std::vector<std::string> m_region_order;          // paths in order of first entry
 public:
  void printRegions(void) const;
};
class BlockTimer {
 public:
  BlockTimer(Timing &timer_obj, std::string const &timer_name) : timer(timer_obj), name(timer_name) { timer.start(name); }
  ~BlockTimer() { timer.stop(name); }
 private:
  Timing &timer;
  std::string name;
};
#define KRIPKE_TIMER(DS, NAME) Kripke::BlockTimer BLK_TIMER_##NAME(DS.getVariable<Kripke::Timing>("timing"), #NAME);

// ---- generation: src/Kripke/Generate.h ---------------------------------------------------------------------
void generateProblem(Core::DataStore &data_store, InputVariables const &input_vars);
namespace Generate {
void generateDecomp(Core::DataStore &data_store, InputVariables const &input_vars);
void generateEnergy(Core::DataStore &data_store, InputVariables const &input_vars);
void generateQuadrature(Core::DataStore &data_store, InputVariables const &input_vars);
void generateSpace(Core::DataStore &data_store, InputVariables const &input_vars);
void generateData(Core::DataStore &data_store, InputVariables const &input_vars);
}  // namespace Generate

// ---- kernels: src/Kripke/Kernel.h:19-81 ----------------------------------------------------------------------
namespace Kernel {
void LPlusTimes(Core::DataStore &data_store);
void LTimes(Core::DataStore &data_store);
double population(Core::DataStore &data_store);
// fold_source: the caller promises that Kernel::source(data_store) is the next thing to touch phi_out (as in
// SteadyStateSolver.cpp:59-65); the source term is then added by the scattering kernel's epilogue and that source call
// only checks the promise
void scattering(Core::DataStore &data_store, bool fold_source = false);
void source(Core::DataStore &data_store);
void sweepSubdomain(Core::DataStore &data_store, SdomId sdom_id);
// batched form: all subdomains of the list must be mutually independent (used by SweepSolver)
void sweepSubdomains(Core::DataStore &data_store, std::vector<SdomId> const &sdom_ids, bool deliver_downwind,
                     bool fuse_population = false);
// Population fused into the sweep: SweepSolver brackets its sweeps with begin/end; population()
// returns the sum the sweep kernels accumulated if psi has not been touched since, else recomputes.
void populationFusionBegin(Core::DataStore &data_store);
void populationFusionEnd(Core::DataStore &data_store, std::vector<SdomId> const &swept);

template <typename FieldType>
void kConst(FieldType &field, SdomId sdom_id, typename FieldType::ElementType value) {
  if (value == typename FieldType::ElementType(0)) {
    field.setZeroPending(sdom_id);
    return;
  }
  size_t n = field.size(sdom_id);
  auto *p = field.getData(sdom_id);
  for (size_t i = 0; i < n; ++i) p[i] = value;
}
template <typename FieldType>
void kConst(FieldType &field, typename FieldType::ElementType value) {
  for (SdomId sdom_id : field.getWorkList()) kConst(field, sdom_id, value);
}
// copies field_src -> field_dst.  (The reference's kCopy assigns in the opposite direction,
// Kernel.h:71 -- a defect; this one does what its name says.)
template <typename FieldType>
void kCopy(FieldType &field_dst, SdomId sdom_id_dst, FieldType &field_src, SdomId sdom_id_src) {
  size_t n = field_src.size(sdom_id_src);
  if (field_src.isZeroPending(sdom_id_src)) { field_dst.setZeroPending(sdom_id_dst); return; }
  KB200_CALL(kb200_copy(field_dst.devicePtrOverwrite(sdom_id_dst), field_src.devicePtrConst(sdom_id_src),
                        n * sizeof(typename FieldType::ElementType), nullptr));
}
template <typename FieldType>
void kCopy(FieldType &field_dst, FieldType &field_src) {
  for (SdomId sdom_id : field_dst.getWorkList()) kCopy(field_dst, sdom_id, field_src, sdom_id);
}
}  // namespace Kernel

// ---- exchange over NVLink peer memory (one process per GPU, same node) -------------------------------------------
// Owned by the DataStore ("kb200/p2p"): the plane chunks of this rank that an off-rank upwind subdomain writes are
// exported once (CUDA IPC), the peers' chunks this rank writes are mapped, and every (subdomain, dimension) face has a
// flag in the receiver's memory.  The sweep kernel stores outgoing faces straight into the mapped chunk
// (kb200_sweep_desc.out_plane); PeerExchange::signal / wait order the two GPUs' streams.  Replaces the MPI_Irecv /
// MPI_Isend / MPI_Testany triple of src/Kripke/ParallelComm.cpp:61-251.
class PeerExchange : public Core::BaseVar {
 public:
  // collective over all ranks; returns nullptr when peer mapping is unavailable (then NCCL send/recv is used)
  static PeerExchange *get(Core::DataStore &data_store);
  ~PeerExchange() override;
  bool usable() const { return m_usable; }
  // mapped pointer of the downwind subdomain's plane chunk on its GPU (nullptr: on-rank or no neighbour)
  double *outPlane(SdomId sdom_id, int dim) const { return m_out_ptr[3 * (size_t)*sdom_id + dim]; }
  void beginSweep() { ++m_epoch; }
  // raise the flags of the faces the given subdomains have just produced / hold the stream until the faces the given
  // subdomains need have arrived
  void signal(std::vector<SdomId> const &produced);
  void wait(std::vector<SdomId> const &needed);
 private:
  PeerExchange() = default;
  bool m_usable = false;
  unsigned m_epoch = 0;
  unsigned *m_flags = nullptr;              // [3 * local subdomains] on this GPU
  std::vector<double *> m_out_ptr;          // [3 * local subdomains] mapped downwind chunk
  std::vector<unsigned *> m_out_flag;       // [3 * local subdomains] mapped flag of that face
  std::vector<char> m_in_offrank;           // [3 * local subdomains] the upwind face comes from another rank
  std::vector<void *> m_opened;             // mapped bases to close
};

// ---- exchange: src/Kripke/ParallelComm.h ------------------------------------------------------------------------
class ParallelComm {
 public:
  explicit ParallelComm(Core::DataStore &data_store);
  virtual ~ParallelComm() = default;
  virtual void addSubdomain(Core::DataStore &data_store, SdomId sdom_id) = 0;
  virtual bool workRemaining(void);
  virtual std::vector<SdomId> readySubdomains(void) = 0;
  virtual void markComplete(SdomId sdom_id) = 0;
  // the sweep kernel already wrote the outgoing faces into the on-rank downwind chunks
  void setDelivered(bool d) { m_delivered = d; }
 protected:
  int findSubdomain(SdomId sdom_id);
  void dequeueSubdomain(SdomId sdom_id);
  void postRecvs(Core::DataStore &data_store, SdomId sdom_id);
  void postSends(Core::DataStore &data_store, SdomId sdom_id, Core::FieldStorage<double> *src_fields[3]);
  void testRecieves(void);
  void waitAllSends(void);
  std::vector<SdomId> getReadyList(void);
  Core::DataStore *m_data_store;
  Core::FieldStorage<double> *m_plane_data[3];
  std::vector<int> queue_sdom_ids;
  std::vector<int> queue_depends;
  bool m_delivered = false;
};
class SweepComm : public ParallelComm {
 public:
  explicit SweepComm(Core::DataStore &data_store);
  virtual ~SweepComm();
  void addSubdomain(Core::DataStore &data_store, SdomId sdom_id) override;
  bool workRemaining(void) override;
  std::vector<SdomId> readySubdomains(void) override;
  void markComplete(SdomId sdom_id) override;
};
class BlockJacobiComm : public ParallelComm {
 public:
  explicit BlockJacobiComm(Core::DataStore &data_store);
  virtual ~BlockJacobiComm();
  void addSubdomain(Core::DataStore &data_store, SdomId sdom_id) override;
  bool workRemaining(void) override;
  std::vector<SdomId> readySubdomains(void) override;
  void markComplete(SdomId sdom_id) override;
 private:
  bool posted_sends;
};

// ---- solvers: src/Kripke/SweepSolver.h:20-22, SteadyStateSolver.h:18 -----------------------------------------------
void SweepSolver(Core::DataStore &data_store, std::vector<SdomId> subdomain_list, bool block_jacobi);
// the wavefront schedule SweepSolver runs (also what kripke_b200_sweep_schedule reports): depth of a local subdomain in the
// sweep DAG of its octant over the GLOBAL zone-set grid, and the number of global stages
int sweepDepth(Core::DataStore &data_store, Core::PartitionSpace const &pspace, SdomId local);
int numStages(Core::PartitionSpace const &pspace);
int SteadyStateSolver(Core::DataStore &data_store, size_t max_iter, bool block_jacobi);
// per-iteration particle counts of the last SteadyStateSolver call (full precision; stdout only has %e)
std::vector<double> const &lastParticleCounts();

}  // namespace Kripke

#endif
