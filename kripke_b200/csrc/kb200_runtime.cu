// Runtime half of the C ABI (include/kripke_b200.h): device binding, allocation, memory ops,
// streams/events, descriptor cache, NCCL exchange (dlopen'ed), peak micro-benchmarks.
//
// Replaces, for the hot path only: Core::Comm (src/Kripke/Core/Comm.h), the chunk allocation in
// Core::FieldStorage (src/Kripke/Core/Field.h:61-104), Kernel::kConst/kCopy (src/Kripke/Kernel.h:37-81)
// and the MPI calls of ParallelComm (src/Kripke/ParallelComm.cpp:61-251).
#include "kb200_common.cuh"
#include <unordered_map>
#include <vector>
#include <mutex>
#include <dlfcn.h>
#include <stdarg.h>
#include <stdlib.h>
#include <atomic>
#include <mutex>
#include <vector>

namespace kb200 {

static thread_local char g_err[1024] = "";
static cudaStream_t g_stream = nullptr;
static int g_device = -1;
static int g_sm_count = 0;
static std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char *what, const char *file, int line) {
  if (e == cudaSuccess) return 0;
  set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
  return 1000 + (int)e;
}

cudaStream_t resolve_stream(kb200_stream_t s) { return s ? (cudaStream_t)s : g_stream; }
void count_launch(int n) { g_launches += (uint64_t)n; }
static int g_exact = -1;
bool exact_mode() {
  if (g_exact < 0) {
    const char *e = getenv("KB200_EXACT");
    g_exact = (e && e[0] == '1') ? 1 : 0;
  }
  return g_exact == 1;
}
int sm_count() { return g_sm_count > 0 ? g_sm_count : 148; }

int post_launch(const char *kernel) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("launch of %s failed: %s", kernel, cudaGetErrorString(e));
    return 1000 + (int)e;
  }
  count_launch();
  return 0;
}

// ---- descriptor cache -------------------------------------------------------------------------
struct DescEntry {
  std::vector<unsigned char> host;
  void *dev = nullptr;
  uint64_t last_use = 0;
};
static std::vector<DescEntry> g_desc_cache;
static std::mutex g_desc_mutex;
static uint64_t g_desc_clock = 0;
static const size_t kDescCacheMax = 512;

int device_descs(const void *h, size_t bytes, const void **d_out, cudaStream_t stream) {
  std::lock_guard<std::mutex> lock(g_desc_mutex);
  ++g_desc_clock;
  for (auto &e : g_desc_cache)
    if (e.host.size() == bytes && memcmp(e.host.data(), h, bytes) == 0) {
      e.last_use = g_desc_clock;
      *d_out = e.dev;
      return 0;
    }
  if (g_desc_cache.size() >= kDescCacheMax) {  // evict least recently used (after the device is idle)
    size_t victim = 0;
    for (size_t i = 1; i < g_desc_cache.size(); ++i)
      if (g_desc_cache[i].last_use < g_desc_cache[victim].last_use) victim = i;
    KB_CUDA(cudaDeviceSynchronize());
    cudaFree(g_desc_cache[victim].dev);
    g_desc_cache.erase(g_desc_cache.begin() + victim);
  }
  DescEntry e;
  e.host.assign((const unsigned char *)h, (const unsigned char *)h + bytes);
  KB_CUDA(cudaMalloc(&e.dev, bytes));
  // synchronous w.r.t. the host so the caller's buffer may die; ordered before later launches
  KB_CUDA(cudaMemcpy(e.dev, h, bytes, cudaMemcpyHostToDevice));
  (void)stream;
  e.last_use = g_desc_clock;
  *d_out = e.dev;
  g_desc_cache.push_back(std::move(e));
  return 0;
}

static void clear_desc_cache() {
  std::lock_guard<std::mutex> lock(g_desc_mutex);
  for (auto &e : g_desc_cache) cudaFree(e.dev);
  g_desc_cache.clear();
}

__global__ void fill_f64_kernel(double *__restrict__ p, double v, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  // 128-bit stores on the aligned body
  size_t n2 = n / 2;
  if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    double2 vv = make_double2(v, v);
    double2 *p2 = reinterpret_cast<double2 *>(p);
    for (size_t k = i; k < n2; k += stride) p2[k] = vv;
    if (i == 0 && (n & 1)) p[n - 1] = v;
  } else {
    for (size_t k = i; k < n; k += stride) p[k] = v;
  }
}

// ---- fp64 peak micro-benchmarks --------------------------------------------------------------
__global__ void dfma_peak_kernel(double *out, int iters, double seed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
      a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void dmma_peak_kernel(double *out, int iters, double seed) {
  double a = seed + (threadIdx.x & 31) * 1e-3, b = 1.0 + (threadIdx.x & 31) * 1e-6;
  double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(a), "d"(b));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(a), "d"(b));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(a), "d"(b));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(a), "d"(b));
    }
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
}

__global__ void copy_peak_kernel(const double2 *__restrict__ src, double2 *__restrict__ dst, size_t n2) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n2; i += stride) dst[i] = src[i];
}

// ---- NCCL through dlopen (no link-time dependency; the same libnccl.so.2 torch uses if loaded) --
typedef struct { char internal[128]; } nccl_uid;
typedef void *nccl_comm;
struct Nccl {
  void *handle = nullptr;
  int (*GetUniqueId)(nccl_uid *) = nullptr;
  int (*CommInitRank)(nccl_comm *, int, nccl_uid, int) = nullptr;
  int (*CommDestroy)(nccl_comm) = nullptr;
  int (*Send)(const void *, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, nccl_comm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  nccl_comm comm = nullptr;
  int rank = 0, nranks = 1;
};
static Nccl g_nccl;
enum { NCCL_INT64 = 4, NCCL_FLOAT64 = 8, NCCL_SUM = 0 };

static int nccl_load() {
  if (g_nccl.handle) return 0;
  const char *cands[] = {getenv("KB200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char *c : cands) {
    if (!c) continue;
    g_nccl.handle = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle) break;
  }
  KB_REQUIRE(g_nccl.handle, "cannot dlopen libnccl.so.2 (set KB200_NCCL_LIB): %s", dlerror());
#define SYM(field, name)                                                  \
  *(void **)(&g_nccl.field) = dlsym(g_nccl.handle, name);                 \
  KB_REQUIRE(g_nccl.field, "libnccl: missing symbol %s", name)
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(AllReduce, "ncclAllReduce");
  SYM(AllGather, "ncclAllGather");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  return 0;
}
#define KB_NCCL(x)                                                                         \
  do {                                                                                     \
    int _r = (x);                                                                          \
    if (_r) {                                                                              \
      kb200::set_error("NCCL error %d (%s) in %s", _r, g_nccl.GetErrorString(_r), #x);     \
      return 2000 + _r;                                                                    \
    }                                                                                      \
  } while (0)

}  // namespace kb200

using namespace kb200;

extern "C" {

int kb200_abi_version(void) { return KB200_ABI_VERSION; }
int kb200_set_exact(int on) {
  g_exact = on ? 1 : 0;
  return 0;
}
const char *kb200_last_error(void) { return g_err; }

int kb200_device_count(int *count) {
  KB_REQUIRE(count, "kb200_device_count: null argument");
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) {
    *count = 0;
    cudaGetLastError();
    set_error("no CUDA device visible: %s", cudaGetErrorString(e));
    return 1000 + (int)e;
  }
  return 0;
}

int kb200_init(int device) {
  int n = 0;
  int rc = kb200_device_count(&n);
  if (rc) return rc;
  KB_REQUIRE(n > 0, "kb200_init: no CUDA device visible (this library has no CPU fallback)");
  KB_REQUIRE(device >= 0 && device < n, "kb200_init: device %d out of range (0..%d)", device, n - 1);
  KB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  KB_CUDA(cudaGetDeviceProperties(&prop, device));
  KB_REQUIRE(prop.major == 10, "kb200_init: device %d is sm_%d%d; this library is built for sm_100a only",
             device, prop.major, prop.minor);
  g_device = device;
  g_sm_count = prop.multiProcessorCount;
  if (!g_stream) KB_CUDA(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
  return 0;
}

namespace { void clear_pool(); }

int kb200_finalize(void) {
  if (g_device < 0) return 0;
  cudaDeviceSynchronize();
  clear_desc_cache();
  clear_pool();
  if (g_nccl.comm) {
    g_nccl.CommDestroy(g_nccl.comm);
    g_nccl.comm = nullptr;
  }
  if (g_stream) {
    cudaStreamDestroy(g_stream);
    g_stream = nullptr;
  }
  g_device = -1;
  return 0;
}

int kb200_device_info(char *name, size_t name_len, int *sms, int *maj, int *min, size_t *free_b, size_t *total_b) {
  KB_REQUIRE(g_device >= 0, "kb200_device_info: call kb200_init first");
  cudaDeviceProp prop;
  KB_CUDA(cudaGetDeviceProperties(&prop, g_device));
  if (name && name_len) {
    strncpy(name, prop.name, name_len - 1);
    name[name_len - 1] = 0;
  }
  if (sms) *sms = prop.multiProcessorCount;
  if (maj) *maj = prop.major;
  if (min) *min = prop.minor;
  size_t f = 0, t = 0;
  KB_CUDA(cudaMemGetInfo(&f, &t));
  if (free_b) *free_b = f;
  if (total_b) *total_b = t;
  return 0;
}

// Small allocation pool: fields that live for one SweepSolver call (the block-Jacobi "old" planes, src/Kripke/
// ParallelComm/BlockJacobiComm.cpp:26-44) would otherwise pay a cudaMalloc + a synchronising cudaFree per chunk and
// iteration (measured: 14-30 ms per iteration at 16 subdomains).  Blocks up to kPoolMaxBlock are kept by exact size
// and handed out again; all work runs on the library's stream, so reuse is stream-ordered.
namespace {
constexpr size_t kPoolMaxBlock = 64u << 20, kPoolMaxTotal = 4ull << 30;
std::atomic<int> g_user_streams{0};  // streams handed out by kb200_stream_create and not yet destroyed
std::mutex g_pool_mutex;
std::unordered_map<void *, size_t> g_pool_sizes;            // live blocks that may be pooled when freed
std::unordered_map<size_t, std::vector<void *>> g_pool_free;
size_t g_pool_bytes = 0;
void clear_pool() {
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  for (auto &kv : g_pool_free)
    for (void *q : kv.second) cudaFree(q);
  g_pool_free.clear();  // live blocks stay tracked in g_pool_sizes: they may still be pooled when freed
  g_pool_bytes = 0;
}
}  // namespace

int kb200_alloc(size_t bytes, void **p) {
  KB_REQUIRE(p, "kb200_alloc: null argument");
  KB_REQUIRE(g_device >= 0, "kb200_alloc: call kb200_init first");
  *p = nullptr;
  if (bytes == 0) bytes = 16;
  if (bytes <= kPoolMaxBlock) {
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    auto it = g_pool_free.find(bytes);
    if (it != g_pool_free.end() && !it->second.empty()) {
      *p = it->second.back();
      it->second.pop_back();
      g_pool_bytes -= bytes;
      g_pool_sizes[*p] = bytes;
      return 0;
    }
  }
  cudaError_t e = cudaMalloc(p, bytes);
  if (e == cudaErrorMemoryAllocation) {  // give the pooled blocks back and retry once
    cudaGetLastError();
    cudaDeviceSynchronize();
    clear_pool();
    e = cudaMalloc(p, bytes);
  }
  KB_CUDA(e);
  if (bytes <= kPoolMaxBlock) {
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    g_pool_sizes[*p] = bytes;
  }
  return 0;
}
int kb200_free(void *p) {
  if (!p) return 0;
  // A pooled block can be handed to its next owner at once.  That is only safe if no work that still uses it is
  // outstanding: work on the library stream is ordered with the next owner's work on the same stream, but as soon as the
  // caller has created streams of its own (kb200_stream_create) a kernel on one of them may still be running -- then
  // behave like cudaFree and wait for the device first.
  if (g_user_streams.load() > 0) KB_CUDA(cudaDeviceSynchronize());
  {
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    auto it = g_pool_sizes.find(p);
    if (it != g_pool_sizes.end()) {
      const size_t bytes = it->second;
      g_pool_sizes.erase(it);
      if (g_pool_bytes + bytes <= kPoolMaxTotal) {
        g_pool_free[bytes].push_back(p);
        g_pool_bytes += bytes;
        return 0;
      }
    }
  }
  KB_CUDA(cudaFree(p));
  return 0;
}
// give the pooled (free) blocks back to the driver; live blocks are untouched
int kb200_pool_trim(void) {
  KB_CUDA(cudaDeviceSynchronize());
  clear_pool();
  return 0;
}
int kb200_alloc_host(size_t bytes, void **p) {
  KB_REQUIRE(p, "kb200_alloc_host: null argument");
  KB_CUDA(cudaMallocHost(p, bytes ? bytes : 16));
  return 0;
}
int kb200_free_host(void *p) {
  if (p) KB_CUDA(cudaFreeHost(p));
  return 0;
}
int kb200_upload(void *d, const void *h, size_t bytes, kb200_stream_t s) {
  if (!bytes) return 0;
  KB_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, resolve_stream(s)));
  return 0;
}
int kb200_download(void *h, const void *d, size_t bytes, kb200_stream_t s) {
  if (!bytes) return 0;
  KB_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, resolve_stream(s)));
  return 0;
}
int kb200_copy(void *dst, const void *src, size_t bytes, kb200_stream_t s) {
  if (!bytes) return 0;
  KB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, resolve_stream(s)));
  return 0;
}
int kb200_fill_f64(double *p, double v, size_t n, kb200_stream_t s) {
  if (!n) return 0;
  if (v == 0.0) {  // +0.0 is all-zero bytes: use the copy engine-free memset path
    KB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(double), resolve_stream(s)));
    return 0;
  }
  size_t want = (n / 2 + 255) / 256;
  int blocks = (int)(want < (size_t)sm_count() * 8 ? (want ? want : 1) : (size_t)sm_count() * 8);
  fill_f64_kernel<<<blocks, 256, 0, resolve_stream(s)>>>(p, v, n);
  return post_launch("fill_f64");
}
int kb200_stream_create(kb200_stream_t *s) {
  cudaStream_t st;
  KB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  *s = (kb200_stream_t)st;
  ++g_user_streams;
  return 0;
}
int kb200_stream_destroy(kb200_stream_t s) {
  if (s) {
    KB_CUDA(cudaStreamDestroy((cudaStream_t)s));
    --g_user_streams;
  }
  return 0;
}
int kb200_device_bound(void) { return g_device >= 0 ? 1 : 0; }
int kb200_memset(void *p, int byte_value, size_t bytes, kb200_stream_t s) {
  if (!bytes) return 0;
  KB_CUDA(cudaMemsetAsync(p, byte_value, bytes, resolve_stream(s)));
  return 0;
}
int kb200_stream_sync(kb200_stream_t s) {
  KB_CUDA(cudaStreamSynchronize(resolve_stream(s)));
  return 0;
}
int kb200_device_sync(void) {
  KB_CUDA(cudaDeviceSynchronize());
  return 0;
}
int kb200_event_create(kb200_event_t *ev) {
  cudaEvent_t e;
  KB_CUDA(cudaEventCreate(&e));
  *ev = (kb200_event_t)e;
  return 0;
}
int kb200_event_destroy(kb200_event_t ev) {
  if (ev) KB_CUDA(cudaEventDestroy((cudaEvent_t)ev));
  return 0;
}
int kb200_event_record(kb200_event_t ev, kb200_stream_t s) {
  KB_CUDA(cudaEventRecord((cudaEvent_t)ev, resolve_stream(s)));
  return 0;
}
int kb200_event_sync(kb200_event_t ev) {
  KB_CUDA(cudaEventSynchronize((cudaEvent_t)ev));
  return 0;
}
int kb200_event_elapsed_ms(kb200_event_t a, kb200_event_t b, float *ms) {
  KB_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b));
  return 0;
}
int kb200_stream_wait_event(kb200_stream_t s, kb200_event_t ev) {
  KB_CUDA(cudaStreamWaitEvent(resolve_stream(s), (cudaEvent_t)ev, 0));
  return 0;
}
int kb200_launch_count(uint64_t *count, int reset) {
  if (count) *count = g_launches.load();
  if (reset) g_launches = 0;
  return 0;
}

// ---- exchange ---------------------------------------------------------------------------------
int kb200_comm_unique_id(void *id128) {
  int rc = nccl_load();
  if (rc) return rc;
  nccl_uid id;
  KB_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return 0;
}
int kb200_comm_init(int rank, int nranks, const void *id128) {
  KB_REQUIRE(g_device >= 0, "kb200_comm_init: call kb200_init first");
  int rc = nccl_load();
  if (rc) return rc;
  nccl_uid id;
  memcpy(&id, id128, sizeof(id));
  KB_NCCL(g_nccl.CommInitRank(&g_nccl.comm, nranks, id, rank));
  g_nccl.rank = rank;
  g_nccl.nranks = nranks;
  return 0;
}
int kb200_comm_destroy(void) {
  if (g_nccl.comm) {
    KB_NCCL(g_nccl.CommDestroy(g_nccl.comm));
    g_nccl.comm = nullptr;
  }
  g_nccl.rank = 0;
  g_nccl.nranks = 1;
  return 0;
}
int kb200_comm_rank(int *rank, int *nranks) {
  if (rank) *rank = g_nccl.rank;
  if (nranks) *nranks = g_nccl.nranks;
  return 0;
}
int kb200_comm_group_start(void) {
  KB_REQUIRE(g_nccl.comm, "kb200_comm_group_start: no communicator");
  KB_NCCL(g_nccl.GroupStart());
  return 0;
}
int kb200_comm_group_end(void) {
  KB_REQUIRE(g_nccl.comm, "kb200_comm_group_end: no communicator");
  KB_NCCL(g_nccl.GroupEnd());
  return 0;
}
int kb200_comm_send(const double *buf, size_t count, int peer, kb200_stream_t s) {
  KB_REQUIRE(g_nccl.comm, "kb200_comm_send: no communicator");
  KB_NCCL(g_nccl.Send(buf, count, NCCL_FLOAT64, peer, g_nccl.comm, resolve_stream(s)));
  return 0;
}
int kb200_comm_recv(double *buf, size_t count, int peer, kb200_stream_t s) {
  KB_REQUIRE(g_nccl.comm, "kb200_comm_recv: no communicator");
  KB_NCCL(g_nccl.Recv(buf, count, NCCL_FLOAT64, peer, g_nccl.comm, resolve_stream(s)));
  return 0;
}
int kb200_comm_allreduce_sum_f64(double *buf, size_t count, kb200_stream_t s) {
  if (!g_nccl.comm || g_nccl.nranks == 1) return 0;
  KB_NCCL(g_nccl.AllReduce(buf, buf, count, NCCL_FLOAT64, NCCL_SUM, g_nccl.comm, resolve_stream(s)));
  return 0;
}
int kb200_comm_allreduce_sum_i64(long long *buf, size_t count, kb200_stream_t s) {
  if (!g_nccl.comm || g_nccl.nranks == 1) return 0;
  KB_NCCL(g_nccl.AllReduce(buf, buf, count, NCCL_INT64, NCCL_SUM, g_nccl.comm, resolve_stream(s)));
  return 0;
}

// host buffers: h_recv[r * bytes .. (r+1) * bytes) = rank r's h_send (setup-time exchange of IPC handles, MPI_Allgather's role)
int kb200_comm_allgather(const void *h_send, size_t bytes, void *h_recv) {
  if (!g_nccl.comm || g_nccl.nranks == 1) {
    memcpy(h_recv, h_send, bytes);
    return 0;
  }
  if (!bytes) return 0;
  cudaStream_t st = resolve_stream(nullptr);
  unsigned char *d = nullptr;
  KB_CUDA(cudaMalloc(&d, bytes * (size_t)(g_nccl.nranks + 1)));
  KB_CUDA(cudaMemcpyAsync(d, h_send, bytes, cudaMemcpyHostToDevice, st));
  KB_NCCL(g_nccl.AllGather(d, d + bytes, bytes, 0 /* ncclInt8 */, g_nccl.comm, st));
  KB_CUDA(cudaMemcpyAsync(h_recv, d + bytes, bytes * (size_t)g_nccl.nranks, cudaMemcpyDeviceToHost, st));
  KB_CUDA(cudaStreamSynchronize(st));
  KB_CUDA(cudaFree(d));
  return 0;
}
// stream-ordered barrier over all ranks (a one-element all-reduce)
int kb200_comm_barrier(kb200_stream_t s) {
  if (!g_nccl.comm || g_nccl.nranks == 1) return 0;
  static long long *d_one = nullptr;
  if (!d_one) {
    KB_CUDA(cudaMalloc(&d_one, sizeof(long long)));
    KB_CUDA(cudaMemset(d_one, 0, sizeof(long long)));
  }
  KB_NCCL(g_nccl.AllReduce(d_one, d_one, 1, NCCL_INT64, NCCL_SUM, g_nccl.comm, resolve_stream(s)));
  return 0;
}

// ---- peaks --------------------------------------------------------------------------------------
int kb200_peak_fp64_gflops(int use_dmma, int iters, double *gflops) {
  KB_REQUIRE(g_device >= 0, "kb200_peak_fp64_gflops: call kb200_init first");
  int blocks = sm_count() * 4, threads = 512;
  double *out;
  KB_CUDA(cudaMalloc(&out, sizeof(double) * blocks * threads));
  cudaEvent_t a, b;
  KB_CUDA(cudaEventCreate(&a));
  KB_CUDA(cudaEventCreate(&b));
  cudaStream_t st = resolve_stream(nullptr);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    KB_CUDA(cudaEventRecord(a, st));
    if (use_dmma) dmma_peak_kernel<<<blocks, threads, 0, st>>>(out, iters, 1.0);
    else dfma_peak_kernel<<<blocks, threads, 0, st>>>(out, iters, 1.0);
    KB_CUDA(cudaEventRecord(b, st));
    KB_CUDA(cudaEventSynchronize(b));
    float ms;
    KB_CUDA(cudaEventElapsedTime(&ms, a, b));
    if (rep > 0 && ms < best) best = ms;
  }
  double flops;
  if (use_dmma) flops = (double)blocks * (threads / 32) * (double)iters * 16.0 * (8 * 8 * 4 * 2);
  else flops = (double)blocks * threads * (double)iters * 64.0 * 2.0;
  *gflops = flops / (best * 1e-3) * 1e-9;
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(out);
  return post_launch("peak_fp64");
}

int kb200_peak_copy_gbs(size_t bytes, int iters, double *gbs) {
  KB_REQUIRE(g_device >= 0, "kb200_peak_copy_gbs: call kb200_init first");
  double2 *src, *dst;
  size_t n2 = bytes / sizeof(double2);
  KB_CUDA(cudaMalloc(&src, n2 * sizeof(double2)));
  KB_CUDA(cudaMalloc(&dst, n2 * sizeof(double2)));
  KB_CUDA(cudaMemset(src, 0, n2 * sizeof(double2)));
  cudaEvent_t a, b;
  KB_CUDA(cudaEventCreate(&a));
  KB_CUDA(cudaEventCreate(&b));
  cudaStream_t st = resolve_stream(nullptr);
  float best = 1e30f;
  for (int rep = 0; rep < iters + 1; ++rep) {
    KB_CUDA(cudaEventRecord(a, st));
    copy_peak_kernel<<<sm_count() * 16, 512, 0, st>>>(src, dst, n2);
    KB_CUDA(cudaEventRecord(b, st));
    KB_CUDA(cudaEventSynchronize(b));
    float ms;
    KB_CUDA(cudaEventElapsedTime(&ms, a, b));
    if (rep > 0 && ms < best) best = ms;
  }
  *gbs = 2.0 * (double)(n2 * sizeof(double2)) / (best * 1e-3) * 1e-9;
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(src);
  cudaFree(dst);
  return post_launch("peak_copy");
}

}  // extern "C"
