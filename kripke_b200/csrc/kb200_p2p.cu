// Face exchange over NVLink peer memory (CUDA IPC) for one process per GPU.
//
// Replaces ParallelComm::postSends / postRecvs / testRecieves (src/Kripke/ParallelComm.cpp:61-251) for neighbours that
// live on another GPU of the same node: the receiver's plane chunk is mapped into the sender's address space once
// (cudaIpcGetMemHandle / cudaIpcOpenMemHandle -- MPI_Irecv "straight into the downwind plane chunk", :99-107), the sweep
// kernel of the sender stores its outgoing faces directly there through kb200_sweep_desc.out_plane while it computes (the
// transfer rides under the sweep, MPI_Isend :170-178), and completion is a flag in the receiver's memory: the sender's
// stream raises it after the kernel, the receiver's stream waits for it before the dependent sweep (MPI_Testany :222-230).
// No host synchronisation and no staging copies are involved.
#include "kb200_common.cuh"

namespace kb200 {

__global__ void p2p_signal_kernel(unsigned *const *__restrict__ flags, int n, unsigned value) {
  // the stores of every earlier kernel of this stream have been performed; make that hold system-wide before the flag
  __threadfence_system();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    unsigned *f = flags[i];
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(value) : "memory");
  }
}

__global__ void p2p_wait_kernel(const unsigned *const *__restrict__ flags, int n, unsigned value) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const unsigned *f = flags[i];
    unsigned v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    } while ((int)(v - value) < 0);  // epochs only grow (wrap-safe comparison)
  }
  __threadfence_system();
}

}  // namespace kb200

using namespace kb200;

extern "C" int kb200_ipc_export(const void *d_ptr, void *handle64) {
  KB_REQUIRE(d_ptr && handle64, "kb200_ipc_export: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  KB_CUDA(cudaIpcGetMemHandle(&h, const_cast<void *>(d_ptr)));
  memcpy(handle64, &h, sizeof(h));
  return 0;
}

extern "C" int kb200_ipc_open(const void *handle64, void **d_peer_ptr) {
  KB_REQUIRE(d_peer_ptr && handle64, "kb200_ipc_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  KB_CUDA(cudaIpcOpenMemHandle(d_peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

extern "C" int kb200_ipc_close(void *d_peer_ptr) {
  if (d_peer_ptr) KB_CUDA(cudaIpcCloseMemHandle(d_peer_ptr));
  return 0;
}

// *flag = value on every listed flag (device pointers, typically peer memory), after all earlier work of the stream
extern "C" int kb200_p2p_signal(unsigned *const *h_flags, int n, unsigned value, kb200_stream_t stream) {
  if (n <= 0) return 0;
  cudaStream_t st = resolve_stream(stream);
  const void *d = nullptr;
  int rc = device_descs(h_flags, sizeof(unsigned *) * (size_t)n, &d, st);
  if (rc) return rc;
  p2p_signal_kernel<<<1, 64, 0, st>>>((unsigned *const *)d, n, value);
  return post_launch("p2p_signal");
}

// later work of the stream starts only when every listed flag (local device memory) has reached `value`
extern "C" int kb200_p2p_wait(const unsigned *const *h_flags, int n, unsigned value, kb200_stream_t stream) {
  if (n <= 0) return 0;
  cudaStream_t st = resolve_stream(stream);
  const void *d = nullptr;
  int rc = device_descs(h_flags, sizeof(unsigned *) * (size_t)n, &d, st);
  if (rc) return rc;
  p2p_wait_kernel<<<1, 64, 0, st>>>((const unsigned *const *)d, n, value);
  return post_launch("p2p_wait");
}
