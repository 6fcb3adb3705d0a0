// How many warps per SM sub-partition does the fp64 tensor pipe need?  mma.sync.m8n8k4.f64 chains per warp (8 or 16
// independent accumulators), optionally with the operand traffic of the scattering inner loop (3 ld.shared.v2.f64 +
// 2 DMUL per 8 DMMAs).  Prints TFLOP/s per (warps per sub-partition, chains, with loads).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int CH, bool LOADS>
__global__ void k(double *out, int iters, double seed) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = seed * 1e-3 * (i & 7);
  __syncthreads();
  double acc[CH][2];
  for (int c = 0; c < CH; ++c) acc[c][0] = acc[c][1] = 0.0;
  double a0 = seed, b0 = seed * 0.5;
  const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 16;
  for (int it = 0; it < iters; ++it) {
    if (LOADS) {
#pragma unroll
      for (int h = 0; h < CH / 8; ++h) {
        double2 bv, a01, a23;
        const unsigned ad = base + ((it * 2 + h) & 15) * 1536;
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(bv.x), "=d"(bv.y) : "r"(ad));
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a01.x), "=d"(a01.y) : "r"(ad + 512));
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a23.x), "=d"(a23.y) : "r"(ad + 1024));
        const double b0s = bv.x * seed, b1s = bv.y * seed;
        dmma(acc[8 * h + 0][0], acc[8 * h + 0][1], a01.x, b0s);
        dmma(acc[8 * h + 1][0], acc[8 * h + 1][1], a01.x, b1s);
        dmma(acc[8 * h + 2][0], acc[8 * h + 2][1], a01.y, b0s);
        dmma(acc[8 * h + 3][0], acc[8 * h + 3][1], a01.y, b1s);
        dmma(acc[8 * h + 4][0], acc[8 * h + 4][1], a23.x, b0s);
        dmma(acc[8 * h + 5][0], acc[8 * h + 5][1], a23.x, b1s);
        dmma(acc[8 * h + 6][0], acc[8 * h + 6][1], a23.y, b0s);
        dmma(acc[8 * h + 7][0], acc[8 * h + 7][1], a23.y, b1s);
      }
    } else {
#pragma unroll
      for (int c = 0; c < CH; ++c) dmma(acc[c][0], acc[c][1], a0, b0);
    }
  }
  double s = 0;
  for (int c = 0; c < CH; ++c) s += acc[c][0] + acc[c][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH, bool LOADS>
void run(int wps, double *out) {
  const int threads = 32 * 4 * wps, blocks = 148, iters = 20000 / (CH / 8);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(a);
    k<CH, LOADS><<<blocks, threads, 32768>>>(out, iters, 1.0);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  const double flops = (double)blocks * (threads / 32) * (double)iters * CH * 512.0;
  printf("warps/subpartition %2d  chains %2d  loads %d : %7.2f TFLOP/s  (%s)\n", wps, CH, (int)LOADS, flops / (best * 1e-3) * 1e-12,
         cudaGetErrorString(cudaGetLastError()));
}
int main() {
  double *out; cudaMalloc(&out, sizeof(double) * 148 * 1024);
  for (int wps : {1, 2, 3, 4, 8}) {
    run<8, false>(wps, out); run<16, false>(wps, out); run<8, true>(wps, out); run<16, true>(wps, out);
  }
  return 0;
}
