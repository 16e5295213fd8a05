// Micro-benchmark: throughput of warp REDUX (CREDUX) vs SHFL+FMNMX on sm_100a, per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o redux redux.cu && ./redux
#include <cuda_runtime.h>
#include <stdio.h>
__global__ void k_redux(unsigned *out, int iters) {
  unsigned v[8];
  for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * 2654435761u + i;
  unsigned acc = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc += __reduce_min_sync(0xffffffffu, v[i]); v[i] = v[i] * 1664525u + 1013904223u; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_base(unsigned *out, int iters) {  // same loop without the REDUX
  unsigned v[8];
  for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * 2654435761u + i;
  unsigned acc = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc += v[i]; v[i] = v[i] * 1664525u + 1013904223u; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_shfl(unsigned *out, int iters) {
  float v[8];
  for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * 0.37f + i;
  float acc = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = fminf(v[i] * 1.0001f + 0.5f, __shfl_xor_sync(0xffffffffu, v[i], 16)); }
    acc += v[0];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = __float_as_uint(acc + v[1] + v[2] + v[3] + v[4] + v[5] + v[6] + v[7]);
}
template <typename F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize(); float best = 1e30f;
  for (int r = 0; r < 5; ++r) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
  return best;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount; const int iters = 20000;
  unsigned *out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
  for (int warps = 4; warps <= 32; warps *= 2) {
    int threads = warps * 32 > 1024 ? 1024 : warps * 32, ctas = sms * (warps * 32 / threads);
    float tr = timeit([&] { k_redux<<<ctas, threads>>>(out, iters); });
    float tb = timeit([&] { k_base<<<ctas, threads>>>(out, iters); });
    float ts = timeit([&] { k_shfl<<<ctas, threads>>>(out, iters); });
    double clk = 1.965e9, n = (double)iters * 8 * warps;  // warp-level ops per SM
    printf("warps/SM %2d: redux loop %.2f cyc/op/SM (base loop %.2f), shfl+fma+fmnmx loop %.2f cyc/op/SM\n", warps,
           tr * 1e-3 * clk / n, tb * 1e-3 * clk / n, ts * 1e-3 * clk / n);
  }
  return 0;
}
