// Micro-benchmark: does non-FMA work co-issue with packed FFMA2 on sm_100a?  One SM-filling grid per variant;
// every variant runs NF packed FFMA2 per iteration plus K "other" instructions of one kind (FMNMX on the ALU pipe,
// warp REDUX.MIN, LDS.128, scalar FFMA) and reports SM cycles per iteration per SMSP-resident warp set.
// If the packed op held the dispatch port for both of its FMA-pipe cycles, cycles/iter = 2*NF + K; if the other
// instruction can slip into the second cycle, cycles/iter stays 2*NF until K exceeds NF.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_mix issue_mix.cu && ./issue_mix
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define FMA2(acc, a, b) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc) : "l"(a), "l"(b))
#define FMNMX(r, x) asm volatile("min.f32 %0, %0, %1;" : "+f"(r) : "f"(x))
#define FFMA(r, a, b) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(r) : "f"(a), "f"(b))
#define REDUX(r, x) asm volatile("redux.sync.min.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(x))
#define LDS128(a, b, c, d, addr) asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(addr))

constexpr int NF = 16;

template <int KIND, int K>
__global__ void __launch_bounds__(1024) kern(float *out, int iters, float fa, float fb, long long *cyc) {
  __shared__ __align__(16) float sm[1024];
  sm[threadIdx.x] = fa * threadIdx.x;
  __syncthreads();
  uint64_t acc[NF];
  float2 A = make_float2(fa, fa), B = make_float2(fb, fb);
  const uint64_t a = *(uint64_t *)&A, b = *(uint64_t *)&B;
  for (int i = 0; i < NF; ++i) { float2 t = make_float2(threadIdx.x + i, i); acc[i] = *(uint64_t *)&t; }
  float m[8];
  for (int i = 0; i < 8; ++i) m[i] = 1e30f + i;
  unsigned red[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float l0 = 0, l1 = 0, l2 = 0, l3 = 0;
  const unsigned saddr = (unsigned)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 0;  // broadcast read
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NF; ++i) {
      FMA2(acc[i], a, b);
      if (i < K) {
        if (KIND == 1) FMNMX(m[i & 7], fa);
        if (KIND == 2) REDUX(red[i & 7], (unsigned)it + i);
        if (KIND == 3) LDS128(l0, l1, l2, l3, saddr + 16 * (i & 7));
        if (KIND == 4) FFMA(m[i & 7], fa, fb);
      }
    }
  }
  const long long t1 = clock64();
  float s = l0 + l1 + l2 + l3;
  for (int i = 0; i < NF; ++i) { float2 t = *(float2 *)&acc[i]; s += t.x + t.y; }
  for (int i = 0; i < 8; ++i) s += m[i] + red[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int KIND, int K>
static double run(int warps_per_smsp, float *out, long long *cyc) {
  const int iters = 4000, threads = warps_per_smsp * 4 * 32;
  kern<KIND, K><<<148, threads>>>(out, iters, 1.0001f, 0.5f, cyc);
  cudaDeviceSynchronize();
  kern<KIND, K><<<148, threads>>>(out, iters, 1.0001f, 0.5f, cyc);
  cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  return (double)c / iters / warps_per_smsp;  // cycles per iteration per warp on one SMSP
}

int main() {
  float *out; long long *cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  printf("{\"nf_packed_per_iter\": %d, \"unit\": \"SM cycles per iteration per warp (SMSP time share)\"", NF);
  for (int w = 1; w <= 8; w *= 2) {
    printf(",\n \"warps_per_smsp_%d\": {", w);
    printf("\"ffma2_only\": %.2f", run<0, 0>(w, out, cyc));
    printf(", \"plus_4_fmnmx\": %.2f, \"plus_8_fmnmx\": %.2f, \"plus_16_fmnmx\": %.2f", run<1, 4>(w, out, cyc), run<1, 8>(w, out, cyc), run<1, 16>(w, out, cyc));
    printf(", \"plus_1_redux\": %.2f, \"plus_2_redux\": %.2f, \"plus_4_redux\": %.2f, \"plus_8_redux\": %.2f", run<2, 1>(w, out, cyc), run<2, 2>(w, out, cyc), run<2, 4>(w, out, cyc), run<2, 8>(w, out, cyc));
    printf(", \"plus_2_lds128\": %.2f, \"plus_4_lds128\": %.2f, \"plus_8_lds128\": %.2f", run<3, 2>(w, out, cyc), run<3, 4>(w, out, cyc), run<3, 8>(w, out, cyc));
    printf(", \"plus_8_ffma\": %.2f, \"plus_16_ffma\": %.2f}", run<4, 8>(w, out, cyc), run<4, 16>(w, out, cyc));
  }
  printf("}\n");
  return 0;
}
