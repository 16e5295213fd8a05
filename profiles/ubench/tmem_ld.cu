// Micro-benchmark (sm_100a): (A) tcgen05.ld 32x32b.x32 throughput per SM with 4 / 8 warps, alone and with the 16 FMNMX3 a
// row-minimum epilogue spends per load; (B) issue rate of tcgen05.mma kind::tf32 M128 N256 K8 from shared-memory operands.
// Decides whether a tensor-core distance filter for the Chamfer forward is read-bound or MMA-bound.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld tmem_ld.cu && ./tmem_ld
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ float min3(float a, float b, float c) { return fminf(fminf(a, b), c); }

#define LD32(taddr, r)                                                                                                    \
  asm volatile(                                                                                                           \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20," \
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                              \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),       \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),            \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),           \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                         \
      : "r"(taddr))

template <int MODE>  // 0 = loads only, 1 = loads + 16 FMNMX3 per load
__global__ void k_ld(float *out, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  float acc = 3.0e38f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c0 = 0; c0 < 128; c0 += 64) {
      uint32_t r[32], s[32];
      LD32(base + c0, r);
      LD32(base + c0 + 32, s);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (MODE == 0) {
        acc = fminf(acc, __uint_as_float(r[3] ^ s[17]));
      } else {
        float m0 = acc, m1 = acc, m2 = acc, m3 = acc;
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          m0 = min3(m0, __uint_as_float(r[e]), __uint_as_float(r[e + 1]));
          m1 = min3(m1, __uint_as_float(s[e]), __uint_as_float(s[e + 1]));
          m2 = min3(m2, __uint_as_float(r[e + 2]), __uint_as_float(r[e + 3]));
          m3 = min3(m3, __uint_as_float(s[e + 2]), __uint_as_float(s[e + 3]));
        }
        acc = min3(fminf(m0, m1), m2, m3);
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "n"(512));
}

__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3ffffu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}
__host__ __device__ constexpr uint32_t instr_desc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// one thread issues `per` MMAs (M128 N256 K8 tf32) per tile into alternating accumulators, commit + wait per tile
template <int N, int DEPTH>  // DEPTH accumulators of N columns in flight; DEPTH == 0: never wait (issue throughput)
__global__ void k_mma(float *out, int tiles, int per) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar[8];
  constexpr int NB = DEPTH > 0 ? DEPTH : 2;
  float *a = reinterpret_cast<float *>(smem);            // 128 x 8
  float *b = a + 128 * 8;                                // N x 8
  for (int i = threadIdx.x; i < (128 + N) * 8; i += blockDim.x) a[i] = 1.0f;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[i])), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0) {
    const uint64_t ad = smem_desc(smem_u32(a), 16 * 128, 128), bd = smem_desc(smem_u32(b), (N / 8) * 128, 128);
    constexpr uint32_t idesc = instr_desc_tf32(128, N);
    auto wait = [&](int t) {
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok)
                     : "r"(smem_u32(&bar[t % NB])), "r"((t / NB) & 1)
                     : "memory");
    };
    for (int t = 0; t < tiles; ++t) {
      const uint32_t d = slot + (t % NB) * N;
      for (int m = 0; m < per; ++m)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(d),
                     "l"(ad), "l"(bd), "r"(idesc), "r"(m)
                     : "memory");
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[t % NB])) : "memory");
      if (DEPTH > 0 && t >= DEPTH - 1) wait(t - (DEPTH - 1));  // the oldest tile in flight
      if (DEPTH == 0 && t >= 2) wait(t - 2);  // phases must not lap the barrier (two accumulators reused without draining)
    }
    for (int t = tiles - NB; t < tiles; ++t) wait(t);
    out[blockIdx.x] = 1.f;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "n"(512));
}

// MMA issue throughput without any wait in the loop: commit every `every` tiles to a barrier nobody waits on
template <int N>
__global__ void k_mma_nowait(float *out, int tiles, int per, int every) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar[2];
  float *a = reinterpret_cast<float *>(smem);
  float *b = a + 128 * 8;
  for (int i = threadIdx.x; i < (128 + N) * 8; i += blockDim.x) a[i] = 1.0f;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[0])), "r"(1) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[1])), "r"(every < 0 ? 2 : 1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int issuers = every < 0 ? 2 : 1;
  if (every < 0) every = 0;
  if ((threadIdx.x & 31) == 0 && warp < issuers) {
    const uint64_t ad = smem_desc(smem_u32(a), 16 * 128, 128), bd = smem_desc(smem_u32(b), (N / 8) * 128, 128);
    constexpr uint32_t idesc = instr_desc_tf32(128, N);
    for (int t = warp; t < tiles; t += issuers) {
      const uint32_t d = slot + (t % (512 / N)) * N;
      for (int m = 0; m < per; ++m)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(d),
                     "l"(ad), "l"(bd), "r"(idesc), "r"(m)
                     : "memory");
      if (every > 0 && (t % every) == every - 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[0])) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[1])) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                   : "=r"(ok)
                   : "r"(smem_u32(&bar[1])), "r"(0)
                   : "memory");
    out[blockIdx.x * 2 + warp] = 1.f;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "n"(512));
}

template <typename F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize(); float best = 1e30f;
  for (int r = 0; r < 5; ++r) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
  return best;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount; const double clk = 1.90e9;
  float *out; cudaMalloc(&out, sms * 256 * 4);
  const int iters = 20000;
  for (int threads = 128; threads <= 256; threads *= 2) {
    const float t0 = timeit([&] { k_ld<0><<<sms, threads>>>(out, iters); });
    const float t1 = timeit([&] { k_ld<1><<<sms, threads>>>(out, iters); });
    // per SM: iters * 4 loads of 4 KB per warp
    const double loads = (double)iters * 4 * (threads / 32);
    printf("{\"bench\": \"tcgen05.ld.32x32b.x32\", \"warps\": %d, \"cycles_per_load_per_sm\": %.2f, \"bytes_per_cycle_per_sm\": %.1f, "
           "\"with_16_fmnmx3_cycles_per_load_per_sm\": %.2f, \"cycles_per_128x256_tile\": %.1f, \"with_min_cycles_per_tile\": %.1f}\n",
           threads / 32, t0 * 1e-3 * clk / loads, 4096.0 / (t0 * 1e-3 * clk / loads), t1 * 1e-3 * clk / loads,
           t0 * 1e-3 * clk / loads * 32, t1 * 1e-3 * clk / loads * 32);
  }
  cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
  const int tiles = 20000;
#define RUN(N, DEPTH)                                                                                                     \
  for (int per = 1; per <= 4; ++per) {                                                                                   \
    const float t = timeit([&] { k_mma<N, DEPTH><<<sms, 128, (128 + N) * 8 * 4>>>(out, tiles, per); });                  \
    printf("{\"bench\": \"tcgen05.mma tf32 m128 k8\", \"n\": %d, \"tiles_in_flight\": %d, \"mmas_per_tile\": %d, "               \
           "\"cycles_per_tile\": %.1f, \"cycles_per_mma\": %.1f}\n", N, DEPTH, per, t * 1e-3 * clk / tiles,                    \
           t * 1e-3 * clk / tiles / per);                                                                                \
  }
  RUN(256, 2)
#define RUNNW(N, EVERY)                                                                                                   \
  for (int per = 1; per <= 4; per += 1) {                                                                                \
    const float t = timeit([&] { k_mma_nowait<N><<<sms, 128, (128 + N) * 8 * 4>>>(out, tiles, per, EVERY); });           \
    printf("{\"bench\": \"tcgen05.mma tf32 m128 k8, no wait in the loop\", \"n\": %d, \"commit_every_tiles\": %d, "                \
           "\"mmas_per_tile\": %d, \"cycles_per_tile\": %.1f, \"cycles_per_mma\": %.1f}\n", N, EVERY, per,                     \
           t * 1e-3 * clk / tiles, t * 1e-3 * clk / tiles / per);                                                       \
  }
  RUNNW(256, 1) RUNNW(256, -1) RUNNW(128, -1)
  e = cudaDeviceSynchronize(); if (e != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
  return 0;
}
