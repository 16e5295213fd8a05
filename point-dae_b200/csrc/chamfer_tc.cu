// chamfer_tc.cu -- Chamfer forward with the 5th-generation tensor cores as an exact FILTER (round 2).
//
// The reference (extensions/chamfer_dist/chamfer.cu:15-145) evaluates every pair (a_i, b_j) in fp32 on the CUDA cores and
// keeps, per point, the minimum distance and its lowest index; chamfer.cu here does the same on the FP32 FMA pipe and is
// bound by it (6 lane-ops per pair).  This file gets the same bits from ~1/4 of the issue slots:
//
//   1. tcgen05.mma (kind::tf32, M128 x N128|256 x K8, accumulators in TENSOR MEMORY) evaluates an APPROXIMATE
//      D[i][j] = |b'_j|^2 - 2 a'_i . b'_j   (= |a'_i - b'_j|^2 - |a'_i|^2: the row constant does not move a row's argmin)
//      for centred points a' = a - c, b' = b - c.  Every coordinate is split into tf32 hi + lo; one K = 8 operand row
//      carries [ah.x ah.y ah.z al.x | al.y al.z 1 1] against [-2bh.x -2bh.y -2bh.z -2bh.x | -2bh.y -2bh.z m_hi m_lo] and a
//      second MMA adds [.. same A ..] x [-2bl.x -2bl.y -2bl.z -2bl.x | -2bl.y -2bl.z 0 0]: all four hi/lo cross terms, so
//      |D - exact| <= eps = eps_rel * (max|a'|^2 + max|b'|^2) with eps_rel = 2^-16 by default (measured error: see
//      DESIGN.md 4.1b; the tune hook lowers eps_rel until results change, which is how the margin is tested).
//   2. The epilogue warps read D with tcgen05.ld (32 columns per instruction, a thread = a row), reduce every 32-column
//      group to its minimum g with FMNMX3 and keep, per row, the groups with g <= (running best) + 2 eps -- a superset of
//      the groups that can hold the exact argmin; a new best more than 2 eps below the old one clears the list, so it
//      holds one or two entries.
//   3. The surviving groups (32 columns each, ~1.0 per row) are evaluated EXACTLY -- the reference's own expression
//      fma(dz,dz, fma(dx,dx, dy*dy)) on the original coordinates -- and the (distance bits, index) minimum is the
//      reference's result bit for bit (strict `<`: lowest index on ties).  Rows whose list overflowed (mass ties), rows
//      without a finite candidate and clouds with non-finite bounds are scanned exactly over all columns: the result
//      never depends on the filter being right, only its speed does.
//
// Persistent CTAs (one per SM: a CTA owns all 512 TMEM columns), each walks a contiguous range of 128-row blocks; the
// operand image of the whole reference cloud (<= 2048 points, 128 KB) is built once per (cloud, direction) and stays in
// shared memory.  Warp 8 builds the row operands and issues the MMAs; warps 0-7 are the epilogue (warp w reads TMEM lanes
// 32 (w & 3).., columns of half w >> 2); accumulator buffers cycle through full / empty mbarriers (tcgen05.commit).
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"

namespace pdae {

namespace tcc {
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
  }
}
// busy poll (no suspended wait): the MMA issuer and the accumulator consumers sit on the pipeline's critical path
__device__ __forceinline__ void mbar_poll(uint64_t *bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// K-major, no swizzle: 8-row x 16-byte core matrices; LBO = bytes between the two 16-byte K chunks, SBO = between 8-row groups
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
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D fp32, A / B fp16 (K = 16 per instruction), both K-major
__host__ __device__ constexpr uint32_t instr_desc_f16(int m, int n) {
  return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// v = hi + lo + r with hi, lo fp16 and |r| <= 2^-24 for |v| <= 1 (fp16 subnormals carry the small lo parts)
__device__ __forceinline__ void split_h(float v, uint32_t &hi, uint32_t &lo) {
  const __half h = __float2half_rn(v);
  const __half l = __float2half_rn(__fsub_rn(v, __half2float(h)));
  hi = __half_as_ushort(h), lo = __half_as_ushort(l);
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float tf32_rn(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
#define PDAE_TMEM_LD32(taddr, r)                                                                                          \
  asm volatile(                                                                                                           \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20," \
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                              \
      : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]),       \
        "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15]), "=f"(r[16]),            \
        "=f"(r[17]), "=f"(r[18]), "=f"(r[19]), "=f"(r[20]), "=f"(r[21]), "=f"(r[22]), "=f"(r[23]), "=f"(r[24]),           \
        "=f"(r[25]), "=f"(r[26]), "=f"(r[27]), "=f"(r[28]), "=f"(r[29]), "=f"(r[30]), "=f"(r[31])                         \
      : "r"(taddr))
}  // namespace tcc

constexpr int TCC_M = 128;         // rows of a row block (MMA M, the 128 TMEM lanes)
constexpr int TCC_MAXCOLS = 2048;  // reference points whose operand image stays in shared memory
constexpr int TCC_GROUP = 32;      // columns per candidate group (one tcgen05.ld.x32)
constexpr int TCC_CAP = 8;         // candidate groups kept per (row, column half)
constexpr int TCC_EPI = 256;       // epilogue threads (warps 0-7)
constexpr int TCC_VER_WARPS = 3;   // verifier warps (exact evaluation of the surviving groups)
constexpr int TCC_PROD_WARPS = 1;  // MMA issuer warps (two are supported: tile k goes to issuer k & 1; one keeps up with two accumulators)
constexpr int TCC_THREADS = TCC_EPI + 32 * TCC_PROD_WARPS + 32 * TCC_VER_WARPS;
constexpr int TCC_RSTRIDE = 36;    // floats per 32-column group in the shared-memory copy of the searched cloud
constexpr float TCC_EPS_TF32 = 7.62939453125e-6f;  // 2^-17: default error bound of the tf32 filter relative to the scale
                                                   // (largest observed error 2^-21.9, results change below 2^-26)
constexpr float TCC_EPS_F16 = 1.52587890625e-5f;   // 2^-16: ... of the fp16 filter (observed 2^-21.0, results change below 2^-23)
constexpr float TCC_BIG = 1.0e30f;  // "distance" of a padded column

struct TccDir {
  const float *q;  // (b, nq, 3) the points that receive a minimum (rows)
  const float *r;  // (b, nr, 3) the cloud that is searched (columns)
  float *dist;     // (b, nq)
  int *idx;        // (b, nq)
  uint64_t *keys;  // (b, nq) when the searched cloud is cut into column chunks: (distance bits << 32 | index), RED.MIN target
  int nq, nr, rbs;  // rbs = ceil(nq / 128)
  int nch, chunk;   // column chunks of the searched cloud (each <= 2048 points, a multiple of 256 but the last) and their size
  int idx_off;      // added to the indices in the keys (reference-set sharding: global index of r[0])
};
constexpr int TCC_MAX_BOUNDS = 160;  // CTAs + 1 of a cost-balanced launch (148 SMs on B200)
struct TccArgs {
  TccDir d[2];
  long long units;  // b * (d[0].rbs * d[0].nch + d[1].rbs * d[1].nch) row blocks x column chunks
  float eps_rel;
  unsigned long long *stats;  // optional probe: [0] max |g - (exact group minimum - |a'|^2)| / (max|a'|^2 + max|b'|^2) as float
                              // bits, [1] rows decided by the full scan, [2] groups evaluated exactly, [3] rows
  long long *trace;           // optional (probe): clock64 of CTA 0's first 256 tiles, 6 stamps each (see pdae_chamfer_tc_probe)
  int nbounds;                // > 0: CTA c walks units [bounds[c], bounds[c + 1]) (cost-balanced shares, tcc_partition)
  long long bounds[TCC_MAX_BOUNDS];
};

// literal reference scan of one row (chamfer.cu:42-79), for clouds with non-finite coordinates
__device__ __forceinline__ void tcc_exact_row(const float *__restrict__ R, int nr, float ax, float ay, float az, float &best,
                                              int &bi) {
  best = 0.f;
  bi = 0;
  for (int j = 0; j < nr; ++j) {
    const float d = dist_yxz(__fsub_rn(__ldg(R + 3 * j), ax), __fsub_rn(__ldg(R + 3 * j + 1), ay), __fsub_rn(__ldg(R + 3 * j + 2), az));
    if (j == 0 || d < best) best = d, bi = j;
  }
}

// TN = columns per accumulator buffer (256: two buffers, 128: four).
// F16: the operands are fp16 hi/lo pairs of the coordinates scaled by a power of two into [-1/2, 1/2] and ONE
// kind::f16 MMA (K = 16: ah.bh + al.bh + ah.bl + al.bl + three pieces of |b|^2) makes a tile, instead of two kind::tf32
// MMAs (K = 8 each): half the tensor time and half the issue work per tile for a 4x wider error bound.
template <int TN, bool F16>
__global__ void __launch_bounds__(TCC_THREADS, 1) chamfer_tc_kernel(const TccArgs args) {
  constexpr int NBUF = 512 / TN;
  constexpr int CHUNK = TN * 16;          // bytes of one 16-byte-wide K chunk of a B tile
  constexpr int BTILE = 2 * CHUNK;        // bytes of one B tile image (K = 8 floats)
  constexpr int ACHUNK = TCC_M * 16;      // 2 KB
  constexpr int MAXT = TCC_MAXCOLS / TN;
  constexpr int NIMG = F16 ? 1 : 2;       // operand images of the searched cloud
  constexpr int HALF = TN / 2;            // columns of a tile one epilogue thread reads
  constexpr int STEPS = HALF / TCC_GROUP; // tcgen05.ld.x32 per tile and thread
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t *full = reinterpret_cast<uint64_t *>(smem);         // [NBUF] accumulator written (tcgen05.commit)
  uint64_t *empty = full + NBUF;                               // [NBUF] accumulator read by all 256 epilogue threads
  uint64_t *lfull = empty + NBUF;                              // [2] candidate lists of a row block complete
  uint64_t *lempty = lfull + 2;                                // [2] ... and consumed by the verifier warps
  uint64_t *aready = lempty + 2;                               // [3] row operand image written (first issuer -> second)
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + 128);
  float *scal = reinterpret_cast<float *>(smem + 160);         // [0..2] centre, [3] eps2, [4] fallback flag, [5] bound scale
  float *red = reinterpret_cast<float *>(smem + 256);          // [13 warps][8] reduction scratch
  unsigned char *b1 = smem + 1024;                             // [MAXT][2 chunks][TN/8][8][16 B]
  unsigned char *aimg = b1 + NIMG * MAXT * BTILE;              // [3][2 chunks][16][8][16 B]
  // the searched cloud as given, planar, every 32-column group padded to 36 floats (16-byte aligned rows whose starts
  // fall on different banks); padded columns hold x = +inf, so their distance is +inf
  float *rpx = reinterpret_cast<float *>(aimg + 3 * 2 * ACHUNK);         // [3][TCC_MAXCOLS / 32][36]
  constexpr int RPLANE = TCC_MAXCOLS / TCC_GROUP * TCC_RSTRIDE;
  uint2 *lists = reinterpret_cast<uint2 *>(rpx + 3 * RPLANE);            // [2][CAP][256]
  float *sbest = reinterpret_cast<float *>(lists + 2 * TCC_CAP * TCC_EPI);  // [2][256]
  int *scnt = reinterpret_cast<int *>(sbest + 2 * TCC_EPI);              // [2][256]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // roles: warps [0, VER) verify, [VER, VER + PROD) issue MMAs, the last eight are the epilogue (their TMEM lane quarter is
  // warp % 4, so the block of eight must start at a multiple of four)
  constexpr int W_ISS = TCC_VER_WARPS, W_EPI = TCC_VER_WARPS + TCC_PROD_WARPS;
  static_assert(W_EPI % 4 == 0, "epilogue warps must start at a multiple of four");
  const int etid = tid - W_EPI * 32;  // index among the epilogue threads
  if (tid == 0) {
    for (int i = 0; i < NBUF; ++i) {
      tcc::mbar_init(full + i, 1);
      tcc::mbar_init(empty + i, TCC_EPI);
    }
    for (int i = 0; i < 2; ++i) {
      tcc::mbar_init(lfull + i, TCC_EPI);
      tcc::mbar_init(lempty + i, 32 * TCC_VER_WARPS);
    }
    for (int i = 0; i < 3; ++i) tcc::mbar_init(aready + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tcc::smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcc::tc_fence_before();
  __syncthreads();
  tcc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // programmatic dependent launch: a kernel the caller has queued behind this one as INDEPENDENT of its results (the
  // patchifier of a training step, pdae_fps_group_ex_f32 with PDAE_LAUNCH_OVERLAP_PREVIOUS) may be handed to the block
  // scheduler now -- this grid is fully resident (one CTA per SM, all of shared and tensor memory), so the dependent's CTAs
  // start exactly when and where a CTA of this grid exits instead of after the whole grid has drained.  No-op otherwise.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  constexpr uint32_t IDESC = F16 ? tcc::instr_desc_f16(TCC_M, TN) : tcc::instr_desc_tf32(TCC_M, TN);

  // this CTA's contiguous share of the row blocks
  const int units0 = args.d[0].rbs * args.d[0].nch;
  const int per_cloud = units0 + args.d[1].rbs * args.d[1].nch;
  const long long u0 = args.nbounds ? args.bounds[blockIdx.x] : static_cast<long long>(blockIdx.x) * args.units / gridDim.x;
  const long long u1 = args.nbounds ? args.bounds[blockIdx.x + 1] : static_cast<long long>(blockIdx.x + 1) * args.units / gridDim.x;
  uint32_t k = 0;   // accumulator tiles issued / consumed so far (producer and epilogue count alike)
  uint32_t kb = 0;  // row blocks handed from the epilogue to the verifier so far
  long long u = u0;
  while (u < u1) {
    // ---- a run of row blocks of one (cloud, direction, column chunk): build the chunk's operand image ------------------
    const long long cloud = u / per_cloud;
    const int t0 = static_cast<int>(u - cloud * per_cloud);
    const int dir = t0 >= units0 ? 1 : 0;
    const float *Q = (dir ? args.d[1].q : args.d[0].q), *R = (dir ? args.d[1].r : args.d[0].r);
    float *odist = dir ? args.d[1].dist : args.d[0].dist;
    int *oidx = dir ? args.d[1].idx : args.d[0].idx;
    uint64_t *okeys = dir ? args.d[1].keys : args.d[0].keys;
    const int nq = dir ? args.d[1].nq : args.d[0].nq, nr_all = dir ? args.d[1].nr : args.d[0].nr;
    const int rbs = dir ? args.d[1].rbs : args.d[0].rbs, chunk = dir ? args.d[1].chunk : args.d[0].chunk;
    const int tt = dir ? t0 - units0 : t0;
    const int cc = tt / rbs, rb0 = tt - cc * rbs;
    const int col_off = cc * chunk;                                  // index of the chunk's first column in the cloud
    const int idx_base = col_off + (dir ? args.d[1].idx_off : args.d[0].idx_off);  // ... in the indices of the keys
    const int nr = min(chunk, nr_all - col_off);                     // columns of this chunk
    long long uend = cloud * per_cloud + (dir ? units0 : 0) + static_cast<long long>(cc + 1) * rbs;
    if (uend > u1) uend = u1;
    const int nrb = static_cast<int>(uend - u);  // row blocks rb0 .. rb0 + nrb - 1
    Q += static_cast<size_t>(cloud) * nq * 3;
    R += (static_cast<size_t>(cloud) * nr_all + col_off) * 3;
    odist += cloud * nq;
    oidx += cloud * nq;
    if (okeys) okeys += cloud * nq;
    const int ntiles = (nr + TN - 1) / TN;

    {  // centre = middle of the searched cloud's bounding box; raw copy for the exact evaluation
      float lx = INFINITY, ly = INFINITY, lz = INFINITY, hx = -INFINITY, hy = -INFINITY, hz = -INFINITY;
      constexpr int NIT = (TCC_MAXCOLS + TCC_THREADS - 1) / TCC_THREADS;
      float vx[NIT], vy[NIT], vz[NIT];
#pragma unroll
      for (int it = 0; it < NIT; ++it) {  // all loads in flight before the first use
        const int j = tid + it * TCC_THREADS;
        const int jc = j < nr ? j : nr - 1;
        vx[it] = __ldg(R + 3 * jc), vy[it] = __ldg(R + 3 * jc + 1), vz[it] = __ldg(R + 3 * jc + 2);
      }
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int j = tid + it * TCC_THREADS;
        if (j < ntiles * TN) {
          const int o = (j >> 5) * TCC_RSTRIDE + (j & 31);
          if (j < nr) {
            const float x = vx[it], y = vy[it], z = vz[it];
            rpx[o] = x, rpx[RPLANE + o] = y, rpx[2 * RPLANE + o] = z;
            lx = fminf(lx, x), hx = fmaxf(hx, x);
            ly = fminf(ly, y), hy = fmaxf(hy, y);
            lz = fminf(lz, z), hz = fmaxf(hz, z);
          } else {
            rpx[o] = INFINITY, rpx[RPLANE + o] = 0.f, rpx[2 * RPLANE + o] = 0.f;
          }
        }
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        lx = fminf(lx, __shfl_xor_sync(0xffffffffu, lx, o)), hx = fmaxf(hx, __shfl_xor_sync(0xffffffffu, hx, o));
        ly = fminf(ly, __shfl_xor_sync(0xffffffffu, ly, o)), hy = fmaxf(hy, __shfl_xor_sync(0xffffffffu, hy, o));
        lz = fminf(lz, __shfl_xor_sync(0xffffffffu, lz, o)), hz = fmaxf(hz, __shfl_xor_sync(0xffffffffu, hz, o));
      }
      if (lane == 0) {
        float *w = red + warp * 8;
        w[0] = lx, w[1] = ly, w[2] = lz, w[3] = hx, w[4] = hy, w[5] = hz;
      }
      __syncthreads();
      if (tid < 3) {
        float l = red[tid], h = red[3 + tid];
        for (int w = 1; w < TCC_THREADS / 32; ++w) l = fminf(l, red[w * 8 + tid]), h = fmaxf(h, red[w * 8 + 3 + tid]);
        scal[tid] = 0.5f * l + 0.5f * h;
      }
      __syncthreads();
    }
    const float cx = scal[0], cy = scal[1], cz = scal[2];
    float rmax = 0.f;  // max |b'|^2 and max |a'|^2 over the rows of this run
    float bad = 0.f;   // becomes NaN when any squared norm is NaN or infinite (fmaxf would drop a NaN)
    for (int j = tid; j < nr; j += TCC_THREADS) {
      const int o = (j >> 5) * TCC_RSTRIDE + (j & 31);
      const float x = __fsub_rn(rpx[o], cx), y = __fsub_rn(rpx[RPLANE + o], cy), z = __fsub_rn(rpx[2 * RPLANE + o], cz);
      const float m = __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x)));
      rmax = fmaxf(rmax, m);
      bad += m * 0.f;
    }
    float amax = 0.f;
    {
      const int r_lo = rb0 * TCC_M, r_hi = min(nq, (rb0 + nrb) * TCC_M);
      constexpr int NIT = (TCC_MAXCOLS + TCC_THREADS - 1) / TCC_THREADS;
      for (int base = r_lo; base < r_hi; base += NIT * TCC_THREADS) {
        float vx[NIT], vy[NIT], vz[NIT];
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
          const int i = base + tid + it * TCC_THREADS;
          const int ic = i < r_hi ? i : r_hi - 1;
          vx[it] = __ldg(Q + 3 * ic), vy[it] = __ldg(Q + 3 * ic + 1), vz[it] = __ldg(Q + 3 * ic + 2);
        }
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
          const float x = __fsub_rn(vx[it], cx), y = __fsub_rn(vy[it], cy), z = __fsub_rn(vz[it], cz);
          const float m = __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x)));
          amax = fmaxf(amax, m);
          bad += m * 0.f;
        }
      }
    }
    {
      float s = bad + (cx + cy + cz) * 0.f;  // NaN when anything above was not finite
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        s = s + __shfl_xor_sync(0xffffffffu, s, o);
      }
      if (lane == 0) red[warp * 8] = rmax, red[warp * 8 + 1] = amax, red[warp * 8 + 2] = s;
    }
    __syncthreads();
    if (tid == 0) {
      float rm = 0.f, am = 0.f, s = 0.f;
      for (int w = 0; w < TCC_THREADS / 32; ++w) rm = fmaxf(rm, red[w * 8]), am = fmaxf(am, red[w * 8 + 1]), s += red[w * 8 + 2];
      // F16: the power of two that brings every centred coordinate into [-1/2, 1/2] (so that -2 b fits [-1, 1])
      float sc = 1.f;
      if (F16) {
        int e = 0;
        (void)frexpf(sqrtf(fmaxf(rm, am)), &e);  // largest norm = f * 2^e, f in [1/2, 1)
        sc = (rm + am > 0.f && rm + am < 1.0e30f) ? ldexpf(1.f, -(e + 1)) : 1.f;
      }
      const float eps2 = 2.f * args.eps_rel * (rm + am);
      scal[3] = eps2 * sc * sc;  // the accumulators hold (scaled) distances minus the row's own squared norm
      scal[4] = (s == 0.f && eps2 < 1.0e25f) ? 0.f : 1.f;  // non-finite or huge coordinates: literal scan of every row
      scal[5] = rm + am;
      scal[6] = sc;
    }
    __syncthreads();
    const float sc = scal[6];
    for (int j = tid; j < ntiles * TN; j += TCC_THREADS) {
      const int t = j / TN, r = j - t * TN;
      const int off = t * BTILE + (r >> 3) * 128 + (r & 7) * 16;
      const int o = (j >> 5) * TCC_RSTRIDE + (j & 31);
      const float x = __fsub_rn(rpx[o], cx) * sc, y = __fsub_rn(rpx[RPLANE + o], cy) * sc, z = __fsub_rn(rpx[2 * RPLANE + o], cz) * sc;
      const float m = __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x)));
      if (F16) {
        uint4 c0 = make_uint4(0u, 0u, 0u, 0u), c1 = make_uint4(0u, 0u, 0x7b53u /* 60000: a padded column */, 0u);
        if (j < nr) {
          uint32_t hx, lx, hy, ly, hz, lz, m1, m2, m3, m4;
          tcc::split_h(-2.f * x, hx, lx), tcc::split_h(-2.f * y, hy, ly), tcc::split_h(-2.f * z, hz, lz);
          tcc::split_h(m, m1, m2);
          tcc::split_h(__fsub_rn(__fsub_rn(m, __half2float(__ushort_as_half(m1))), __half2float(__ushort_as_half(m2))), m3, m4);
          c0 = make_uint4(hx | (hy << 16), hz | (hx << 16), hy | (hz << 16), lx | (ly << 16));
          c1 = make_uint4(lz | (lx << 16), ly | (lz << 16), m1 | (m2 << 16), m3);
        }
        *reinterpret_cast<uint4 *>(b1 + off) = c0;
        *reinterpret_cast<uint4 *>(b1 + off + CHUNK) = c1;
      } else {
        float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = make_float4(0.f, 0.f, TCC_BIG, 0.f), l0 = c0, l1 = c0;
        if (j < nr) {
          const float hx = tcc::tf32_rn(x), hy = tcc::tf32_rn(y), hz = tcc::tf32_rn(z);
          const float lx = tcc::tf32_rn(__fsub_rn(x, hx)), ly = tcc::tf32_rn(__fsub_rn(y, hy)), lz = tcc::tf32_rn(__fsub_rn(z, hz));
          const float mh = tcc::tf32_rn(m), ml = tcc::tf32_rn(__fsub_rn(m, mh));
          c0 = make_float4(-2.f * hx, -2.f * hy, -2.f * hz, -2.f * hx);
          c1 = make_float4(-2.f * hy, -2.f * hz, mh, ml);
          l0 = make_float4(-2.f * lx, -2.f * ly, -2.f * lz, -2.f * lx);
          l1 = make_float4(-2.f * ly, -2.f * lz, 0.f, 0.f);
        }
        unsigned char *b2 = b1 + MAXT * BTILE;
        *reinterpret_cast<float4 *>(b1 + off) = c0;
        *reinterpret_cast<float4 *>(b1 + off + CHUNK) = c1;
        *reinterpret_cast<float4 *>(b2 + off) = l0;
        *reinterpret_cast<float4 *>(b2 + off + CHUNK) = l1;
      }
    }
    tcc::fence_proxy_async();  // the operand image was written through the generic proxy
    __syncthreads();
    const float eps2 = scal[3];
    const bool fallback_all = scal[4] != 0.f;

    if (fallback_all) {
      for (int i = rb0 * TCC_M + tid; i < min(nq, (rb0 + nrb) * TCC_M); i += TCC_THREADS) {
        float best;
        int bi;
        tcc_exact_row(R, nr, __ldg(Q + 3 * i), __ldg(Q + 3 * i + 1), __ldg(Q + 3 * i + 2), best, bi);
        if (okeys) {
          atomicMin(reinterpret_cast<unsigned long long *>(okeys + i), pack_key(best, static_cast<uint32_t>(bi + idx_base)));
        } else {
          odist[i] = best;
          oidx[i] = bi;
        }
      }
    } else if (warp >= W_ISS && warp < W_EPI) {
      // ================= issuer warps: row operands + MMA issue =======================================================
      // Tiles alternate between the two issuers (tile k belongs to issuer k & 1).  The first issuer also fetches the rows
      // of block rbl + 1 while block rbl's tiles are issued and writes their operand image (one of three buffers: the
      // MMAs of block rbl - 1 may still be reading theirs) half way through; `aready` hands it to the second issuer.
      const int p = warp - W_ISS;
      float qx[TCC_M / 32], qy[TCC_M / 32], qz[TCC_M / 32];
      auto fetch_rows = [&](int rbl) {
#pragma unroll
        for (int t = 0; t < TCC_M / 32; ++t) {
          int i = (rb0 + rbl) * TCC_M + lane + 32 * t;
          i = i < nq ? i : nq - 1;
          qx[t] = __ldg(Q + 3 * i), qy[t] = __ldg(Q + 3 * i + 1), qz[t] = __ldg(Q + 3 * i + 2);
        }
      };
      auto build_rows = [&](uint32_t kbn) {  // kbn = running number of the row block
        unsigned char *A = aimg + (kbn % 3) * 2 * ACHUNK;
#pragma unroll
        for (int t = 0; t < TCC_M / 32; ++t) {
          const int r = lane + 32 * t;
          const float x = __fsub_rn(qx[t], cx) * sc, y = __fsub_rn(qy[t], cy) * sc, z = __fsub_rn(qz[t], cz) * sc;
          const int off = (r >> 3) * 128 + (r & 7) * 16;
          if (F16) {
            uint32_t hx, lx, hy, ly, hz, lz;
            tcc::split_h(x, hx, lx), tcc::split_h(y, hy, ly), tcc::split_h(z, hz, lz);
            constexpr uint32_t ONE = 0x3c00u;
            *reinterpret_cast<uint4 *>(A + off) = make_uint4(hx | (hy << 16), hz | (lx << 16), ly | (lz << 16), hx | (hy << 16));
            *reinterpret_cast<uint4 *>(A + off + ACHUNK) = make_uint4(hz | (lx << 16), ly | (lz << 16), ONE | (ONE << 16), ONE);
          } else {
            const float hx = tcc::tf32_rn(x), hy = tcc::tf32_rn(y), hz = tcc::tf32_rn(z);
            const float lx = tcc::tf32_rn(__fsub_rn(x, hx)), ly = tcc::tf32_rn(__fsub_rn(y, hy)), lz = tcc::tf32_rn(__fsub_rn(z, hz));
            *reinterpret_cast<float4 *>(A + off) = make_float4(hx, hy, hz, lx);
            *reinterpret_cast<float4 *>(A + off + ACHUNK) = make_float4(ly, lz, 1.f, 1.f);
          }
        }
        tcc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) tcc::mbar_arrive(aready + kbn % 3);
      };
      if (p == 0) {
        fetch_rows(0);
        build_rows(kb);
      }
      for (int rbl = 0; rbl < nrb; ++rbl, ++kb) {
        const bool next = rbl + 1 < nrb;
        if (p == 0) {
          if (next) fetch_rows(rbl + 1);
        } else if (lane == 0) {
          tcc::mbar_wait(aready + kb % 3, (kb / 3) & 1u);
        }
        const uint64_t ad = tcc::smem_desc(tcc::smem_u32(aimg + (kb % 3) * 2 * ACHUNK), ACHUNK, 128);
        for (int t = 0; t < ntiles; ++t, ++k) {
          if (lane == 0 && static_cast<int>(k % TCC_PROD_WARPS) == p) {
            const uint32_t buf = k % NBUF, use = k / NBUF;
            tcc::mbar_wait(empty + buf, (use & 1u) ^ 1u);  // the epilogue has read this accumulator's previous tile
            tcc::tc_fence_after();
            if (args.trace && blockIdx.x == 0 && k < 256) args.trace[k * 6 + 0] = clock64();
            const uint64_t bd1 = tcc::smem_desc(tcc::smem_u32(b1 + t * BTILE), CHUNK, 128);
            if (F16) {
              tcc::mma_f16(tmem + buf * TN, ad, bd1, IDESC, 0u);
            } else {
              const uint64_t bd2 = tcc::smem_desc(tcc::smem_u32(b1 + (MAXT + t) * BTILE), CHUNK, 128);
              tcc::mma_tf32(tmem + buf * TN, ad, bd1, IDESC, 0u);
              tcc::mma_tf32(tmem + buf * TN, ad, bd2, IDESC, 1u);
            }
            tcc::mma_commit(full + buf);
            if (args.trace && blockIdx.x == 0 && k < 256) args.trace[k * 6 + 1] = clock64();
          }
          __syncwarp();
          if (p == 0 && next && t == (ntiles - 1) / 2) build_rows(kb + 1);
        }
      }
    } else if (warp >= W_EPI) {
      // ================= epilogue warps: group minima and candidate lists ===========================================
      const int quarter = warp & 3, half = (warp - W_EPI) >> 2;
      const uint32_t tbase = tmem + (static_cast<uint32_t>(quarter * 32) << 16) + half * HALF;
      for (int rbl = 0; rbl < nrb; ++rbl, ++kb) {
        const uint32_t slot = kb & 1u;
        tcc::mbar_wait(lempty + slot, ((kb >> 1) & 1u) ^ 1u);  // the verifier is done with this slot's previous row block
        uint2 *mylist = lists + slot * (TCC_CAP * TCC_EPI) + etid;
        float best = INFINITY, thr = INFINITY, low = INFINITY;
        int cnt = 0;
        for (int t = 0; t < ntiles; ++t, ++k) {
          const uint32_t buf = k % NBUF, use = k / NBUF;
          const bool tr = args.trace && blockIdx.x == 0 && k < 256 && etid == 0;
          if (tr) args.trace[k * 6 + 2] = clock64();
          tcc::mbar_wait(full + buf, use & 1u);
          tcc::tc_fence_after();
          if (tr) args.trace[k * 6 + 3] = clock64();
          // the whole 128-column share in registers at once: the accumulator is released before any of it is processed
          // (loading it in two halves keeps the kernel under 128 registers but releases later: 131 vs 122 us)
          float v[STEPS][32];
#pragma unroll
          for (int s = 0; s < STEPS; ++s) PDAE_TMEM_LD32(tbase + buf * TN + s * TCC_GROUP, v[s]);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          tcc::tc_fence_before();
          tcc::mbar_arrive(empty + buf);
          if (args.trace && blockIdx.x == 0 && k < 256 && lane == 0)  // the last warp's release is the one that counts
            atomicMax(reinterpret_cast<unsigned long long *>(args.trace + k * 6 + 4), static_cast<unsigned long long>(clock64()));
#pragma unroll
          for (int s = 0; s < STEPS; ++s) {
            float m0 = min3(v[s][0], v[s][1], v[s][2]), m1 = min3(v[s][3], v[s][4], v[s][5]);
            float m2 = min3(v[s][6], v[s][7], v[s][8]), m3 = min3(v[s][9], v[s][10], v[s][11]);
            m0 = min3(m0, v[s][12], v[s][13]), m1 = min3(m1, v[s][14], v[s][15]);
            m2 = min3(m2, v[s][16], v[s][17]), m3 = min3(m3, v[s][18], v[s][19]);
            m0 = min3(m0, v[s][20], v[s][21]), m1 = min3(m1, v[s][22], v[s][23]);
            m2 = min3(m2, v[s][24], v[s][25]), m3 = min3(m3, v[s][26], v[s][27]);
            m0 = min3(m0, v[s][28], v[s][29]), m1 = min3(m1, v[s][30], v[s][31]);
            const float g = fminf(min3(m0, m1, m2), m3);
            const int gid = t * (TN / TCC_GROUP) + half * (HALF / TCC_GROUP) + s;
            cnt = g < low ? 0 : cnt;  // everything kept so far is more than 2 eps above this group
            const bool take = g <= thr;
            const int pos = cnt < TCC_CAP ? cnt : TCC_CAP - 1;  // an overflowing list is flagged by cnt > CAP
            if (take) mylist[pos * TCC_EPI] = make_uint2(__float_as_uint(g), static_cast<uint32_t>(gid));
            cnt += take ? 1 : 0;
            best = fminf(best, g);
            thr = best + eps2;
            low = best - eps2;
          }
          if (tr) args.trace[k * 6 + 5] = clock64();
        }
        sbest[slot * TCC_EPI + etid] = best;
        scnt[slot * TCC_EPI + etid] = cnt;
        tcc::mbar_arrive(lfull + slot);
      }
    } else {
      // ================= verifier warps: exact evaluation of the surviving groups ====================================
      // a warp owns 32 rows of the block: every lane filters its own row's lists, then the warp evaluates the surviving
      // (row, group) items one after the other with a lane per column
      const int vw = warp;
      // the four 32-row quarters of a block rotate over the verifier warps: a warp owns one lane per row
      for (int rbl = 0; rbl < nrb; ++rbl, ++kb) {
        const uint32_t slot = kb & 1u;
        const int first = (vw + TCC_VER_WARPS - static_cast<int>(kb % TCC_VER_WARPS)) % TCC_VER_WARPS;  // (sub + kb) % VER == vw
        float nx, ny, nz;  // the next quarter's rows, fetched one quarter ahead
        {
          const int i = (rb0 + rbl) * TCC_M + first * 32 + lane;
          const int ic = i < nq ? i : nq - 1;
          nx = __ldg(Q + 3 * ic), ny = __ldg(Q + 3 * ic + 1), nz = __ldg(Q + 3 * ic + 2);
        }
        const bool vtr = args.trace && blockIdx.x == 0 && kb < 64 && vw == 0 && lane == 0;
        if (vtr) args.trace[1536 + kb * 4 + 0] = clock64();
        tcc::mbar_wait(lfull + slot, (kb >> 1) & 1u);
        if (vtr) args.trace[1536 + kb * 4 + 1] = clock64();
#pragma unroll 1
       for (int sub = first; sub < TCC_M / 32; sub += TCC_VER_WARPS) {
        const bool last = sub + TCC_VER_WARPS >= TCC_M / 32;
        const int row = sub * 32 + lane;
        const int i = (rb0 + rbl) * TCC_M + row;
        const bool live = i < nq;
        const float ax = nx, ay = ny, az = nz;
        if (!last) {
          const int i2 = i + 32 * TCC_VER_WARPS;
          const int ic = i2 < nq ? i2 : nq - 1;
          nx = __ldg(Q + 3 * ic), ny = __ldg(Q + 3 * ic + 1), nz = __ldg(Q + 3 * ic + 2);
        }
        const uint2 *L = lists + slot * (TCC_CAP * TCC_EPI);
        const float limit = fminf(sbest[slot * TCC_EPI + row], sbest[slot * TCC_EPI + row + 128]) + eps2;
        int g0 = 0, g1 = 0, g2 = 0, g3 = 0, nmine = 0;  // this row's surviving groups
        float a0 = 0.f;                                  // approximate minimum of the first (probe only)
        bool full_scan = false;
        {
          // the first four entries of both halves are fetched unconditionally (independent loads, no loop-carried
          // latency); longer lists are rare and walked by a warp-uniform loop
          const int cnt0 = scnt[slot * TCC_EPI + row], cnt1 = scnt[slot * TCC_EPI + row + 128];
          full_scan |= cnt0 > TCC_CAP || cnt1 > TCC_CAP;
          const int nl0 = cnt0 < TCC_CAP ? cnt0 : TCC_CAP, nl1 = cnt1 < TCC_CAP ? cnt1 : TCC_CAP;
          uint2 ent[2][4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            ent[0][e] = L[e * TCC_EPI + row];
            ent[1][e] = L[e * TCC_EPI + row + 128];
          }
          auto consider = [&](uint2 en, bool valid) {
            const bool ok = valid && __uint_as_float(en.x) <= limit;
            const int g = static_cast<int>(en.y);
            a0 = (ok && nmine == 0) ? __uint_as_float(en.x) : a0;
            g0 = (ok && nmine == 0) ? g : g0;
            g1 = (ok && nmine == 1) ? g : g1;
            g2 = (ok && nmine == 2) ? g : g2;
            g3 = (ok && nmine == 3) ? g : g3;
            nmine += ok ? 1 : 0;
          };
#pragma unroll
          for (int e = 0; e < 4; ++e) consider(ent[0][e], e < nl0);
#pragma unroll
          for (int e = 0; e < 4; ++e) consider(ent[1][e], e < nl1);
          const int nlmax = __reduce_max_sync(0xffffffffu, nl0 > nl1 ? nl0 : nl1);
          for (int e = 4; e < nlmax; ++e) {
            consider(L[e * TCC_EPI + row], e < nl0);
            consider(L[e * TCC_EPI + row + 128], e < nl1);
          }
        }
        full_scan |= nmine > 4 || nmine == 0;
        __syncwarp();
        if (vtr && sub == first) args.trace[1792 + kb * 4 + 0] = clock64();
        if (last) tcc::mbar_arrive(lempty + slot);  // this warp's last lists are in registers
        uint64_t key = ~0ull;
        auto probe_err = [&](float approx, uint32_t exact_bits) {
          const float x = __fsub_rn(ax, cx), y = __fsub_rn(ay, cy), z = __fsub_rn(az, cz);
          const float na = __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x)));
          const float err = fabsf(approx / (scal[6] * scal[6]) - (__uint_as_float(exact_bits) - na)) / scal[5];
          atomicMax(args.stats, static_cast<unsigned long long>(__float_as_uint(err)));
          atomicAdd(args.stats + 2, 1ull);
        };
        {  // nearly every row has exactly one surviving group: every lane walks its own group's 32 columns
          const float *px = rpx + g0 * TCC_RSTRIDE;
          float d[TCC_GROUP];
#pragma unroll
          for (int c4 = 0; c4 < TCC_GROUP / 4; ++c4) {
            const float4 X = *reinterpret_cast<const float4 *>(px + 4 * c4);
            const float4 Y = *reinterpret_cast<const float4 *>(px + RPLANE + 4 * c4);
            const float4 Z = *reinterpret_cast<const float4 *>(px + 2 * RPLANE + 4 * c4);
            d[4 * c4] = dist_yxz(__fsub_rn(X.x, ax), __fsub_rn(Y.x, ay), __fsub_rn(Z.x, az));
            d[4 * c4 + 1] = dist_yxz(__fsub_rn(X.y, ax), __fsub_rn(Y.y, ay), __fsub_rn(Z.y, az));
            d[4 * c4 + 2] = dist_yxz(__fsub_rn(X.z, ax), __fsub_rn(Y.z, ay), __fsub_rn(Z.z, az));
            d[4 * c4 + 3] = dist_yxz(__fsub_rn(X.w, ax), __fsub_rn(Y.w, ay), __fsub_rn(Z.w, az));
          }
          float m0 = min3(d[0], d[1], d[2]), m1 = min3(d[3], d[4], d[5]), m2 = min3(d[6], d[7], d[8]), m3 = min3(d[9], d[10], d[11]);
          m0 = min3(m0, d[12], d[13]), m1 = min3(m1, d[14], d[15]), m2 = min3(m2, d[16], d[17]), m3 = min3(m3, d[18], d[19]);
          m0 = min3(m0, d[20], d[21]), m1 = min3(m1, d[22], d[23]), m2 = min3(m2, d[24], d[25]), m3 = min3(m3, d[26], d[27]);
          m0 = min3(m0, d[28], d[29]), m1 = min3(m1, d[30], d[31]);
          const float dm = fminf(min3(m0, m1, m2), m3);
          int f0 = TCC_GROUP, f1 = TCC_GROUP, f2 = TCC_GROUP, f3 = TCC_GROUP;  // four short chains instead of one of 32
#pragma unroll
          for (int c = TCC_GROUP / 4 - 1; c >= 0; --c) {
            f0 = (d[c] == dm) ? c : f0;
            f1 = (d[c + 8] == dm) ? c + 8 : f1;
            f2 = (d[c + 16] == dm) ? c + 16 : f2;
            f3 = (d[c + 24] == dm) ? c + 24 : f3;
          }
          int found = min(min(f0, f1), min(f2, f3));  // the lowest column attaining it
          found = found < TCC_GROUP ? found : 0;
          if (nmine >= 1) {
            key = pack_key(dm, static_cast<uint32_t>(g0 * TCC_GROUP + found));
            if (args.stats && live) probe_err(a0, __float_as_uint(dm));
          }
        }
        auto eval = [&](int r, int gsel) {  // lane r's further item: every lane takes one column of the group
          const int g = __shfl_sync(0xffffffffu, gsel, r);
          const float bx = __shfl_sync(0xffffffffu, ax, r), by = __shfl_sync(0xffffffffu, ay, r), bz = __shfl_sync(0xffffffffu, az, r);
          const int o = g * TCC_RSTRIDE + lane;
          const float d = dist_yxz(__fsub_rn(rpx[o], bx), __fsub_rn(rpx[RPLANE + o], by), __fsub_rn(rpx[2 * RPLANE + o], bz));
          const uint32_t kd = __float_as_uint(d);
          const uint32_t m = __reduce_min_sync(0xffffffffu, kd);
          const uint32_t who = __ballot_sync(0xffffffffu, kd == m);
          const uint64_t cand = (static_cast<uint64_t>(m) << 32) | static_cast<uint32_t>(g * TCC_GROUP + (__ffs(who) - 1));
          if (lane == r) key = key < cand ? key : cand;
        };
        if (vtr && sub == first) args.trace[1792 + kb * 4 + 1] = clock64();
        for (uint32_t mask = __ballot_sync(0xffffffffu, nmine >= 2); mask; mask &= mask - 1) eval(__ffs(mask) - 1, g1);
        for (uint32_t mask = __ballot_sync(0xffffffffu, nmine >= 3); mask; mask &= mask - 1) eval(__ffs(mask) - 1, g2);
        for (uint32_t mask = __ballot_sync(0xffffffffu, nmine >= 4); mask; mask &= mask - 1) eval(__ffs(mask) - 1, g3);
        if (vtr && sub == first) args.trace[1792 + kb * 4 + 2] = clock64();
        // overflowed list, too many survivors, none, or a non-finite winner: the whole row, the lanes striding the columns
        full_scan |= !(__uint_as_float(static_cast<uint32_t>(key >> 32)) < INFINITY);
        for (uint32_t mask = __ballot_sync(0xffffffffu, full_scan && live); mask; mask &= mask - 1) {
          const int r = __ffs(mask) - 1;
          const float bx = __shfl_sync(0xffffffffu, ax, r), by = __shfl_sync(0xffffffffu, ay, r), bz = __shfl_sync(0xffffffffu, az, r);
          uint64_t mine = ~0ull;
          for (int j = lane; j < nr; j += 32) {
            const int o = (j >> 5) * TCC_RSTRIDE + lane;
            const float d = dist_yxz(__fsub_rn(rpx[o], bx), __fsub_rn(rpx[RPLANE + o], by), __fsub_rn(rpx[2 * RPLANE + o], bz));
            const uint64_t c = pack_key(d, static_cast<uint32_t>(j));
            mine = mine < c ? mine : c;
          }
#pragma unroll
          for (int o = 16; o; o >>= 1) {
            const uint64_t other = __shfl_xor_sync(0xffffffffu, mine, o);
            mine = mine < other ? mine : other;
          }
          if (lane == r) {
            key = mine;
            if (args.stats) atomicAdd(args.stats + 1, 1ull);
          }
        }
        if (live) {
          if (okeys) {  // one of several column chunks: the (distance, index) order of the key is the reference's tie rule
            atomicMin(reinterpret_cast<unsigned long long *>(okeys + i), key + static_cast<uint32_t>(idx_base));
          } else {
            odist[i] = __uint_as_float(static_cast<uint32_t>(key >> 32));
            oidx[i] = static_cast<int>(static_cast<uint32_t>(key));
          }
          if (args.stats) atomicAdd(args.stats + 3, 1ull);
        }
        if (vtr) args.trace[1536 + kb * 4 + 2 + (sub == first ? 0 : 1)] = clock64();
       }
      }
    }
    u = uend;
    k = __shfl_sync(0xffffffffu, k, 0);
    __syncthreads();  // every MMA of this run has been consumed and verified: the operand image may be rebuilt
  }
  tcc::tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

template <int TN, bool F16>
static size_t tcc_smem_bytes() {
  return 1024 + static_cast<size_t>(F16 ? 1 : 2) * (TCC_MAXCOLS / TN) * (2 * TN * 16) + 3 * 2 * TCC_M * 16 + 3 * (TCC_MAXCOLS / TCC_GROUP) * TCC_RSTRIDE * 4 +
         2 * TCC_CAP * TCC_EPI * 8 + 2 * TCC_EPI * (4 + 4);
}

// tuning / test hook state: mode 0 = off, 1 / 2 = tf32 filter with 128- / 256-column accumulators, 3 = fp16 filter (default)
static int g_tcc_mode = -1;
static float g_tcc_eps_rel = 0.f;
static int tcc_mode() {
  if (g_tcc_mode < 0) {
    const char *e = getenv("PDAE_CHAMFER_TC");
    g_tcc_mode = e ? atoi(e) : 3;
    const char *x = getenv("PDAE_CHAMFER_TC_EPS");
    g_tcc_eps_rel = x ? static_cast<float>(atof(x)) : 0.f;  // 0 = the mode's default bound
  }
  return g_tcc_mode;
}

// column chunks of a searched cloud of nr points: as few as fit the resident image, of (nearly) equal size, a multiple of 256
static void tcc_chunks(int nr, int &nch, int &chunk) {
  nch = (nr + TCC_MAXCOLS - 1) / TCC_MAXCOLS;
  chunk = ((nr + nch - 1) / nch + 255) / 256 * 256;
  nch = (nr + chunk - 1) / chunk;
}

// workspace of chamfer_tc_forward: merged (distance, index) keys of the rows whose searched cloud is cut into chunks
static size_t tcc_workspace_bytes(int b, int n, int m) {
  return (static_cast<size_t>(m > TCC_MAXCOLS ? n : 0) + static_cast<size_t>(n > TCC_MAXCOLS ? m : 0)) * b * sizeof(uint64_t);
}

// true when chamfer_tc_forward serves this shape: both clouds at least 512 points; clouds above 2048 points are searched
// in column chunks whose results merge through keys in the caller's workspace
bool chamfer_tc_applies(int b, int n, int m, size_t workspace_bytes) {
  if (tcc_mode() <= 0) return false;
  const int lo = n < m ? n : m;
  return b > 0 && lo >= 512 && workspace_bytes >= tcc_workspace_bytes(b, n, m);
}

__global__ void __launch_bounds__(256) tcc_fill_keys_kernel(uint64_t *__restrict__ keys, long long count) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < count) keys[i] = ~0ull;
}
__global__ void __launch_bounds__(256) tcc_unpack_keys_kernel(const uint64_t *__restrict__ keys, long long count,
                                                              float *__restrict__ dist, int *__restrict__ idx) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint64_t k = keys[i];
  dist[i] = __uint_as_float(static_cast<uint32_t>(k >> 32));
  idx[i] = static_cast<int>(static_cast<uint32_t>(k));
}

// Cost-balanced shares.  A CTA pays for every row block (its accumulator tiles) AND for every operand image it has to
// build (pipeline drain + bounding box + 64 KB image: 13.8 k cycles at 2048 columns = 2.5 row blocks of eight tiles), and
// equal block counts give the CTAs whose range touches three (cloud, direction, chunk) runs 7 % more work than those that
// touch two.  The shares are the contiguous partition with the smallest maximum cost (binary search on the cost, greedy
// fill -- every CTA pays a build for its first run wherever it starts, so the longest feasible prefix is optimal);
// walked run by run, so the host work is O(runs) per probe.  PDAE_TCC_BUILD_COST = cost of a 2048-column build in
// tiles (default 20; 0 = equal block counts).
static double tcc_build_cost() {
  static double c = -1.0;
  if (c < 0.0) {
    const char *e = getenv("PDAE_TCC_BUILD_COST");
    c = e && *e ? atof(e) : 20.0;
    if (c < 0.0) c = 0.0;
  }
  return c;
}

template <int TN>
static void tcc_partition(TccArgs &a, int b, int grid) {
  a.nbounds = 0;
  const double bc = tcc_build_cost();
  if (bc <= 0.0 || grid + 1 > TCC_MAX_BOUNDS || grid < 2 || a.units <= grid) return;
  // the runs of one cloud: (row blocks, tiles per block, build cost), in unit order
  struct Run { long long rbs; double w, build; };
  Run runs[2 * 64];
  int nruns = 0;
  for (int d = 0; d < 2; ++d) {
    if (a.d[d].nch > 64) return;
    for (int c = 0; c < a.d[d].nch; ++c) {
      const int cols = a.d[d].nr - c * a.d[d].chunk < a.d[d].chunk ? a.d[d].nr - c * a.d[d].chunk : a.d[d].chunk;
      // a row block of a short last chunk is not cheaper in proportion to its tiles (row operands, candidate lists and the
      // verifier's pass per 128 rows do not shrink: one rank's share of the sharded 100 000-point forward lost 15 % with
      // per-tile weights), so every block of a direction weighs its nominal chunk
      (void)cols;
      runs[nruns++] = Run{a.d[d].rbs, static_cast<double>((a.d[d].chunk + TN - 1) / TN),
                          bc * (0.4 + 0.6 * a.d[d].chunk / TCC_MAXCOLS)};
    }
  }
  // the shares of the last shape are kept (a training loop repeats one shape); several host threads may launch at once
  static std::mutex mu;
  static TccArgs cached;
  static int cached_b = -1, cached_grid = -1, cached_tn = -1;
  static double cached_bc = -1.0;
  auto same = [&](const TccDir &x, const TccDir &y) { return x.nq == y.nq && x.nr == y.nr && x.rbs == y.rbs && x.nch == y.nch && x.chunk == y.chunk; };
  std::lock_guard<std::mutex> lock(mu);
  if (cached_b == b && cached_grid == grid && cached_tn == TN && cached_bc == bc && same(cached.d[0], a.d[0]) &&
      same(cached.d[1], a.d[1])) {
    a.nbounds = cached.nbounds;
    for (int i = 0; i <= grid; ++i) a.bounds[i] = cached.bounds[i];
    return;
  }
  double per_cloud_cost = 0.0;
  for (int r = 0; r < nruns; ++r) per_cloud_cost += runs[r].rbs * runs[r].w + runs[r].build;
  // fill CTAs up to cost T in unit order; returns the number of CTAs used (their first units go to out[], when given)
  auto fill = [&](double T, long long *out) -> long long {
    long long ctas = 1, u = 0;
    double acc = 0.0;
    if (out) out[0] = 0;
    for (int cl = 0; cl < b; ++cl) {
      for (int r = 0; r < nruns; ++r) {
        long long left = runs[r].rbs;
        bool paid = false;  // has the current CTA built this run's image?
        while (left > 0) {
          const double need = paid ? 0.0 : runs[r].build;
          long long fit = static_cast<long long>((T - acc - need) / runs[r].w + 1e-9);
          if (fit <= 0) {
            if (acc > 0.0) {  // close this CTA, the next one goes on with the run (and builds its image again)
              if (out && ctas <= grid) out[ctas] = u;
              ++ctas, acc = 0.0, paid = false;
              continue;
            }
            fit = 1;  // an empty CTA always takes at least one block
          }
          const long long take = fit < left ? fit : left;
          acc += need + take * runs[r].w;
          paid = true, left -= take, u += take;
        }
      }
    }
    if (out && ctas <= grid) out[ctas] = u;
    return ctas;
  };
  double lo = per_cloud_cost * b / grid * 0.5, hi = per_cloud_cost * b / grid * 2.0 + per_cloud_cost;
  for (int it = 0; it < 48; ++it) {
    const double mid = 0.5 * (lo + hi);
    if (fill(mid, nullptr) <= grid) hi = mid; else lo = mid;
  }
  long long bounds[TCC_MAX_BOUNDS];
  long long used = fill(hi, bounds);
  // costs are discrete, so the cheapest feasible limit may need fewer CTAs than there are: halve the longest shares until
  // every CTA has one (a half never costs more than the whole)
  while (used >= 1 && used < grid) {
    int big = 0;
    for (int i = 1; i < used; ++i)
      if (bounds[i + 1] - bounds[i] > bounds[big + 1] - bounds[big]) big = i;
    if (bounds[big + 1] - bounds[big] < 2) break;
    for (long long i = used; i > big; --i) bounds[i + 1] = bounds[i];
    bounds[big + 1] = bounds[big] + (bounds[big + 2] - bounds[big]) / 2;
    ++used;
  }
  cached = a, cached_b = b, cached_grid = grid, cached_tn = TN, cached_bc = bc;
  cached.nbounds = 0;
  // every CTA must own at least one unit (the kernel's roles assume a non-empty share): otherwise equal block counts
  bool ok = used == grid && bounds[grid] == a.units;
  for (int i = 0; ok && i < grid; ++i) ok = bounds[i + 1] > bounds[i];
  if (!ok) return;
  for (int i = 0; i <= grid; ++i) a.bounds[i] = cached.bounds[i] = bounds[i];
  a.nbounds = cached.nbounds = grid + 1;
}

template <int TN, bool F16>
static int tcc_launch(TccArgs &a, int b, cudaStream_t st) {
  const size_t smem = tcc_smem_bytes<TN, F16>();
  // per launch, not once per process: the attribute belongs to the current device's instance of the kernel (one process may
  // drive several GPUs, e.g. the reference's nn.DataParallel threads)
  PDAE_CUDA_TRY(cudaFuncSetAttribute(chamfer_tc_kernel<TN, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    PDAE_CUDA_TRY(cudaGetDevice(&dev));
    PDAE_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const long long grid = a.units < sms ? a.units : sms;
  tcc_partition<TN>(a, b, static_cast<int>(grid));
  chamfer_tc_kernel<TN, F16><<<static_cast<unsigned>(grid), TCC_THREADS, smem, st>>>(a);
  PDAE_RETURN_IF_LAUNCH_FAILED();
  return 0;
}

// keys1 != nullptr (reference-set sharding): the rows of xyz1 leave as packed keys with ref_offset added to the indices
// (dist1 / idx1 unused), xyz2 being one rank's slice of the searched cloud
static int tcc_forward(const float *xyz1, const float *xyz2, int b, int n, int m, float *dist1, float *dist2, int *idx1,
                       int *idx2, uint64_t *keys1, int ref_offset, void *workspace, size_t workspace_bytes, cudaStream_t st,
                       unsigned long long *stats, long long *trace) {
  TccArgs a;
  a.nbounds = 0;
  a.stats = stats;
  a.trace = trace;
  a.d[0] = TccDir{xyz1, xyz2, dist1, idx1, nullptr, n, m, (n + TCC_M - 1) / TCC_M, 1, TCC_MAXCOLS, ref_offset};
  a.d[1] = TccDir{xyz2, xyz1, dist2, idx2, nullptr, m, n, (m + TCC_M - 1) / TCC_M, 1, TCC_MAXCOLS, 0};
  uint64_t *ws = static_cast<uint64_t *>(workspace);
  long long nkeys = 0;
  for (int d = 0; d < 2; ++d) {
    tcc_chunks(a.d[d].nr, a.d[d].nch, a.d[d].chunk);
    if (d == 0 && keys1) continue;
    if (a.d[d].nch > 1) {
      a.d[d].keys = ws + nkeys;
      nkeys += static_cast<long long>(b) * a.d[d].nq;
    }
  }
  if (static_cast<size_t>(nkeys) * sizeof(uint64_t) > workspace_bytes || (nkeys && !workspace)) return PDAE_E_WORKSPACE;
  const long long per_cloud = static_cast<long long>(a.d[0].rbs) * a.d[0].nch + static_cast<long long>(a.d[1].rbs) * a.d[1].nch;
  if (per_cloud > 0x7fffffffLL) return PDAE_E_UNSUPPORTED;
  a.units = static_cast<long long>(b) * per_cloud;
  if (nkeys) {
    tcc_fill_keys_kernel<<<static_cast<unsigned>((nkeys + 255) / 256), 256, 0, st>>>(ws, nkeys);
    PDAE_RETURN_IF_LAUNCH_FAILED();
  }
  if (keys1) {
    const long long cnt = static_cast<long long>(b) * n;
    tcc_fill_keys_kernel<<<static_cast<unsigned>((cnt + 255) / 256), 256, 0, st>>>(keys1, cnt);
    PDAE_RETURN_IF_LAUNCH_FAILED();
    a.d[0].keys = keys1;
  }
  const int mode = tcc_mode();
  a.eps_rel = g_tcc_eps_rel > 0.f ? g_tcc_eps_rel : (mode == 3 ? TCC_EPS_F16 : TCC_EPS_TF32);
  int rc;
  if (mode == 1) rc = tcc_launch<128, false>(a, b, st);
  else if (mode == 3) rc = tcc_launch<256, true>(a, b, st);
  else rc = tcc_launch<256, false>(a, b, st);
  if (rc) return rc;
  for (int d = 0; d < 2; ++d) {
    if (!a.d[d].keys || (d == 0 && keys1)) continue;
    const long long cnt = static_cast<long long>(b) * a.d[d].nq;
    tcc_unpack_keys_kernel<<<static_cast<unsigned>((cnt + 255) / 256), 256, 0, st>>>(a.d[d].keys, cnt, a.d[d].dist, a.d[d].idx);
    PDAE_RETURN_IF_LAUNCH_FAILED();
  }
  return 0;
}

int chamfer_tc_forward(const float *xyz1, const float *xyz2, int b, int n, int m, float *dist1, float *dist2, int *idx1,
                       int *idx2, void *workspace, size_t workspace_bytes, cudaStream_t st, unsigned long long *stats,
                       long long *trace) {
  return tcc_forward(xyz1, xyz2, b, n, m, dist1, dist2, idx1, idx2, nullptr, 0, workspace, workspace_bytes, st, stats, trace);
}

// one rank's share of a reference-set-sharded forward (pdae_chamfer_sharded_f32): true when the tensor-core path serves it
// (workspace: b * m_local keys, the entry point's own requirement, covers the merged keys of the slice's rows)
bool chamfer_tc_sharded_applies(int b, int n, int m_local) {
  if (tcc_mode() <= 0) return false;
  return b > 0 && n >= 512 && m_local >= 512;
}
int chamfer_tc_sharded(const float *xyz1, const float *xyz2_local, int b, int n, int m_local, int ref_offset, uint64_t *keys1,
                       float *dist2_local, int *idx2_local, void *workspace, size_t workspace_bytes, cudaStream_t st) {
  return tcc_forward(xyz1, xyz2_local, b, n, m_local, nullptr, dist2_local, nullptr, idx2_local, keys1, ref_offset, workspace,
                     workspace_bytes, st, nullptr, nullptr);
}

}  // namespace pdae

// tuning / test hook: mode 0 = FP32-pipe kernels only, 1 / 2 = tf32 filter with 128- / 256-column accumulators, 3 = fp16
// filter (default); eps_rel > 0 sets the filter's error bound relative to max|a'|^2 + max|b'|^2, < 0 restores the mode's
// default (2^-17 tf32, 2^-16 fp16).  Negative mode only queries.
// Returns the previous mode.
extern "C" int pdae_tune_chamfer_tc(int mode, float eps_rel) {
  const int old = pdae::tcc_mode();
  if (mode >= 0) pdae::g_tcc_mode = mode;
  if (eps_rel > 0.f) pdae::g_tcc_eps_rel = eps_rel;
  if (eps_rel < 0.f) pdae::g_tcc_eps_rel = 0.f;  // back to the mode's default
  return old;
}

// host-only diagnostic (no device work): the cost-balanced shares tcc_partition gives `grid` persistent CTAs for a
// forward of b clouds of n against m points (256-column tiles).  bounds_out[0 .. grid] receives the first unit of every
// CTA (units = 128-row blocks in (cloud, direction, chunk, block) order); returns grid + 1, or 0 when the launch would use
// equal block counts (PDAE_TCC_BUILD_COST=0, too few units, or more CTAs than the parameter block holds), < 0 on bad arguments.
extern "C" int pdae_chamfer_tc_shares(int b, int n, int m, int grid, long long *bounds_out, long long *units_out) {
  if (b <= 0 || n <= 0 || m <= 0 || grid <= 0 || !bounds_out) return PDAE_E_INVALID;
  pdae::TccArgs a;
  a.nbounds = 0;
  a.stats = nullptr;
  a.trace = nullptr;
  a.d[0] = pdae::TccDir{nullptr, nullptr, nullptr, nullptr, nullptr, n, m, (n + pdae::TCC_M - 1) / pdae::TCC_M, 1, pdae::TCC_MAXCOLS, 0};
  a.d[1] = pdae::TccDir{nullptr, nullptr, nullptr, nullptr, nullptr, m, n, (m + pdae::TCC_M - 1) / pdae::TCC_M, 1, pdae::TCC_MAXCOLS, 0};
  for (int d = 0; d < 2; ++d) pdae::tcc_chunks(a.d[d].nr, a.d[d].nch, a.d[d].chunk);
  a.units = static_cast<long long>(b) * (static_cast<long long>(a.d[0].rbs) * a.d[0].nch + static_cast<long long>(a.d[1].rbs) * a.d[1].nch);
  if (units_out) *units_out = a.units;
  if (grid > a.units) grid = static_cast<int>(a.units);
  pdae::tcc_partition<256>(a, b, grid);
  for (int i = 0; i < a.nbounds; ++i) bounds_out[i] = a.bounds[i];
  return a.nbounds;
}

// probe: the tensor-core forward regardless of the mode switch, plus filter statistics (4 x uint64, zeroed by the caller):
// [0] float bits of the largest observed |approximate - exact| group minimum relative to max|a'|^2 + max|b'|^2,
// [1] rows decided by the literal scan, [2] 32-column groups evaluated exactly, [3] rows written.
// trace (optional, 256 x 6 int64): clock64 of CTA 0's first 256 accumulator tiles -- producer: accumulator free, MMAs
// committed; epilogue thread 0: starts waiting, accumulator ready, accumulator released, tile processed.
extern "C" int pdae_chamfer_tc_probe(const float *xyz1, const float *xyz2, int b, int n, int m, float *dist1, float *dist2,
                                     int *idx1, int *idx2, unsigned long long *stats4, long long *trace,
                                     pdae_stream_t stream) {
  if (b <= 0 || !xyz1 || !xyz2 || !dist1 || !dist2 || !idx1 || !idx2) return PDAE_E_INVALID;
  const int lo = n < m ? n : m, hi = n < m ? m : n;
  if (lo < 512 || hi > pdae::TCC_MAXCOLS) return PDAE_E_UNSUPPORTED;  // (the probe takes no workspace: single-chunk clouds)
  (void)pdae::tcc_mode();
  return pdae::chamfer_tc_forward(xyz1, xyz2, b, n, m, dist1, dist2, idx1, idx2, nullptr, 0, static_cast<cudaStream_t>(stream),
                                  stats4, trace);
}
