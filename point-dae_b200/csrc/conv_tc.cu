// conv_tc.cu -- 1x1 convolutions of the path's dense consumers on the 5th-generation tensor cores (SURVEY.md 8f row 4:
// the EdgeConv layers of `dgcnn_encoder`, models/dgcnn_util.py:96-133, and the patch `Encoder`,
// models/PointCAE_transformer.py:20-51).
//
//   Z[b][j][n] = sum_c W[j][c] * X[b][c][n]            X (B,C,N) channel-major, W (J,C) row-major, Z (B,J,N)
//
// written for sm_100a by hand: tcgen05.mma kind::tf32 issued by one elected thread, operands in shared memory in the
// K-major no-swizzle canonical layout (8-row x 16-byte core matrices), the fp32 accumulator tiles (128 points x up to 256
// output channels, one for the main and one for the correction terms) in TENSOR MEMORY, read back with tcgen05.ld for the epilogue; completion through tcgen05.commit on an
// mbarrier.  fp32 accuracy comes from the 3xTF32 split: every operand is staged twice, hi = the value truncated to
// tf32 and lo = value - hi (exact in fp32), and each K-step issues hi*hi + hi*lo + lo*hi into the same accumulator
// (the dropped lo*lo term is 2^-22 relative), so the result holds the reference's 1e-5 bound without changing any
// neighbour order downstream.  The split is arithmetic on every element, so the activations are staged by the CTA's
// threads (coalesced loads along the points, one 128-bit shared store per 4 channels -- the (C,N) -> K-major
// transposition is free here); the weights are split once per launch and arrive with the TMA unit (cp.async.bulk).
#include "common.cuh"

namespace pdae {

namespace tc {
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
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
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor): start address, leading byte offset
// (between the two 16-byte K chunks of one MMA), stride byte offset (between 8-row groups), all in 16-byte units
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3ffffu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell)
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A / B tf32, both K-major, M x N
__host__ __device__ constexpr uint32_t instr_desc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
}  // namespace tc

// Weights split once per launch into the shared-memory image the MMA wants: for every K chunk of KC channels and every
// N tile, [hi | lo] x [KC/4 sixteen-byte columns][NT/8 row groups][8 rows][4 floats] -- one contiguous block per
// (n tile, k chunk), so a single bulk copy brings it.
constexpr int TC_M = 128;   // points per CTA tile (MMA M)
constexpr int TC_KC = 32;   // channels per K chunk

__global__ void __launch_bounds__(256) conv_pack_weights_kernel(const float *__restrict__ w, int j, int c, int nt, int cpad,
                                                                float *__restrict__ packed) {
  // packed index space: (ntile, kchunk, part(hi/lo), kc16, rowgroup, row8, e4)
  const int nchunks = cpad / TC_KC, ntiles = (j + nt - 1) / nt;
  const long long total = static_cast<long long>(ntiles) * nchunks * 2 * TC_KC * nt;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long r = t;
    const int e4 = static_cast<int>(r % 4); r /= 4;
    const int row8 = static_cast<int>(r % 8); r /= 8;
    const int rg = static_cast<int>(r % (nt / 8)); r /= (nt / 8);
    const int kc16 = static_cast<int>(r % (TC_KC / 4)); r /= (TC_KC / 4);
    const int part = static_cast<int>(r % 2); r /= 2;
    const int kch = static_cast<int>(r % nchunks); r /= nchunks;
    const int ntile = static_cast<int>(r);
    const int row = ntile * nt + rg * 8 + row8, ch = kch * TC_KC + kc16 * 4 + e4;
    const float v = (row < j && ch < c) ? w[static_cast<size_t>(row) * c + ch] : 0.f;
    const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    packed[t] = part == 0 ? hi : __fsub_rn(v, hi);
  }
}

// One CTA = 128 points of one cloud x NT output channels.  128 threads (the four warps own the four 32-lane quarters of the
// accumulator in the epilogue).
// IN_PM / OUT_PM: the activations / the result are point-major ((b, n, c) / (b, n, j)) instead of channel-major.
template <int NT, bool IN_PM, bool OUT_PM>
__global__ void __launch_bounds__(128) conv1x1_tc_kernel(const float *__restrict__ x, const float *__restrict__ wpacked,
                                                        const float *__restrict__ bias, int c, int cpad, int n, int j,
                                                        float *__restrict__ z) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t *bar_w = reinterpret_cast<uint64_t *>(smem_raw);       // weights landed (TMA)
  uint64_t *bar_mma = bar_w + 1;                                  // MMAs of the chunk retired
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_w + 2);
  float *a_hi = reinterpret_cast<float *>(smem_raw + 128);        // [KC/4][16 row groups][8][4]
  float *a_lo = a_hi + TC_M * TC_KC;
  float *b_img = a_lo + TC_M * TC_KC;                             // [hi | lo][KC/4][NT/8][8][4]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * TC_M, ntile = blockIdx.y, cloud = blockIdx.z;
  const int nchunks = cpad / TC_KC;
  const float *X = x + static_cast<size_t>(cloud) * c * n;

  if (tid == 0) {
    tc::mbar_init(bar_w, 1);
    tc::mbar_init(bar_mma, 1);
  }
  if (warp == 0) {  // one warp allocates the accumulator columns and lets other CTAs allocate too
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_slot)), "n"(2 * NT));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  constexpr uint32_t IDESC = tc::instr_desc_tf32(TC_M, NT);
  constexpr uint32_t B_PART = NT * TC_KC;  // floats per hi / lo image

  for (int kch = 0; kch < nchunks; ++kch) {
    const uint32_t par = kch & 1;
    if (kch > 0) {  // the previous chunk's MMAs have finished reading shared memory
      tc::mbar_wait(bar_mma, par ^ 1u);
    }
    __syncthreads();
    if (tid == 0) {  // weights of this (n tile, k chunk): one contiguous block
      const float *src = wpacked + (static_cast<size_t>(ntile) * nchunks + kch) * (2 * B_PART);
      tc::mbar_expect_tx(bar_w, 2 * B_PART * 4);
      tc::bulk_g2s(b_img, src, 2 * B_PART * 4, bar_w);
    }
    // activations: thread t stages point n0 + t for all KC channels (4 channels -> one 16-byte element of its row)
    {
      const int p = n0 + tid;
      const bool in = p < n;
      const int rg = tid >> 3, row8 = tid & 7;
#pragma unroll 2
      for (int kc16 = 0; kc16 < TC_KC / 4; ++kc16) {
        float v[4];
        const int ch0 = kch * TC_KC + kc16 * 4;
        if (IN_PM && in && (c & 3) == 0 && ch0 + 3 < c) {  // a point's channels are contiguous: one 128-bit load
          const float4 t = __ldg(reinterpret_cast<const float4 *>(X + static_cast<size_t>(p) * c + ch0));
          v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int ch = ch0 + e;
            v[e] = (in && ch < c) ? __ldg(IN_PM ? X + static_cast<size_t>(p) * c + ch : X + static_cast<size_t>(ch) * n + p) : 0.f;
          }
        }
        float4 hi, lo;
        hi.x = __uint_as_float(__float_as_uint(v[0]) & 0xffffe000u), lo.x = __fsub_rn(v[0], hi.x);
        hi.y = __uint_as_float(__float_as_uint(v[1]) & 0xffffe000u), lo.y = __fsub_rn(v[1], hi.y);
        hi.z = __uint_as_float(__float_as_uint(v[2]) & 0xffffe000u), lo.z = __fsub_rn(v[2], hi.z);
        hi.w = __uint_as_float(__float_as_uint(v[3]) & 0xffffe000u), lo.w = __fsub_rn(v[3], hi.w);
        const int o = ((kc16 * (TC_M / 8) + rg) * 8 + row8) * 4;
        *reinterpret_cast<float4 *>(a_hi + o) = hi;
        *reinterpret_cast<float4 *>(a_lo + o) = lo;
      }
    }
    tc::fence_proxy_async();  // the generic-proxy stores above must be visible to the tensor core's reads
    __syncthreads();
    if (warp == 0) {
      tc::mbar_wait(bar_w, par);
      tc::tc_fence_after();
      if (lane == 0) {
        const uint32_t a_hi_s = tc::smem_u32(a_hi), a_lo_s = tc::smem_u32(a_lo);
        const uint32_t b_hi_s = tc::smem_u32(b_img), b_lo_s = b_hi_s + B_PART * 4;
        constexpr uint32_t A_LBO = (TC_M / 8) * 128, B_LBO = (NT / 8) * 128, SBO = 128;
#pragma unroll
        for (int ks = 0; ks < TC_KC / 8; ++ks) {  // one MMA = 8 channels = two 16-byte columns
          const uint64_t ah = tc::smem_desc(a_hi_s + ks * 2 * A_LBO, A_LBO, SBO), al = tc::smem_desc(a_lo_s + ks * 2 * A_LBO, A_LBO, SBO);
          const uint64_t bh = tc::smem_desc(b_hi_s + ks * 2 * B_LBO, B_LBO, SBO), bl = tc::smem_desc(b_lo_s + ks * 2 * B_LBO, B_LBO, SBO);
          // two accumulator tiles: the correction terms (2^-11 of the main term) are summed on their own, so they are
          // not rounded away against a large running sum; the epilogue adds the two tiles once
          const uint32_t first = (kch > 0 || ks > 0) ? 1u : 0u;
          tc::mma_tf32(tmem_d + NT, al, bh, IDESC, first);
          tc::mma_tf32(tmem_d + NT, ah, bl, IDESC, 1u);
          tc::mma_tf32(tmem_d, ah, bh, IDESC, first);
        }
        tc::mma_commit(bar_mma);  // arrives when every MMA issued so far has completed (implicit before_thread_sync)
      }
      __syncwarp();
    }
  }
  // ---- epilogue: TMEM -> registers -> Z (coalesced along the points for every output channel) --------------------------
  tc::mbar_wait(bar_mma, (nchunks - 1) & 1);
  tc::tc_fence_after();
  {
    const int p = n0 + warp * 32 + lane;
    float *Z = z + static_cast<size_t>(cloud) * j * n;
#pragma unroll 1
    for (int c0 = 0; c0 < NT; c0 += 16) {
      uint32_t r[16], s[16];
      tc::tmem_ld16(tmem_d + (static_cast<uint32_t>(warp * 32) << 16) + c0, r);
      tc::tmem_ld16(tmem_d + (static_cast<uint32_t>(warp * 32) << 16) + NT + c0, s);
      if (p < n) {
        float v[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = __fadd_rn(__uint_as_float(r[e]), __uint_as_float(s[e]));
        const int jj0 = ntile * NT + c0;
        if (bias) {
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (jj0 + e < j) v[e] = __fadd_rn(v[e], __ldg(bias + jj0 + e));
        }
        if (OUT_PM && (j & 3) == 0 && jj0 + 15 < j) {  // 64 contiguous bytes of the point's row
          float4 *dst = reinterpret_cast<float4 *>(Z + static_cast<size_t>(p) * j + jj0);
#pragma unroll
          for (int e = 0; e < 4; ++e) dst[e] = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int jj = jj0 + e;
            if (jj < j) Z[OUT_PM ? static_cast<size_t>(p) * j + jj : static_cast<size_t>(jj) * n + p] = v[e];
          }
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(2 * NT));
}

static int conv_tile_n(int j) { return j <= 128 ? 128 : 256; }

}  // namespace pdae

using namespace pdae;

extern "C" size_t pdae_conv1x1_workspace_bytes(int c, int j) {
  if (c <= 0 || j <= 0) return 0;
  const int nt = conv_tile_n(j), cpad = (c + TC_KC - 1) / TC_KC * TC_KC;
  const int ntiles = (j + nt - 1) / nt;
  return static_cast<size_t>(ntiles) * nt * cpad * 2 * sizeof(float);
}

template <int NT, bool IN_PM, bool OUT_PM>
static int conv_launch(const float *x, const float *packed, const float *bias, int b, int c, int cpad, int n, int j, int ntiles,
                       float *z, cudaStream_t st) {
  const dim3 grid(ceil_div(n, TC_M), ntiles, b);
  const size_t smem = 128 + static_cast<size_t>(2) * TC_M * TC_KC * 4 + static_cast<size_t>(2) * NT * TC_KC * 4;
  // per launch: the attribute belongs to the current device's instance of the kernel (a process may drive several GPUs)
  PDAE_CUDA_TRY(cudaFuncSetAttribute(conv1x1_tc_kernel<NT, IN_PM, OUT_PM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem)));
  conv1x1_tc_kernel<NT, IN_PM, OUT_PM><<<grid, 128, smem, st>>>(x, packed, bias, c, cpad, n, j, z);
  PDAE_RETURN_IF_LAUNCH_FAILED();
  return 0;
}

extern "C" int pdae_conv1x1_tf32x3_f32(const float *x, const float *w, const float *bias, int b, int c, int n, int j,
                                       int in_point_major, int out_point_major, float *z, void *workspace,
                                       size_t workspace_bytes, pdae_stream_t stream) {
  if (b < 0 || c <= 0 || n < 0 || j <= 0) return PDAE_E_INVALID;
  if (b == 0 || n == 0) return 0;
  if (!x || !w || !z || !workspace) return PDAE_E_INVALID;
  if (workspace_bytes < pdae_conv1x1_workspace_bytes(c, j)) return PDAE_E_INVALID;
  if (b > 65535) return PDAE_E_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nt = conv_tile_n(j), cpad = (c + TC_KC - 1) / TC_KC * TC_KC;
  const int ntiles = (j + nt - 1) / nt;
  float *packed = static_cast<float *>(workspace);
  const long long total = static_cast<long long>(ntiles) * nt * cpad * 2;
  conv_pack_weights_kernel<<<static_cast<unsigned>((total + 255) / 256 > 1184 ? 1184 : (total + 255) / 256), 256, 0, st>>>(
      w, j, c, nt, cpad, packed);
  PDAE_RETURN_IF_LAUNCH_FAILED();
  const int sel = (nt == 256 ? 4 : 0) | (in_point_major ? 2 : 0) | (out_point_major ? 1 : 0);
  switch (sel) {
    case 0: return conv_launch<128, false, false>(x, packed, bias, b, c, cpad, n, j, ntiles, z, st);
    case 1: return conv_launch<128, false, true>(x, packed, bias, b, c, cpad, n, j, ntiles, z, st);
    case 2: return conv_launch<128, true, false>(x, packed, bias, b, c, cpad, n, j, ntiles, z, st);
    case 3: return conv_launch<128, true, true>(x, packed, bias, b, c, cpad, n, j, ntiles, z, st);
    case 4: return conv_launch<256, false, false>(x, packed, bias, b, c, cpad, n, j, ntiles, z, st);
    case 5: return conv_launch<256, false, true>(x, packed, bias, b, c, cpad, n, j, ntiles, z, st);
    case 6: return conv_launch<256, true, false>(x, packed, bias, b, c, cpad, n, j, ntiles, z, st);
    default: return conv_launch<256, true, true>(x, packed, bias, b, c, cpad, n, j, ntiles, z, st);
  }
}
