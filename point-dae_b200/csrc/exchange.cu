// exchange.cu -- the exchange step of reference-set-sharded Chamfer (BASELINE config 5) as ONE kernel over
// NVLink peer memory instead of a collective call: every rank keeps its packed row keys
// (float_bits(min d) << 32 | global argmin, ascending order = (distance, lower index)) in a SYMMETRIC buffer; after a
// cross-rank barrier each rank reduces ITS slice of the rows over all ranks' buffers and writes the unpacked
// (distance, index) pairs straight into every rank's result buffer -- reduce-scatter + all-gather + unpack fused:
//   * peer form: W peer loads (ld.volatile over NVLink) per row, min, W peer stores;
//   * in-switch form (NVLS, when the symmetric allocation has a multicast address): ONE multimem.ld_reduce.min.u64 per
//     row -- the NVSwitch reduces the W copies on the way -- and one multimem.st per output, broadcast by the switch.
// 0.8 MB of keys at N = 100 000: the NCCL all-reduce it replaces is latency-bound at this size; here the traffic per rank
// is 2 x N / W x 8 bytes and the latency two barriers and one short kernel.
#include "common.cuh"

namespace pdae {

constexpr int EX_MAX_WORLD = 16;

struct ExchangePeers {
  const uint64_t *keys[EX_MAX_WORLD];
  float *dist[EX_MAX_WORLD];
  int *idx[EX_MAX_WORLD];
};

__global__ void __launch_bounds__(256) exchange_keys_peer_kernel(const ExchangePeers p, int world, long long lo, long long hi) {
  for (long long row = lo + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; row < hi;
       row += static_cast<long long>(gridDim.x) * blockDim.x) {
    uint64_t k = 0xffffffffffffffffull;
    for (int w = 0; w < world; ++w) {
      const uint64_t v = *reinterpret_cast<const volatile uint64_t *>(p.keys[w] + row);  // never served from a stale L1 line
      k = v < k ? v : k;
    }
    const float d = __uint_as_float(static_cast<uint32_t>(k >> 32));
    const int i = static_cast<int>(static_cast<uint32_t>(k));
    for (int w = 0; w < world; ++w) {
      p.dist[w][row] = d;
      p.idx[w][row] = i;
    }
  }
  __threadfence_system();
}

// multicast addresses: one load reduces over every rank's copy inside the switch, one store reaches every rank
__global__ void __launch_bounds__(256) exchange_keys_multimem_kernel(const uint64_t *mc_keys, float *mc_dist, int *mc_idx,
                                                                     long long lo, long long hi) {
  for (long long row = lo + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; row < hi;
       row += static_cast<long long>(gridDim.x) * blockDim.x) {
    uint64_t k;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.min.u64 %0, [%1];" : "=l"(k) : "l"(mc_keys + row) : "memory");
    const uint32_t dbits = static_cast<uint32_t>(k >> 32), ibits = static_cast<uint32_t>(k);
    asm volatile("multimem.st.relaxed.sys.global.b32 [%0], %1;" ::"l"(mc_dist + row), "r"(dbits) : "memory");
    asm volatile("multimem.st.relaxed.sys.global.b32 [%0], %1;" ::"l"(mc_idx + row), "r"(ibits) : "memory");
  }
  __threadfence_system();
}

}  // namespace pdae

using namespace pdae;

// keys / dist / idx: host arrays of `world` device pointers (this rank's own buffers included, at index `rank`), all
// peer-mapped (torch symmetric memory / cudaIpc / cuMem); rows [lo, hi) are this rank's share.  The caller brackets the
// call with cross-rank barriers (every rank's keys written before, every rank's results read after).
extern "C" int pdae_chamfer_exchange_keys_peer(const void *const *keys, void *const *dist, void *const *idx, int world,
                                               long long lo, long long hi, pdae_stream_t stream) {
  if (world <= 0 || world > EX_MAX_WORLD || lo < 0 || hi < lo) return PDAE_E_INVALID;
  if (hi == lo) return 0;
  if (!keys || !dist || !idx) return PDAE_E_INVALID;
  ExchangePeers p;
  for (int w = 0; w < world; ++w) {
    if (!keys[w] || !dist[w] || !idx[w]) return PDAE_E_INVALID;
    p.keys[w] = static_cast<const uint64_t *>(keys[w]);
    p.dist[w] = static_cast<float *>(dist[w]);
    p.idx[w] = static_cast<int *>(idx[w]);
  }
  const long long n = hi - lo;
  const unsigned grid = static_cast<unsigned>((n + 255) / 256 > 4 * 148 ? 4 * 148 : (n + 255) / 256);
  exchange_keys_peer_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p, world, lo, hi);
  PDAE_RETURN_IF_LAUNCH_FAILED();
  return 0;
}

extern "C" int pdae_chamfer_exchange_keys_multimem(const void *mc_keys, void *mc_dist, void *mc_idx, long long lo, long long hi,
                                                   pdae_stream_t stream) {
  if (lo < 0 || hi < lo) return PDAE_E_INVALID;
  if (hi == lo) return 0;
  if (!mc_keys || !mc_dist || !mc_idx) return PDAE_E_INVALID;
  const long long n = hi - lo;
  const unsigned grid = static_cast<unsigned>((n + 255) / 256 > 4 * 148 ? 4 * 148 : (n + 255) / 256);
  exchange_keys_multimem_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint64_t *>(mc_keys), static_cast<float *>(mc_dist), static_cast<int *>(mc_idx), lo, hi);
  PDAE_RETURN_IF_LAUNCH_FAILED();
  return 0;
}
