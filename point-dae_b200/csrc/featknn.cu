// featknn.cu -- DGCNN `knn` dispatch and the fused `get_graph_feature` gather (+ backward).
//
// Semantics follow models/dgcnn_util.py:7-36 of the reference: idx = k nearest neighbours of
// every point in feature space (self included, nearest first); feature = cat(x_j - x_i, x_i)
// returned as a (B, 2C, N, k) view of a physically (B, N, k, 2C) tensor.  The neighbour order is
// this repo's canonical one (direct-form distance, ties -> lower index).
//
// The graph-feature kernels are HBM-bound (output B*N*k*2C*4 bytes): x is transposed once into a
// (B, N, C) workspace so every neighbour row is one contiguous read, and the reference's
// index-select + repeat + cat + permute chain (4 passes) becomes one write pass.
#include "common.cuh"

namespace pdae {

// (b, rows, cols) -> (b, cols, rows), 32x32 shared-memory tiles
__global__ void __launch_bounds__(256) transpose_kernel(const float *__restrict__ in, float *__restrict__ out, int rows,
                                                        int cols) {
  __shared__ float t[32][33];
  const size_t base = static_cast<size_t>(blockIdx.z) * rows * cols;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int r = r0 + ty + i, c = c0 + tx;
    if (r < rows && c < cols) t[ty + i][tx] = __ldg(in + base + static_cast<size_t>(r) * cols + c);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int c = c0 + ty + i, r = r0 + tx;
    if (r < rows && c < cols) out[base + static_cast<size_t>(c) * rows + r] = t[tx][ty + i];
  }
}

static int launch_transpose(const float *in, float *out, int b, int rows, int cols, cudaStream_t st) {
  if (b > 65535) return PDAE_E_UNSUPPORTED;
  const dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32), b);
  if (grid.y > 65535) return PDAE_E_UNSUPPORTED;
  transpose_kernel<<<grid, 256, 0, st>>>(in, out, rows, cols);
  PDAE_RETURN_IF_LAUNCH_FAILED();
  return 0;
}

// one thread per output element; xt is (b, n, c).  out row (b,i,p) = [xt[j] - xt[i], xt[i]].
__global__ void __launch_bounds__(256) graph_feature_kernel(const float *__restrict__ xt, const int64_t *__restrict__ idx,
                                                            int c, int n, int k, long long total,
                                                            float *__restrict__ out) {
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int c2 = 2 * c;
  const long long row = e / c2;  // (b*n + i)*k + p
  const int col = static_cast<int>(e - row * c2);
  const long long bi = row / k;  // b*n + i
  const long long bb = bi / n;
  const int ch = col < c ? col : col - c;
  const float ctr = __ldg(xt + bi * c + ch);
  float v = ctr;
  if (col < c) {
    const long long j = __ldg(idx + row);
    v = __fsub_rn(__ldg(xt + (bb * n + j) * c + ch), ctr);
  }
  out[e] = v;
}

// backward, scatter part: gxt[b, idx, ch] += G[row, ch]          (thread per (row, ch))
__global__ void __launch_bounds__(256) graph_feature_grad_scatter_kernel(const float *__restrict__ gout,
                                                                         const int64_t *__restrict__ idx, int c, int n,
                                                                         int k, long long total,
                                                                         float *__restrict__ gxt) {
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const long long row = e / c;
  const int ch = static_cast<int>(e - row * c);
  const long long bb = row / (static_cast<long long>(k) * n);
  const long long j = __ldg(idx + row);
  atomicAdd(gxt + (bb * n + j) * c + ch, __ldg(gout + row * 2 * c + ch));
}

// backward, centre part: gxt[b, i, ch] += sum_p (G[.., c+ch] - G[.., ch])   (thread per (b,i,ch))
__global__ void __launch_bounds__(256) graph_feature_grad_center_kernel(const float *__restrict__ gout, int c, int k,
                                                                        long long total, float *__restrict__ gxt) {
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const long long bi = e / c;
  const int ch = static_cast<int>(e - bi * c);
  const float *g = gout + bi * k * 2 * c;
  float s = 0.f;
  for (int p = 0; p < k; ++p) s += __ldg(g + static_cast<size_t>(p) * 2 * c + c + ch) - __ldg(g + static_cast<size_t>(p) * 2 * c + ch);
  atomicAdd(gxt + e, s);
}

}  // namespace pdae

using namespace pdae;

extern "C" int pdae_feat_knn_f32(const float *x, int b, int c, int n, int k, int64_t *idx, pdae_stream_t stream) {
  return feat_knn_generic(x, b, c, n, k, idx, static_cast<cudaStream_t>(stream));
}

extern "C" size_t pdae_graph_feature_workspace_bytes(int b, int c, int n) {
  if (b <= 0 || c <= 0 || n <= 0) return 0;
  return static_cast<size_t>(b) * c * n * sizeof(float);
}

extern "C" int pdae_graph_feature_f32(const float *x, const int64_t *idx, int b, int c, int n, int k, float *out,
                                      void *workspace, size_t workspace_bytes, pdae_stream_t stream) {
  if (b < 0 || c <= 0 || n < 0 || k < 0) return PDAE_E_INVALID;
  const long long total = static_cast<long long>(b) * n * k * 2 * c;
  if (total == 0) return 0;
  if (!x || !idx || !out) return PDAE_E_INVALID;
  if (!workspace || workspace_bytes < pdae_graph_feature_workspace_bytes(b, c, n)) return PDAE_E_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float *xt = static_cast<float *>(workspace);
  const int rc = launch_transpose(x, xt, b, c, n, st);  // (b,c,n) -> (b,n,c)
  if (rc) return rc;
  const long long grid = (total + 255) / 256;
  if (grid > 0x7fffffffLL) return PDAE_E_UNSUPPORTED;
  graph_feature_kernel<<<static_cast<unsigned>(grid), 256, 0, st>>>(xt, idx, c, n, k, total, out);
  PDAE_RETURN_IF_LAUNCH_FAILED();
  return 0;
}

extern "C" int pdae_graph_feature_grad_f32(const float *gout, const int64_t *idx, int b, int c, int n, int k,
                                           float *gx, void *workspace, size_t workspace_bytes, pdae_stream_t stream) {
  if (b < 0 || c <= 0 || n < 0 || k < 0) return PDAE_E_INVALID;
  const size_t gsz = static_cast<size_t>(b) * c * n;
  if (gsz == 0) return 0;
  if (!gx) return PDAE_E_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (k == 0) {
    PDAE_CUDA_TRY(cudaMemsetAsync(gx, 0, gsz * sizeof(float), st));
    return 0;
  }
  if (!gout || !idx) return PDAE_E_INVALID;
  if (!workspace || workspace_bytes < pdae_graph_feature_workspace_bytes(b, c, n)) return PDAE_E_WORKSPACE;
  float *gxt = static_cast<float *>(workspace);
  PDAE_CUDA_TRY(cudaMemsetAsync(gxt, 0, gsz * sizeof(float), st));
  const long long tc = static_cast<long long>(b) * n * c;
  const long long ts = tc * k;
  if ((ts + 255) / 256 > 0x7fffffffLL) return PDAE_E_UNSUPPORTED;
  graph_feature_grad_center_kernel<<<static_cast<unsigned>((tc + 255) / 256), 256, 0, st>>>(gout, c, k, tc, gxt);
  PDAE_RETURN_IF_LAUNCH_FAILED();
  graph_feature_grad_scatter_kernel<<<static_cast<unsigned>((ts + 255) / 256), 256, 0, st>>>(gout, idx, c, n, k, ts,
                                                                                          gxt);
  PDAE_RETURN_IF_LAUNCH_FAILED();
  return launch_transpose(gxt, gx, b, n, c, st);  // (b,n,c) -> (b,c,n)
}
