// edgeconv.cu -- SURVEY.md 8f row 4, first stage: the gather half of an eval-mode EdgeConv layer of the DGCNN encoder
// (models/dgcnn_util.py:114-126: get_graph_feature -> Conv2d(2C, Co, 1, bias=False) -> BatchNorm2d -> LeakyReLU(0.2) ->
// max over the k neighbours).  With the convolution weight split as W = [W1 | W2],
//     W [x_j - x_i; x_i] = W1 x_j + (W2 - W1) x_i = P[j] + Q[i],
// and BatchNorm (running statistics) + LeakyReLU is monotone per channel, increasing where the folded scale s_o >= 0
// and decreasing where s_o < 0, so
//     max_j act(s_o (P[j][o] + Q[i][o]) + t_o) = act(s_o (ext_j P[j][o] + Q[i][o]) + t_o),  ext = max (s_o >= 0) / min.
// P and Q are two small GEMMs (library calls on the host side, N x C x Co per cloud); this kernel does the rest: per
// point it gathers the k neighbour rows of P (row-major (b, n, co): one coalesced row read per neighbour, L2-resident),
// keeps the per-channel extremum, applies the affine + activation and writes the reference's (b, co, n) layout through
// a shared-memory transpose.  The (b, 2C, n, k) graph feature and the (b, Co, n, k) convolution output never exist:
// traffic is b*n*k*co*4 bytes of gathered reads instead of writing and re-reading both tensors.
// Bound: L2 gather bandwidth.  STATUS: bit-exact with its oracle on B200 (tests/test_row4_edgeconv.py); not yet timed.
#include "common.cuh"

namespace pdae {

constexpr int EC_POINTS = 32;   // points per CTA (one output row segment of 128 bytes per channel)
constexpr int EC_WARPS = 8;     // each warp owns EC_POINTS / EC_WARPS points
constexpr int EC_CHUNK = 256;   // channels handled per pass: 8 per lane

__global__ void __launch_bounds__(EC_WARPS * 32) edge_gather_extremum_kernel(const float *__restrict__ P, const float *__restrict__ Q,
                                                                             const int64_t *__restrict__ idx,
                                                                             const float *__restrict__ scale,
                                                                             const float *__restrict__ shift, float slope, int n,
                                                                             int k, int co, float *__restrict__ out) {
  __shared__ float tile[EC_CHUNK][EC_POINTS + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t cloud = blockIdx.y;
  const int i0 = blockIdx.x * EC_POINTS;
  const float *__restrict__ Pc = P + cloud * n * co;
  const float *__restrict__ Qc = Q + cloud * n * co;
  const int64_t *__restrict__ Ic = idx + cloud * n * k;
  constexpr int PER_LANE = EC_CHUNK / 32;
  constexpr int PER_WARP = EC_POINTS / EC_WARPS;

  for (int c0 = 0; c0 < co; c0 += EC_CHUNK) {
    float s[PER_LANE], t[PER_LANE], sgn[PER_LANE];
#pragma unroll
    for (int u = 0; u < PER_LANE; ++u) {
      const int o = c0 + u * 32 + lane;
      s[u] = o < co ? __ldg(scale + o) : 0.0f;
      t[u] = o < co ? __ldg(shift + o) : 0.0f;
      sgn[u] = s[u] >= 0.0f ? 1.0f : -1.0f;  // ext = sgn * max_j(sgn * P): max for s >= 0, min for s < 0 (exact)
    }
    for (int pw = 0; pw < PER_WARP; ++pw) {
      const int pl = warp * PER_WARP + pw;  // point inside the CTA's segment
      const int i = i0 + pl;
      if (i < n) {  // uniform for the warp
        float acc[PER_LANE];
#pragma unroll
        for (int u = 0; u < PER_LANE; ++u) acc[u] = -__int_as_float(0x7f800000);
        for (int j = 0; j < k; ++j) {
          const size_t row = static_cast<size_t>(__ldg(Ic + static_cast<size_t>(i) * k + j)) * co;
#pragma unroll
          for (int u = 0; u < PER_LANE; ++u) {
            const int o = c0 + u * 32 + lane;
            if (o < co) acc[u] = fmaxf(acc[u], __fmul_rn(sgn[u], __ldg(Pc + row + o)));
          }
        }
#pragma unroll
        for (int u = 0; u < PER_LANE; ++u) {
          const int o = c0 + u * 32 + lane;
          if (o < co) {
            const float v = __fadd_rn(__fmul_rn(sgn[u], acc[u]), __ldg(Qc + static_cast<size_t>(i) * co + o));
            float y = __fmaf_rn(s[u], v, t[u]);
            y = y >= 0.0f ? y : __fmul_rn(y, slope);
            tile[u * 32 + lane][pl] = y;
          }
        }
      }
    }
    __syncthreads();
    // (b, co, n) layout: for every channel of the chunk, the CTA's points are contiguous
    for (int e = threadIdx.x; e < EC_CHUNK * EC_POINTS; e += EC_WARPS * 32) {
      const int ol = e / EC_POINTS, pl = e - ol * EC_POINTS;
      const int o = c0 + ol, i = i0 + pl;
      if (o < co && i < n) out[(cloud * co + o) * n + i] = tile[ol][pl];
    }
    __syncthreads();
  }
}

}  // namespace pdae

using namespace pdae;

extern "C" int pdae_edge_gather_extremum_f32(const float *p, const float *q, const int64_t *idx, const float *scale,
                                             const float *shift, float slope, int b, int n, int k, int co, float *out,
                                             pdae_stream_t stream) {
  if (b < 0 || n < 0 || k <= 0 || co <= 0) return PDAE_E_INVALID;
  if (b == 0 || n == 0) return 0;
  if (!p || !q || !idx || !scale || !shift || !out) return PDAE_E_INVALID;
  if (b > 65535) return PDAE_E_UNSUPPORTED;
  const dim3 grid(static_cast<unsigned>((n + EC_POINTS - 1) / EC_POINTS), static_cast<unsigned>(b));
  edge_gather_extremum_kernel<<<grid, EC_WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(p, q, idx, scale, shift, slope, n, k,
                                                                                          co, out);
  PDAE_RETURN_IF_LAUNCH_FAILED();
  return 0;
}
