// knn3.cu -- 3-D k nearest neighbours, fast path (knn_cuda.KNN with dim 3, the fused Group tail and
// the first DGCNN EdgeConv layer): exact selection with a lane-local threshold pre-pass.
//
// Same results as knn.cu (ascending by (squared distance, index), distance accumulated as
// fma(dz,dz, fma(dy,dy, dx*dx)) like KNN_CUDA's `ssd += tmp*tmp` loop); different schedule:
//   pass 1  every lane streams its share of the reference cloud (two adjacent points per step,
//           packed FADD2/FMUL2/FFMA2) and keeps only its own t smallest distances (t = 2 for
//           k <= 32, 3 for k <= 64) -- 3 or 5 FMNMX per point, no queue, no ballot;
//           the 32*t lane-local minima are real distances, so their k-th smallest (one bitonic sort
//           of fp32 values across the warp) is an upper bound tau0 of the true k-th distance, and
//           a tight one: typically only ~1.2 k points satisfy d <= tau0;
//   pass 2  the same stream again; points with d <= tau0 are appended to lane-private lists
//           (predicated stores, still no ballot), compacted with one warp scan, sorted as 64-bit
//           (distance, index) keys and the first k written out.
// If a query has more candidates than the lists hold (massive ties), it is re-done with the
// streaming warp-select of knn.cu, so the result is exact for every input.
#include "knn_select.cuh"

namespace pdae {

constexpr int KNN3_LC = 8;  // lane-private candidate slots

struct Knn3Args {
  const float *ref;    // PLANAR ? (b, 3, r) : (b, r, 3)
  const float *query;  // PLANAR ? unused : (b, q, 3)
  float *dist;         // optional, Euclidean
  int64_t *idx;        // optional
  float *group;        // optional (b, q, k, 3): ref[idx] - query
  uint64_t *keys;      // optional (b, q, k): raw (squared-distance bits << 32 | ref_offset + index) for sharded merges
  uint32_t ref_offset; // global index of ref[0] (keys output only)
  int raw_group;       // 1: `group` receives ref[idx] itself instead of ref[idx] - query
  GroupAffine aff;     // aff.mats != NULL: also write the corrupted patches / centres, and `group` becomes
                       // ((x - c) + c) - c, the value the reference's forward ends up with
                       // (models/PointCAE_transformer.py:1011-1017)
  int r, q, k;
  int tile;            // reference points per shared-memory tile (multiple of 64)
  int qpw;             // queries per warp (1 when the cloud spans several tiles)
  int out_kq;
};

template <bool PLANAR, int E /*keys per lane in the final sort: 2 (k<=32) or 4 (k<=64)*/, int NW /*warps per CTA*/,
          bool AFF /*fused corruption epilogue (a.aff)*/>
__global__ void __launch_bounds__(NW * 32) knn3_kernel(const Knn3Args a) {
  constexpr int KNN_WARPS = NW, KNN_THREADS = NW * 32;  // shadow the file-scope defaults
  constexpr int CAP = 32 * E;
  constexpr int NS = E / 2;  // fallback warp-select slots (k <= 32 -> 1, k <= 64 -> 2)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t *queue_all = reinterpret_cast<uint64_t *>(smem_raw);                   // [W][CAP]
  uint64_t *lq_all = queue_all + KNN_WARPS * CAP;                                 // [W][LC][32]
  float *planes = reinterpret_cast<float *>(lq_all + KNN_WARPS * KNN3_LC * 32);   // [3][tile]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cloud = blockIdx.y;
  const int r = a.r, q = a.q, k = a.k, tile = a.tile;
  const float *__restrict__ R = a.ref + static_cast<size_t>(cloud) * r * 3;
  const float *__restrict__ Qp = PLANAR ? R : a.query + static_cast<size_t>(cloud) * q * 3;
  uint64_t *queue = queue_all + warp * CAP;
  uint64_t *lq = lq_all + warp * KNN3_LC * 32;
  const float *sx = planes, *sy = planes + tile, *sz = planes + 2 * tile;
  const int ntiles = (r + tile - 1) / tile;
  const int kslot = (k - 1) >> 5, klane = (k - 1) & 31;
  const float INF = __int_as_float(0x7f800000);

  // stage reference points [tbase, tbase+tn) as planes, padded to a multiple of 64 with x = +inf
  auto load_tile = [&](int tbase, int tn) {
    __syncthreads();
    const int tn64 = (tn + 63) & ~63;
    if (PLANAR) {
      for (int c = 0; c < 3; ++c)
        for (int p = tid; p < tn64; p += KNN_THREADS)
          planes[c * tile + p] = p < tn ? __ldg(R + static_cast<size_t>(c) * r + tbase + p) : (c == 0 ? INF : 0.f);
    } else if ((r & 3) == 0 && (tbase & 3) == 0) {
      // four points = three aligned float4 loads (the cloud base is 16-byte aligned when r % 4 == 0); all loads of a
      // thread are in flight together, then one 128-bit store per plane
      const float4 *src4 = reinterpret_cast<const float4 *>(R + static_cast<size_t>(tbase) * 3);
      const int nquad = tn >> 2;  // whole quads of real points
      for (int g = tid; g < (tn64 >> 2); g += KNN_THREADS) {
        float4 X = make_float4(INF, INF, INF, INF), Y = make_float4(0.f, 0.f, 0.f, 0.f), Z = Y;
        if (g < nquad) {
          const float4 a0 = __ldg(src4 + 3 * g), a1 = __ldg(src4 + 3 * g + 1), a2 = __ldg(src4 + 3 * g + 2);
          X = make_float4(a0.x, a0.w, a1.z, a2.y);
          Y = make_float4(a0.y, a1.x, a1.w, a2.z);
          Z = make_float4(a0.z, a1.y, a2.x, a2.w);
        } else if (4 * g < tn) {  // the last, partial quad
          float *xs = &X.x, *ys = &Y.x, *zs = &Z.x;
          for (int e = 0; e < 4 && 4 * g + e < tn; ++e) {
            const float *pt = R + (static_cast<size_t>(tbase) + 4 * g + e) * 3;
            xs[e] = __ldg(pt); ys[e] = __ldg(pt + 1); zs[e] = __ldg(pt + 2);
          }
        }
        *reinterpret_cast<float4 *>(planes + 4 * g) = X;
        *reinterpret_cast<float4 *>(planes + tile + 4 * g) = Y;
        *reinterpret_cast<float4 *>(planes + 2 * tile + 4 * g) = Z;
      }
    } else {
      const float *src = R + static_cast<size_t>(tbase) * 3;
      for (int f = tid; f < tn64 * 3; f += KNN_THREADS) {
        const int p = f / 3, c = f - p * 3;
        planes[c * tile + p] = p < tn ? __ldg(src + f) : (c == 0 ? INF : 0.f);
      }
    }
    __syncthreads();
  };

  for (int qi = 0; qi < a.qpw; ++qi) {
    const int qidx = (blockIdx.x * KNN_WARPS + warp) * a.qpw + qi;
    const bool qvalid = qidx < q;
    float q0 = 0.f, q1 = 0.f, q2 = 0.f;
    if (qvalid) {
      if (PLANAR) {
        q0 = __ldg(Qp + qidx); q1 = __ldg(Qp + r + qidx); q2 = __ldg(Qp + 2 * r + qidx);
      } else {
        q0 = __ldg(Qp + 3 * qidx); q1 = __ldg(Qp + 3 * qidx + 1); q2 = __ldg(Qp + 3 * qidx + 2);
      }
    }
    const float2 qx2 = make_float2(q0, q0), qy2 = make_float2(q1, q1), qz2 = make_float2(q2, q2);
    auto dist_pair = [&](int j) {  // squared distances of points j, j+1 of the staged tile
      const float2 X = *reinterpret_cast<const float2 *>(sx + j);
      const float2 Y = *reinterpret_cast<const float2 *>(sy + j);
      const float2 Z = *reinterpret_cast<const float2 *>(sz + j);
      const float2 dx = sub2(X, qx2), dy = sub2(Y, qy2), dz = sub2(Z, qz2);
      return fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
    };

    // ---- pass 1: lane-local t smallest distances ------------------------------------------------
    float m0 = INF, m1 = INF, m2 = INF;
    for (int tl = 0; tl < ntiles; ++tl) {
      const int tbase = tl * tile;
      const int tn = (r - tbase) < tile ? (r - tbase) : tile;
      if (ntiles > 1 || qi == 0) load_tile(tbase, tn);
      if (!qvalid) continue;
      const int tn64 = (tn + 63) & ~63;
#pragma unroll 2
      for (int j0 = 0; j0 < tn64; j0 += 64) {
        const float2 d = dist_pair(j0 + 2 * lane);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float v = h ? d.y : d.x;
          const float t1 = fmaxf(m0, v);
          m0 = fminf(m0, v);
          if (E == 2) {
            m1 = fminf(m1, t1);
          } else {
            const float t2 = fmaxf(m1, t1);
            m1 = fminf(m1, t1);
            m2 = fminf(m2, t2);
          }
        }
      }
    }
    float sv[E];
    sv[0] = m0;
    sv[1] = m1;
    if (E == 4) { sv[2] = m2; sv[3] = INF; }
    warp_sort_multi_f32<E>(sv, lane);
    float kth = sv[0];
#pragma unroll
    for (int e = 1; e < E; ++e) kth = (e == kslot) ? sv[e] : kth;
    const float tau0 = __shfl_sync(0xffffffffu, kth, klane);

    // ---- pass 2: collect every point with d <= tau0 into lane-private lists ---------------------
    // Branch-free: a predicated 64-bit store through a bumped pointer per value (most 64-point steps of a warp hold
    // at least one candidate, so a warp-level skip would not pay).  Padding has d = +inf and can only pass the test
    // when tau0 itself is +inf (fewer than k finite distances): that case goes to the exact fallback below.
    int cnt = 0;
    const bool degenerate = !(tau0 < INF);
    for (int tl = 0; tl < ntiles; ++tl) {
      const int tbase = tl * tile;
      const int tn = (r - tbase) < tile ? (r - tbase) : tile;
      if (ntiles > 1) load_tile(tbase, tn);
      if (!qvalid) continue;
      const int tn64 = (tn + 63) & ~63;
      uint32_t jg = static_cast<uint32_t>(tbase + 2 * lane);
#pragma unroll 2
      for (int j0 = 0; j0 < tn64; j0 += 64, jg += 64) {
        const float2 d = dist_pair(j0 + 2 * lane);
        const bool c0 = d.x <= tau0, c1 = d.y <= tau0;
        if (c0 && cnt < KNN3_LC) lq[cnt * 32 + lane] = pack_key(d.x, jg);
        cnt += c0;
        if (c1 && cnt < KNN3_LC) lq[cnt * 32 + lane] = pack_key(d.y, jg + 1);
        cnt += c1;
      }
    }
    if (degenerate && qvalid) cnt = KNN3_LC + 1;  // idle warps (qvalid false) keep cnt = 0 and touch no list
    // compaction: exclusive scan of the lane counts
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    bool overflow = __any_sync(0xffffffffu, cnt > KNN3_LC) || total > CAP;
    if (!qvalid) overflow = false;
    if (ntiles > 1) overflow = __syncthreads_or(overflow);  // tiles are re-streamed by the whole CTA

    uint64_t keys[E];
    if (!overflow) {
      const int off = incl - cnt;
      for (int i = 0; i < cnt; ++i) queue[off + i] = lq[i * 32 + lane];
      __syncwarp();
#pragma unroll
      for (int e = 0; e < E; ++e) keys[e] = (e * 32 + lane) < total ? queue[e * 32 + lane] : KEY_INF;
      __syncwarp();
      warp_sort_multi<E>(keys, lane);
    } else {
      // ---- exact fallback: streaming warp-select over the whole cloud ---------------------------
      WarpSelect<NS> sel;
      sel.init();
      for (int tl = 0; tl < ntiles; ++tl) {
        const int tbase = tl * tile;
        const int tn = (r - tbase) < tile ? (r - tbase) : tile;
        if (ntiles > 1) load_tile(tbase, tn);
        if (!qvalid) continue;
        for (int j0 = 0; j0 < tn; j0 += 32) {
          const int j = j0 + lane;
          const bool in = j < tn;
          const float d = dist_seq3(__fsub_rn(sx[in ? j : 0], q0), __fsub_rn(sy[in ? j : 0], q1), __fsub_rn(sz[in ? j : 0], q2));
          const uint64_t key = pack_key(d, static_cast<uint32_t>(tbase + j));
          sel.offer(in && key < sel.tau, key, queue, lane, kslot, klane);
        }
      }
      sel.finish(queue, lane);
#pragma unroll
      for (int e = 0; e < E; ++e) keys[e] = e < NS ? sel.L[e < NS ? e : 0] : KEY_INF;
    }
    if (!qvalid) continue;

    // ---- epilogue: ascending keys -> outputs ------------------------------------------------------
    const size_t bq = static_cast<size_t>(cloud) * q + qidx;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int p = e * 32 + lane;
      if (p < k) {
        const uint64_t key = keys[e];
        const uint32_t ji = static_cast<uint32_t>(key);
        const size_t o = a.out_kq ? (static_cast<size_t>(cloud) * k + p) * q + qidx : bq * k + p;
        if (a.idx) a.idx[o] = static_cast<int64_t>(ji);
        if (a.dist) a.dist[o] = __fsqrt_rn(__uint_as_float(static_cast<uint32_t>(key >> 32)));
        if (a.keys) a.keys[bq * k + p] = key == KEY_INF ? KEY_INF : key + a.ref_offset;  // fewer than k points: +inf keys
        if (!PLANAR && a.group) {
          // Group: neighbours relative to the centre; dropout_patch_random: the neighbours themselves
          const bool raw = a.raw_group != 0;
          float *g = a.group + (bq * k + p) * 3;
          const float x = __ldg(R + 3 * static_cast<size_t>(ji)), y = __ldg(R + 3 * static_cast<size_t>(ji) + 1);
          const float z = __ldg(R + 3 * static_cast<size_t>(ji) + 2);
          if constexpr (!AFF) {
            g[0] = raw ? x : __fsub_rn(x, q0);
            g[1] = raw ? y : __fsub_rn(y, q1);
            g[2] = raw ? z : __fsub_rn(z, q2);
          } else {
            // fused corrupt_data: the reference re-adds the centre to the centred patch, transforms patch and
            // centre with the same matrices, and subtracts the centres again
            const float *mats = a.aff.mats + static_cast<size_t>(cloud) * a.aff.t * 9;
            float ax = __fadd_rn(__fsub_rn(x, q0), q0), ay = __fadd_rn(__fsub_rn(y, q1), q1);
            float az = __fadd_rn(__fsub_rn(z, q2), q2);
            g[0] = __fsub_rn(ax, q0), g[1] = __fsub_rn(ay, q1), g[2] = __fsub_rn(az, q2);
            float cx = q0, cy = q1, cz = q2;
            affine_seq(mats, a.aff.t, ax, ay, az);
            affine_seq(mats, a.aff.t, cx, cy, cz);
            float *tg = a.aff.tgroup + (bq * k + p) * 3;
            tg[0] = __fsub_rn(ax, cx), tg[1] = __fsub_rn(ay, cy), tg[2] = __fsub_rn(az, cz);
            if (p == 0) {
              float *tc = a.aff.tcenter + bq * 3;
              tc[0] = cx, tc[1] = cy, tc[2] = cz;
            }
          }
        }
      }
    }
  }
}

template <bool PLANAR, int E, int NW, bool AFF>
static int launch_knn3_cfg(const Knn3Args &a, int b, cudaStream_t st) {
  const size_t smem = static_cast<size_t>(NW) * (32 * E + KNN3_LC * 32) * sizeof(uint64_t) +
                      static_cast<size_t>(3) * a.tile * sizeof(float);
  const dim3 grid(ceil_div(a.q, NW * a.qpw), b);
  if (smem > 48 * 1024)
    PDAE_CUDA_TRY(cudaFuncSetAttribute(knn3_kernel<PLANAR, E, NW, AFF>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  knn3_kernel<PLANAR, E, NW, AFF><<<grid, NW * 32, smem, st>>>(a);
  PDAE_RETURN_IF_LAUNCH_FAILED();
  return 0;
}

template <bool PLANAR>
static int launch_knn3(Knn3Args a, int b, cudaStream_t st) {
  if (b > 65535) return PDAE_E_UNSUPPORTED;
  // tile: whole cloud when it fits in 96 KB of planes (8192 points), else 4096-point tiles
  const int r64 = (a.r + 63) & ~63;
  a.tile = r64 <= 8192 ? r64 : 4096;
  // big tiles allow only one CTA per SM: give it 16 warps so every scheduler still has 4 to switch between
  const int nw = a.tile > 4096 ? 16 : 8;
  if (r64 <= 8192) {
    // enough CTAs for >= ~6 per SM while amortising the tile load over a few queries per warp
    long long per = (static_cast<long long>(b) * a.q) / (148LL * nw * 6);
    a.qpw = per < 1 ? 1 : (per > 8 ? 8 : static_cast<int>(per));
  } else {
    a.qpw = 1;
  }
  if constexpr (!PLANAR) {
    if (a.aff.mats != nullptr) {
      if (a.k <= 32) return nw == 16 ? launch_knn3_cfg<false, 2, 16, true>(a, b, st) : launch_knn3_cfg<false, 2, 8, true>(a, b, st);
      return nw == 16 ? launch_knn3_cfg<false, 4, 16, true>(a, b, st) : launch_knn3_cfg<false, 4, 8, true>(a, b, st);
    }
  }
  if (a.k <= 32) return nw == 16 ? launch_knn3_cfg<PLANAR, 2, 16, false>(a, b, st) : launch_knn3_cfg<PLANAR, 2, 8, false>(a, b, st);
  return nw == 16 ? launch_knn3_cfg<PLANAR, 4, 16, false>(a, b, st) : launch_knn3_cfg<PLANAR, 4, 8, false>(a, b, st);
}

// entry points used by knn.cu / featknn.cu dispatch (k <= 64 only)
int knn3_points(const float *ref, const float *query, int b, int r, int q, int k, int out_kq, float *dist, int64_t *idx,
                float *group, cudaStream_t st, uint64_t *keys, uint32_t ref_offset, int raw_group,
                const GroupAffine *affine) {
  Knn3Args a{ref, query, dist, idx, group, keys, ref_offset, raw_group, affine ? *affine : GroupAffine{nullptr, 0, nullptr, nullptr},
             r, q, k, 0, 1, out_kq};
  return launch_knn3<false>(a, b, st);
}
int knn3_planar(const float *x, int b, int n, int k, int64_t *idx, cudaStream_t st) {
  Knn3Args a{x, nullptr, nullptr, idx, nullptr, nullptr, 0u, 0, GroupAffine{nullptr, 0, nullptr, nullptr}, n, n, k, 0, 1, 0};
  return launch_knn3<true>(a, b, st);
}

}  // namespace pdae
