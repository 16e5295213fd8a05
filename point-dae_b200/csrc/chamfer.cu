// chamfer.cu -- Chamfer distance forward (fused min/argmin, both directions in one launch),
// backward (gradient scatter through the saved argmin) and the reference-set-sharded variant
// that emits packed (distance, index) keys for a MIN all-reduce.
//
// Semantics follow extensions/chamfer_dist/chamfer.cu:15-145 (forward) and :173-201 (backward)
// of the reference: dist = min_k fma(dz,dz, fma(dx,dx, dy*dy)) with d* = b_k - a, idx = lowest k
// attaining it; backward g = 2*grad_dist, v = g*(a - b_idx), gx1[j] += v, gx2[idx] -= v.
//
// Design (FP32 FMA-pipe bound):
//   * a CTA owns QT*THREADS query points held in registers (each coordinate duplicated into a
//     64-bit register pair) and streams the other cloud through shared memory in planar
//     (SoA) tiles, so one LDS.128 feeds two packed FADD2/FMUL2/FFMA2 evaluations per query;
//   * the inner loop carries no index: the running minimum is updated with one FMNMX3 per two
//     pairs; the only bookkeeping is "which group of 16 reference points last lowered the
//     minimum" (one FSETP + SEL per sixteen pairs).  Every instruction that is not FMA-pipe work
//     costs a dispatch slot the FMA pipe could have used (packed ops take two), so the loop is
//     built to minimise them: 6 FMA-pipe slots + ~1.06 other slots per point pair;
//   * after the scan every thread re-evaluates the single 16-point group recorded for each of
//     its queries and takes the first point whose distance equals the minimum bit-for-bit ->
//     the lowest index, as the reference's strict `<` does.
#include <stdlib.h>

#include "common.cuh"

namespace pdae {

constexpr int CH_GROUP = 16;  // argmin bookkeeping granularity: four LDS.128 steps = 16 reference points

struct ChamferDir {
  const float *q;   // (b, nq, 3) query cloud
  const float *r;   // (b, nr, 3) reference cloud (or local slice)
  float *dist;      // (b, nq) or null
  int *idx;         // (b, nq) or null
  uint64_t *keys;   // (b, nq) packed output (sharded mode) or null
  uint64_t *colkeys;  // (b, nr) symmetric mode: (min over the queries, 32*QT-query group id), RED.MIN target
  int nq, nr;
  int qtiles;       // CTAs per cloud for this direction (0 = direction absent)
  int ref_offset;   // global index of r[0] (sharded mode)
  int csplit;       // column chunks per row block (0 or 1 = a CTA streams the whole reference cloud); > 1 needs `keys`,
                    // pre-filled with the MIN identity: the chunks merge their (row minimum, lowest index) with RED.MIN
};

// SYM: the squared distance is symmetric bit for bit (the operands of every product only change
// sign), so one evaluation of pair (a_i, b_j) serves both directions: besides the per-query running
// minimum the kernel reduces, for every reference point, the minimum over the warp's 32*QT queries
// (FMNMX3 across the thread's queries, one REDUX.MIN across the lanes) and posts
// (min bits << 32 | query-group id) with a 64-bit RED.MIN per CTA; chamfer_col_recover_kernel then
// finds the lowest query index inside the winning group.  Halves the FMA-pipe work of the forward.
// One pass of a CTA: rows [row_base, row_base + QT*THREADS) of cloud `Q` (clamped to nq) against all nr points of
// `R`.  VARGROUP selects how a warp names itself in the column keys: false = id of its 32*QT-row group
// (row_base / (32*QT) + warp, uniform groups); true = (first row / 32) << 3 | QT, for the balanced kernel whose
// passes have different QT.  The shared-memory buffers belong to the caller, so several instantiations can share them.
template <int QT, int THREADS, bool SYM, int CH_TILE, int STEP, bool VARGROUP, bool PACKED = true>
__device__ __forceinline__ void chamfer_min_body(const float *__restrict__ Q, const float *__restrict__ R, int nq, int nr,
                                                 int row_base, size_t cloud, float *dist, int *idx, uint64_t *keys,
                                                 int ref_offset, uint64_t *colkeys, float (*tile)[3][CH_TILE],
                                                 unsigned (*colmin)[THREADS / 32][SYM ? CH_TILE : 4], int tile_begin = 0,
                                                 int tile_end = 0x7fffffff, bool merge = false) {
  constexpr int QPW = 32 * QT;                  // queries per warp
  constexpr int W = THREADS / 32;
  constexpr int LD = 3 * CH_TILE / THREADS;     // floats staged per thread per tile
  static_assert(3 * CH_TILE % THREADS == 0, "tile must split evenly over the CTA");
  static_assert(CH_TILE % CH_GROUP == 0, "tile must hold whole bookkeeping groups");

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qbase = row_base + warp * QPW;

  float2 qx[QT], qy[QT], qz[QT];
  float best[QT];
  int bgrp[QT];
#pragma unroll
  for (int s = 0; s < QT; ++s) {
    int q = qbase + s * 32 + lane;
    q = q < nq ? q : nq - 1;
    const float x = __ldg(Q + 3 * q), y = __ldg(Q + 3 * q + 1), z = __ldg(Q + 3 * q + 2);
    qx[s] = make_float2(x, x);
    qy[s] = make_float2(y, y);
    qz[s] = make_float2(z, z);
    best[s] = __int_as_float(0x7f800000);
    bgrp[s] = 0;
  }

  float pre[LD];
  const int nr3 = nr * 3;
  auto fetch = [&](int tile_base) {
#pragma unroll
    for (int i = 0; i < LD; ++i) {
      const int f = tid + i * THREADS;
      const int g = tile_base * 3 + f;
      // padding: x = +inf makes the padded distance +inf, which never lowers a minimum
      pre[i] = g < nr3 ? __ldg(R + g) : ((f % 3 == 0) ? __int_as_float(0x7f800000) : 0.0f);
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < LD; ++i) {
      const int f = tid + i * THREADS;
      const int p = f / 3, c = f - 3 * p;
      tile[buf][c][p] = pre[i];
    }
  };

  // SYM: post the per-warp column minima of tile `tp` (complete since the last barrier)
  auto flush_cols = [&](int tp) {
    if (!SYM) return;
    uint64_t *ck = colkeys + cloud * nr;
    const int tb = tp * CH_TILE;
    for (int j = tid; j < CH_TILE && tb + j < nr; j += THREADS) {
      unsigned v = colmin[tp & 1][0][j];
      int wm = 0;
#pragma unroll
      for (int w = 1; w < W; ++w) {
        const unsigned u = colmin[tp & 1][w][j];
        wm = u < v ? w : wm;
        v = u < v ? u : v;
      }
      const unsigned who = VARGROUP ? (static_cast<unsigned>((row_base + wm * QPW) >> 5) << 3) | static_cast<unsigned>(QT)
                                    : static_cast<unsigned>(row_base / QPW + wm);
      atomicMin(reinterpret_cast<unsigned long long *>(ck + tb + j), (static_cast<unsigned long long>(v) << 32) | who);
    }
  };

  // a CTA streams tiles [tile_begin, ntiles) of the reference cloud: all of them, or one column chunk of a split unit
  const int ntiles_all = (nr + CH_TILE - 1) / CH_TILE;
  const int ntiles = tile_end < ntiles_all ? tile_end : ntiles_all;
  fetch(tile_begin * CH_TILE);
  stash(tile_begin & 1);
  __syncthreads();
  for (int tl = tile_begin; tl < ntiles; ++tl) {
    const bool more = tl + 1 < ntiles;
    if (more) fetch((tl + 1) * CH_TILE);
    if (tl > tile_begin) flush_cols(tl - 1);
    const float *sx = tile[tl & 1][0], *sy = tile[tl & 1][1], *sz = tile[tl & 1][2];
    unsigned pend[STEP];
#pragma unroll
    for (int r = 0; r < STEP; ++r) pend[r] = 0u;
    int pend_off = -1;
    auto store_pend = [&]() {  // lane 0 parks the warp's column minima of one step
      uint4 *dst = reinterpret_cast<uint4 *>(&colmin[tl & 1][warp][pend_off]);
#pragma unroll
      for (int r = 0; r < STEP; r += 4) dst[r >> 2] = make_uint4(pend[r], pend[r + 1], pend[r + 2], pend[r + 3]);
    };
    const int left = nr - tl * CH_TILE;
    const int ngroups = ((left < CH_TILE ? left : CH_TILE) + CH_GROUP - 1) / CH_GROUP;
    const int group_base = tl * (CH_TILE / CH_GROUP);
#pragma unroll 2
    for (int g = 0; g < ngroups; ++g) {
      const int gid = group_base + g;
      float cur[QT];
#pragma unroll
      for (int s = 0; s < QT; ++s) cur[s] = best[s];
#pragma unroll
      for (int h = 0; h < CH_GROUP; h += STEP) {
        float4 X[STEP / 4], Y[STEP / 4], Z[STEP / 4];
#pragma unroll
        for (int v = 0; v < STEP / 4; ++v) {
          X[v] = *reinterpret_cast<const float4 *>(sx + g * CH_GROUP + h + 4 * v);
          Y[v] = *reinterpret_cast<const float4 *>(sy + g * CH_GROUP + h + 4 * v);
          Z[v] = *reinterpret_cast<const float4 *>(sz + g * CH_GROUP + h + 4 * v);
        }
        float2 dd[QT][STEP / 2];
#pragma unroll
        for (int s = 0; s < QT; ++s) {
#pragma unroll
          for (int v = 0; v < STEP / 4; ++v) {
            if (PACKED) {
              dd[s][2 * v] = dist_yxz2(sub2(make_float2(X[v].x, X[v].y), qx[s]), sub2(make_float2(Y[v].x, Y[v].y), qy[s]),
                                       sub2(make_float2(Z[v].x, Z[v].y), qz[s]));
              dd[s][2 * v + 1] = dist_yxz2(sub2(make_float2(X[v].z, X[v].w), qx[s]), sub2(make_float2(Y[v].z, Y[v].w), qy[s]),
                                           sub2(make_float2(Z[v].z, Z[v].w), qz[s]));
            } else {  // scalar FADD/FMUL/FFMA (same values): experiment for the issue-slot model, see DESIGN.md
              const float ax = qx[s].x, ay = qy[s].x, az = qz[s].x;
              dd[s][2 * v] = make_float2(dist_yxz(__fsub_rn(X[v].x, ax), __fsub_rn(Y[v].x, ay), __fsub_rn(Z[v].x, az)),
                                         dist_yxz(__fsub_rn(X[v].y, ax), __fsub_rn(Y[v].y, ay), __fsub_rn(Z[v].y, az)));
              dd[s][2 * v + 1] = make_float2(dist_yxz(__fsub_rn(X[v].z, ax), __fsub_rn(Y[v].z, ay), __fsub_rn(Z[v].z, az)),
                                             dist_yxz(__fsub_rn(X[v].w, ax), __fsub_rn(Y[v].w, ay), __fsub_rn(Z[v].w, az)));
            }
          }
          if (STEP == 8) {
            const float t0 = min3(dd[s][0].x, dd[s][0].y, dd[s][1].x);
            const float t1 = min3(dd[s][1].y, dd[s][2].x, dd[s][2].y);
            const float t2 = min3(dd[s][STEP / 2 - 1].x, dd[s][STEP / 2 - 1].y, t0);
            cur[s] = min3(cur[s], t1, t2);
          } else {
            const float t0 = min3(dd[s][0].x, dd[s][0].y, dd[s][1].x);
            cur[s] = min3(cur[s], dd[s][1].y, t0);
          }
        }
        if (SYM) {  // column minima over this warp's 32*QT queries for the STEP reference points of the step
          // the REDUX results of the previous step are stored only now, a full step of FMA work later, so
          // their latency never stalls the warp
          if (pend_off >= 0 && lane == 0) store_pend();
#pragma unroll
          for (int r = 0; r < STEP; ++r) {
            float c = (r & 1) ? dd[0][r >> 1].y : dd[0][r >> 1].x;
#pragma unroll
            for (int s = 1; s + 1 < QT; s += 2)
              c = min3(c, (r & 1) ? dd[s][r >> 1].y : dd[s][r >> 1].x, (r & 1) ? dd[s + 1][r >> 1].y : dd[s + 1][r >> 1].x);
            if ((QT & 1) == 0) c = fminf(c, (r & 1) ? dd[QT - 1][r >> 1].y : dd[QT - 1][r >> 1].x);
            pend[r] = __reduce_min_sync(0xffffffffu, __float_as_uint(c));
          }
          pend_off = g * CH_GROUP + h;
        }
      }
#pragma unroll
      for (int s = 0; s < QT; ++s) {
        bgrp[s] = cur[s] < best[s] ? gid : bgrp[s];
        best[s] = cur[s];
      }
    }
    if (SYM && pend_off >= 0 && lane == 0) store_pend();
    if (more) stash((tl + 1) & 1);
    __syncthreads();
  }
  flush_cols(ntiles - 1);

  // ---- index recovery: each thread rescans the one 16-point group recorded per query ----------
  // (scanned back to front so the lowest matching index is the one kept)
  int myidx[QT];
  const bool vec_ok = (nr & 3) == 0;  // every cloud then starts 16-byte aligned
#pragma unroll
  for (int s = 0; s < QT; ++s) {
    const int base = bgrp[s] * CH_GROUP;
    int found = 0;
    if (vec_ok && base + CH_GROUP <= nr) {
      const float4 *p4 = reinterpret_cast<const float4 *>(R + 3 * static_cast<size_t>(base));
      float v[3 * CH_GROUP];
#pragma unroll
      for (int t = 0; t < 3 * CH_GROUP / 4; ++t) {
        const float4 w = __ldg(p4 + t);
        v[4 * t] = w.x; v[4 * t + 1] = w.y; v[4 * t + 2] = w.z; v[4 * t + 3] = w.w;
      }
#pragma unroll
      for (int t = CH_GROUP - 1; t >= 0; --t) {
        const float d = dist_yxz(__fsub_rn(v[3 * t], qx[s].x), __fsub_rn(v[3 * t + 1], qy[s].x), __fsub_rn(v[3 * t + 2], qz[s].x));
        found = (d == best[s]) ? base + t : found;
      }
    } else {
#pragma unroll
      for (int t = CH_GROUP - 1; t >= 0; --t) {
        const int j = base + t;
        if (j < nr) {
          const float bx = __ldg(R + 3 * j), by = __ldg(R + 3 * j + 1), bz = __ldg(R + 3 * j + 2);
          const float d = dist_yxz(__fsub_rn(bx, qx[s].x), __fsub_rn(by, qy[s].x), __fsub_rn(bz, qz[s].x));
          found = (d == best[s]) ? j : found;
        }
      }
    }
    myidx[s] = found;
  }

#pragma unroll
  for (int s = 0; s < QT; ++s) {
    const int q = qbase + s * 32 + lane;
    if (q < nq) {
      const size_t o = cloud * nq + q;
      if (keys) {
        const uint64_t key = pack_key(best[s], static_cast<uint32_t>(myidx[s] + ref_offset));
        // column-split unit: the (distance, index) order of the key is the reference's tie rule
        if (merge) atomicMin(reinterpret_cast<unsigned long long *>(keys + o), key);
        else keys[o] = key;
      } else {
        dist[o] = best[s];
        idx[o] = myidx[s];
      }
    }
  }
}

template <int QT, int THREADS, int MINB, bool SYM, int CH_TILE = 512 /* reference points per shared-memory tile */,
          int STEP = 8 /* reference points per inner step (4 or 8): QT*STEP distances are live at once */,
          bool PACKED = true /* fp32x2 arithmetic */>
__global__ void __launch_bounds__(THREADS, MINB) chamfer_min_kernel(const ChamferDir d0, const ChamferDir d1) {
  constexpr int W = THREADS / 32;
  __shared__ __align__(16) float tile[2][3][CH_TILE];
  __shared__ __align__(16) unsigned colmin[SYM ? 2 : 1][W][SYM ? CH_TILE : 4];

  // column split (symmetric forward only, d1 absent): consecutive blocks are the chunks of one row block
  const int nc = d0.csplit > 1 ? d0.csplit : 1;
  const int per_cloud = (d0.qtiles + d1.qtiles) * nc;
  const int cloud = blockIdx.x / per_cloud;
  int t = blockIdx.x - cloud * per_cloud;
  const int chunk = t % nc;
  t /= nc;
  const bool second = t >= d0.qtiles;
  if (second) t -= d0.qtiles;
  const ChamferDir &d = second ? d1 : d0;
  int tile_begin = 0, tile_end = 0x7fffffff;
  if (nc > 1) {
    const int ntiles = (d.nr + CH_TILE - 1) / CH_TILE, tpc = (ntiles + nc - 1) / nc;
    tile_begin = chunk * tpc;
    tile_end = tile_begin + tpc;
    if (tile_begin >= ntiles) return;  // uniform for the CTA
  }
  chamfer_min_body<QT, THREADS, SYM, CH_TILE, STEP, false, PACKED>(
      d.q + static_cast<size_t>(cloud) * d.nq * 3, d.r + static_cast<size_t>(cloud) * d.nr * 3, d.nq, d.nr,
      t * (QT * THREADS), static_cast<size_t>(cloud), d.dist, d.idx, d.keys, d.ref_offset, d.colkeys, tile, colmin,
      tile_begin, tile_end, nc > 1);
}

// Balanced variant of the symmetric forward (opt-in, PDAE_CHAMFER_CFG=17; measured NOT faster, kept as the record of
// the experiment).  The uniform kernel above hands every CTA
// 128*QT rows; at 128 x 2048^2 that is 512 equal CTAs for 296 resident slots, and since ONE 4-warp CTA already
// saturates an SM's FMA issue the SMs that receive three CTAs instead of four idle for the last eighth of the kernel
// (FMA pipe 67 % of active but 57 % of elapsed cycles).  Here the rows of the whole batch are cut into 128-row
// slices, the slices are dealt out evenly to a persistent grid (two CTAs per SM), and every CTA walks its share in
// passes of up to four slices of one cloud (QT = 4, 3, 2 or 1 queries per thread): the per-scheduler work differs by
// at most one slice (7 against 6.92 at the headline shape).  Warps name themselves in the column keys by
// (first row / 32) << 3 | QT because passes of different QT give 32*QT-row groups.
// Result on B200 (profiles/r01/probe_chamfer_balanced.csv): 161.9 us against 161.1 us for the uniform kernel under ncu.
// Cloud boundaries force passes narrower than four slices (at 16 slices per cloud and ~7 per CTA: 352 passes of QT 3,
// 208 of QT 4, 64 of QT 2, 32 of QT 1), and a narrow pass pays the per-column work (tile reads, column reductions,
// REDUX, key posts) for fewer rows: the in-loop FMA rate drops from 67 % to 61.5 % and the SMs are no better balanced
// (active cycles min/avg/max 242k/279k/297k).
template <int THREADS, int CH_TILE, int STEP>
__global__ void __launch_bounds__(THREADS, 1) chamfer_min_balanced_kernel(const ChamferDir d, int slices_per_cloud,
                                                                          long long total_slices) {
  constexpr int W = THREADS / 32;
  __shared__ __align__(16) float tile[2][3][CH_TILE];
  __shared__ __align__(16) unsigned colmin[2][W][CH_TILE];
  long long cur = static_cast<long long>(blockIdx.x) * total_slices / gridDim.x;
  const long long end = static_cast<long long>(blockIdx.x + 1) * total_slices / gridDim.x;
  while (cur < end) {
    const long long cloud = cur / slices_per_cloud;
    const int s0 = static_cast<int>(cur - cloud * slices_per_cloud);
    long long seg = end - cur;  // what is left of this CTA's share inside the current cloud ...
    if (seg > slices_per_cloud - s0) seg = slices_per_cloud - s0;
    const long long passes = (seg + 3) / 4;  // ... walked in as few passes as possible, of (nearly) equal width:
    const long long cnt = (seg + passes - 1) / passes;  // 5 -> 3+2, 6 -> 3+3, 7 -> 4+3 (narrow passes run at a lower FMA rate)
    const float *Q = d.q + static_cast<size_t>(cloud) * d.nq * 3, *R = d.r + static_cast<size_t>(cloud) * d.nr * 3;
    const int row_base = s0 * THREADS;
    __syncthreads();  // the previous pass is done with the shared buffers
#define PDAE_PASS(QT)                                                                                                   \
  chamfer_min_body<QT, THREADS, true, CH_TILE, STEP, true>(Q, R, d.nq, d.nr, row_base, static_cast<size_t>(cloud), d.dist, \
                                                          d.idx, d.keys, d.ref_offset, d.colkeys, tile, colmin)
    if (cnt == 4) PDAE_PASS(4);
    else if (cnt == 3) PDAE_PASS(3);
    else if (cnt == 2) PDAE_PASS(2);
    else PDAE_PASS(1);
#undef PDAE_PASS
    cur += cnt;
  }
}

// ---- tiny clouds (both sides <= SMALL_MAX points): one warp per (cloud, direction) -----------
// The reference spends a (32,16)x512 grid on these; Point-DAE's fine loss runs ~5000 clouds of
// 32..36 points per step (models/PointCAE_transformer.py:1066).  Mirrors the reference's
// `k == 0 || d < best` scan literally (including its NaN behaviour).
constexpr int SMALL_MAX = 128;
constexpr int SMALL_WARPS = 8;

__global__ void __launch_bounds__(SMALL_WARPS * 32) chamfer_small_kernel(const float *__restrict__ xyz1,
                                                                         const float *__restrict__ xyz2, int b, int n,
                                                                         int m, float *__restrict__ dist1,
                                                                         float *__restrict__ dist2, int *__restrict__ idx1,
                                                                         int *__restrict__ idx2) {
  __shared__ float sref[SMALL_WARPS][3 * SMALL_MAX];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long gw = static_cast<long long>(blockIdx.x) * SMALL_WARPS + warp;
  if (gw >= 2LL * b) return;
  const int cloud = static_cast<int>(gw >> 1);
  const bool second = gw & 1;
  const int nq = second ? m : n, nr = second ? n : m;
  const float *__restrict__ Q = (second ? xyz2 : xyz1) + static_cast<size_t>(cloud) * nq * 3;
  const float *__restrict__ R = (second ? xyz1 : xyz2) + static_cast<size_t>(cloud) * nr * 3;
  float *dist = (second ? dist2 : dist1) + static_cast<size_t>(cloud) * nq;
  int *idx = (second ? idx2 : idx1) + static_cast<size_t>(cloud) * nq;
  float *s = sref[warp];
  for (int i = lane; i < nr * 3; i += 32) s[i] = __ldg(R + i);
  __syncwarp();
  for (int q = lane; q < nq; q += 32) {
    const float x1 = __ldg(Q + 3 * q), y1 = __ldg(Q + 3 * q + 1), z1 = __ldg(Q + 3 * q + 2);
    float best = 0.0f;
    int besti = 0;
#pragma unroll 4
    for (int k = 0; k < nr; ++k) {
      const float d = dist_yxz(__fsub_rn(s[3 * k], x1), __fsub_rn(s[3 * k + 1], y1), __fsub_rn(s[3 * k + 2], z1));
      const bool take = (k == 0) || (d < best);
      best = take ? d : best;
      besti = take ? k : besti;
    }
    dist[q] = best;
    idx[q] = besti;
  }
}

// ---- backward: one thread per point of either cloud, scatter through the saved argmin ---------
// reference: chamfer.cu:173-201 runs (1,16)x256 = 4096 threads over the whole batch and adds BOTH terms of every
// pair with atomics into zero-filled buffers.  Every point receives exactly one "own" term (g * (a_j - b_idx[j]),
// written by its own thread) plus the scattered terms of the points that chose it, so the work is split in two
// launches over all b*(n+m) points: OWN stores the own term with a plain store (which also initialises the
// buffer: no memset), SCATTER then adds -v to the matched point with RED.ADD.F32.  Half the atomics of the
// reference scheme, accumulation order free as in the reference.
//
// GradSrc supplies d(loss)/d(dist) per point: the arrays handed to chamfer.backward, or -- for the fused mean
// losses (ChamferDistanceL2 / L1) -- the upstream scalar times a constant, with the sqrt derivative for L1.
struct GradArrays {
  const float *gd1, *gd2;
  __device__ __forceinline__ float at(bool second, size_t o) const { return __ldg((second ? gd2 : gd1) + o); }
};
struct GradMean {
  const float *gloss;          // upstream gradient of the scalar loss (device)
  const float *dist1, *dist2;  // squared distances of the forward (L1 only)
  float w1, w2;                // d(loss)/d(mean term) / count
  int l1;
  __device__ __forceinline__ float at(bool second, size_t o) const {
    const float g = __fmul_rn(__ldg(gloss), second ? w2 : w1);
    if (!l1) return g;
    // d sqrt(x) = g / (2 sqrt(x)); x == 0 gives inf like torch's SqrtBackward (and NaN after the multiply by 0)
    return __fdiv_rn(g, __fmul_rn(2.0f, __fsqrt_rn(__ldg((second ? dist2 : dist1) + o))));
  }
};

template <bool SCATTER, typename GradSrc>
__global__ void __launch_bounds__(256) chamfer_bwd_kernel(const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                                                          const int *__restrict__ idx1, const int *__restrict__ idx2,
                                                          const GradSrc gs, int n, int m, int blocks_per_cloud,
                                                          float *__restrict__ gx1, float *__restrict__ gx2) {
  const unsigned cloud_u = blockIdx.x / blocks_per_cloud;  // 1-D grid: no 65535-cloud limit (the fine loss has thousands)
  int i = (blockIdx.x - cloud_u * blocks_per_cloud) * blockDim.x + threadIdx.x;  // point of cloud 1 (i < n) or 2 (i - n < m)
  const size_t cloud = cloud_u;
  const bool second = i >= n;
  const float *A, *Bp;
  const int *idx;
  float *ga, *gb;
  size_t o;
  if (!second) {
    A = xyz1 + cloud * n * 3; Bp = xyz2 + cloud * m * 3; idx = idx1 + cloud * n; o = cloud * n + i;
    ga = gx1 + cloud * n * 3; gb = gx2 + cloud * m * 3;
  } else {
    i -= n;
    if (i >= m) return;
    A = xyz2 + cloud * m * 3; Bp = xyz1 + cloud * n * 3; idx = idx2 + cloud * m; o = cloud * m + i;
    ga = gx2 + cloud * m * 3; gb = gx1 + cloud * n * 3;
  }
  const int j2 = __ldg(idx + i);
  const float g = __fmul_rn(gs.at(second, o), 2.0f);
  float v[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) v[c] = __fmul_rn(g, __fsub_rn(__ldg(A + 3 * i + c), __ldg(Bp + 3 * j2 + c)));
  if (!SCATTER) {
#pragma unroll
    for (int c = 0; c < 3; ++c) ga[3 * i + c] = v[c];
  } else {
    // two reductions per point instead of three: a point's 12 bytes always hold one 8-byte aligned pair (x,y for an even
    // point of an 8-byte aligned cloud, y,z for an odd one) -> one RED.ADD.v2.f32 + one scalar RED.ADD.F32
    float *t = gb + 3 * j2;
    if ((reinterpret_cast<uintptr_t>(t) & 7u) == 0) {
      atomicAdd(reinterpret_cast<float2 *>(t), make_float2(-v[0], -v[1]));
      atomicAdd(t + 2, -v[2]);
    } else if ((reinterpret_cast<uintptr_t>(t + 1) & 7u) == 0) {
      atomicAdd(t, -v[0]);
      atomicAdd(reinterpret_cast<float2 *>(t + 1), make_float2(-v[1], -v[2]));
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) atomicAdd(t + c, -v[c]);
    }
  }
}

template <typename GradSrc>
static int launch_chamfer_bwd(const float *xyz1, const float *xyz2, const int *idx1, const int *idx2, const GradSrc &gs,
                              int b, int n, int m, float *gx1, float *gx2, cudaStream_t st) {
  const long long bpc = (static_cast<long long>(n) + m + 255) / 256;
  if (bpc * b > 0x7fffffffLL) return PDAE_E_UNSUPPORTED;
  const unsigned grid = static_cast<unsigned>(bpc * b);
  chamfer_bwd_kernel<false, GradSrc><<<grid, 256, 0, st>>>(xyz1, xyz2, idx1, idx2, gs, n, m, static_cast<int>(bpc), gx1, gx2);
  PDAE_RETURN_IF_LAUNCH_FAILED();
  chamfer_bwd_kernel<true, GradSrc><<<grid, 256, 0, st>>>(xyz1, xyz2, idx1, idx2, gs, n, m, static_cast<int>(bpc), gx1, gx2);
  PDAE_RETURN_IF_LAUNCH_FAILED();
  return 0;
}

// ---- fused mean losses: (mean f(dist1), mean f(dist2)) with f = identity (L2) or sqrt (L1) ------------------
// reference: extensions/chamfer_dist/__init__.py:43 (L2: mean + mean) and :413-417 (L1: (mean sqrt + mean sqrt) / 2),
// there two or four torch kernels plus the add.  Deterministic: fixed-size partial sums in fp32 per CTA (fp64 across
// the CTA partials), summed in a fixed order by the second launch.
constexpr int LOSS_BLOCKS = 128;

__global__ void __launch_bounds__(256) chamfer_loss_partial_kernel(const float *__restrict__ dist1, const float *__restrict__ dist2,
                                                                   long long c1, long long c2, int l1,
                                                                   float *__restrict__ partial /*[2][LOSS_BLOCKS]*/) {
  __shared__ float red[2][8];
  float acc[2] = {0.f, 0.f};
#pragma unroll
  for (int side = 0; side < 2; ++side) {
    const float *__restrict__ d = side ? dist2 : dist1;
    const long long cnt = side ? c2 : c1;
    // four strided loads in flight per round (the sum is latency-bound: 8 values per thread at the headline shape)
    constexpr long long STRIDE = 256LL * LOSS_BLOCKS;
    for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < cnt; i += 4 * STRIDE) {
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = i + u * STRIDE < cnt ? __ldg(d + i + u * STRIDE) : 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[side] += l1 ? __fsqrt_rn(v[u]) : v[u];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[side] += __shfl_xor_sync(0xffffffffu, acc[side], o);
    if ((threadIdx.x & 31) == 0) red[side][threadIdx.x >> 5] = acc[side];
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
    partial[threadIdx.x * LOSS_BLOCKS + blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(32) chamfer_loss_final_kernel(const float *__restrict__ partial, long long c1, long long c2,
                                                                int l1, float *__restrict__ out /*[3]*/) {
  // lanes 0-15 sum the first term's partials, lanes 16-31 the second's: 8 each, then a fixed shuffle tree (deterministic)
  const int side = threadIdx.x >> 4, l = threadIdx.x & 15;
  double t = 0.0;
#pragma unroll
  for (int i = 0; i < LOSS_BLOCKS / 16; ++i) t += static_cast<double>(partial[side * LOSS_BLOCKS + l * (LOSS_BLOCKS / 16) + i]);
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  const long long cnt = side ? c2 : c1;
  const float mean = static_cast<float>(t / static_cast<double>(cnt));  // 0/0 = NaN like torch.mean of an empty tensor
  const float other = __shfl_xor_sync(0xffffffffu, mean, 16);
  if (l == 0) out[1 + side] = mean;
  if (threadIdx.x == 0) out[0] = l1 ? __fmul_rn(__fadd_rn(mean, other), 0.5f) : __fadd_rn(mean, other);
}

__global__ void __launch_bounds__(256) unpack_keys_kernel(const uint64_t *__restrict__ keys, long long count,
                                                          float *__restrict__ dist, int *__restrict__ idx) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint64_t k = keys[i];
  dist[i] = __uint_as_float(static_cast<uint32_t>(k >> 32));
  idx[i] = static_cast<int>(static_cast<uint32_t>(k));
}

// tuning hook: PDAE_CHAMFER_CFG selects the CTA shape of the large-cloud kernel
// SMs of the current device (cached per device index; the grid of the balanced kernel is two CTAs per SM)
static int sm_count() {
  static int cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    cached[dev] = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
  }
  return cached[dev];
}

static int g_chamfer_variant = -1;
static int chamfer_variant() {
  if (g_chamfer_variant < 0) {
    const char *e = getenv("PDAE_CHAMFER_CFG");
    g_chamfer_variant = e ? atoi(e) : 0;
  }
  return g_chamfer_variant;
}
// tuning hook (profiles/tune_kernels.py): switch the kernel shape inside one process; ids as PDAE_CHAMFER_CFG
extern "C" int pdae_tune_chamfer_variant(int v) {
  const int old = chamfer_variant();
  if (v >= 0) g_chamfer_variant = v;
  return old;
}
// queries per CTA of the kernel launch_min / launch_sym will pick (the callers size `qtiles` with it)
static int chamfer_qpc(int nq_max, bool sym) {
  if (!sym && nq_max <= 256) return 128;
  switch (chamfer_variant() % 25) {
    case 3: return 512;   // <2,256>
    case 4: return 256;   // <2,128>
    case 5: case 6: case 7: return 256;   // <4,64>
    case 8: case 10: return 512;   // <8,64>
    case 9: case 11: return 256;   // <8,32>
    case 12: case 13: return 128;  // <4,32>
    case 14: case 15: return 1024; // <8,128>
    default: return 512;  // <4,128>
  }
}

template <bool SYM>
static int launch_min(const ChamferDir &d0, const ChamferDir &d1, int b, cudaStream_t st) {
  const long long per_cloud = (static_cast<long long>(d0.qtiles) + d1.qtiles) * (d0.csplit > 1 ? d0.csplit : 1);
  const long long grid = per_cloud * b;
  if (grid <= 0) return 0;
  if (grid > 0x7fffffffLL) return PDAE_E_UNSUPPORTED;
  const unsigned g = static_cast<unsigned>(grid);
  const int nq_max = d0.nq > d1.nq ? d0.nq : d1.nq;
  if (!SYM && nq_max <= 256) {
    chamfer_min_kernel<1, 128, 1, false><<<g, 128, 0, st>>>(d0, d1);
  } else {
    switch (chamfer_variant() % 25) {
      case 1: chamfer_min_kernel<4, 128, 4, SYM><<<g, 128, 0, st>>>(d0, d1); break;
      case 2: chamfer_min_kernel<4, 128, 3, SYM><<<g, 128, 0, st>>>(d0, d1); break;
      case 3: chamfer_min_kernel<2, 256, 3, SYM><<<g, 256, 0, st>>>(d0, d1); break;
      case 4: chamfer_min_kernel<2, 128, 6, SYM><<<g, 128, 0, st>>>(d0, d1); break;
      case 5: chamfer_min_kernel<4, 64, 4, SYM, 256><<<g, 64, 0, st>>>(d0, d1); break;
      case 6: chamfer_min_kernel<4, 64, 8, SYM, 256><<<g, 64, 0, st>>>(d0, d1); break;
      case 7: chamfer_min_kernel<4, 64, 6, SYM, 256><<<g, 64, 0, st>>>(d0, d1); break;
      case 8: chamfer_min_kernel<8, 64, 4, SYM, 256, 4><<<g, 64, 0, st>>>(d0, d1); break;
      case 9: chamfer_min_kernel<8, 32, 8, SYM, 128, 4><<<g, 32, 0, st>>>(d0, d1); break;
      case 10: chamfer_min_kernel<8, 64, 6, SYM, 256, 4><<<g, 64, 0, st>>>(d0, d1); break;
      case 11: chamfer_min_kernel<8, 32, 12, SYM, 128, 4><<<g, 32, 0, st>>>(d0, d1); break;
      case 12: chamfer_min_kernel<4, 32, 16, SYM, 128, 8><<<g, 32, 0, st>>>(d0, d1); break;
      case 13: chamfer_min_kernel<4, 32, 12, SYM, 128, 8><<<g, 32, 0, st>>>(d0, d1); break;
      case 14: chamfer_min_kernel<8, 128, 1, SYM, 512, 4><<<g, 128, 0, st>>>(d0, d1); break;
      case 15: chamfer_min_kernel<8, 128, 1, SYM, 512, 8><<<g, 128, 0, st>>>(d0, d1); break;
      case 18: chamfer_min_kernel<4, 128, 1, SYM, 512, 8, false><<<g, 128, 0, st>>>(d0, d1); break;
      default: chamfer_min_kernel<4, 128, 1, SYM><<<g, 128, 0, st>>>(d0, d1); break;
    }
  }
  PDAE_RETURN_IF_LAUNCH_FAILED();
  return 0;
}

// identity of the MIN reduction over packed keys: larger than every real key both as uint64 and as
// int64 (torch / NCCL reduce the keys as signed 64-bit; real keys have a clear top bit because d >= 0).
constexpr uint64_t CHAMFER_KEY_IDENTITY = 0x7fffffffffffffffull;

__global__ void __launch_bounds__(256) fill_keys_kernel(uint64_t *__restrict__ keys, long long count) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < count) keys[i] = CHAMFER_KEY_IDENTITY;
}

// symmetric mode, second half: one warp per column point b_j.  colkeys[j] = (min_i d(a_i, b_j), id of the
// QPG-query group holding a minimiser; lowest such group).  Scanning that group in index order and taking
// the first query whose distance equals the minimum bit for bit yields the lowest index overall.
// QPG = queries per group (32 * QT of the main kernel).  A group's queries are contiguous, so each lane takes
// QPG/32 consecutive ones (float4 loads when the cloud base is 16-byte aligned) and the whole group is examined in
// one step; lanes are in index order, hence ballot + ffs gives the lowest index.  A warp resolves RPW consecutive
// column points with all their loads issued up front: the kernel is pure load latency (key -> group rows), so the
// memory-level parallelism per warp is what sets its speed.
template <int QPG, int RPW>
__global__ void __launch_bounds__(256) chamfer_col_recover_kernel(const float *__restrict__ rows, const float *__restrict__ cols,
                                                                  const uint64_t *__restrict__ colkeys, int n_rows,
                                                                  int n_cols,
                                                                  float *__restrict__ dist, int *__restrict__ idx) {
  constexpr int PER = QPG / 32;  // consecutive queries per lane
  const int lane = threadIdx.x & 31;
  const int j0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * RPW;  // first column point of the warp
  if (j0 >= n_cols) return;
  const size_t cloud = blockIdx.y;
  const float *__restrict__ A = rows + cloud * n_rows * 3;
  const bool vec = (n_rows & 3) == 0 && PER % 4 == 0;
  uint64_t key[RPW];
  float bx[RPW], by[RPW], bz[RPW];
#pragma unroll
  for (int u = 0; u < RPW; ++u) {
    const int j = j0 + u < n_cols ? j0 + u : n_cols - 1;
    const size_t gw = cloud * n_cols + j;
    key[u] = colkeys[gw];
    bx[u] = __ldg(cols + 3 * gw); by[u] = __ldg(cols + 3 * gw + 1); bz[u] = __ldg(cols + 3 * gw + 2);
  }
  float c[RPW][3 * PER];
  int base[RPW];
#pragma unroll
  for (int u = 0; u < RPW; ++u) {
    base[u] = static_cast<int>(static_cast<uint32_t>(key[u])) * QPG + lane * PER;
    if (vec && base[u] + PER <= n_rows) {
      const float4 *p4 = reinterpret_cast<const float4 *>(A + 3 * static_cast<size_t>(base[u]));
#pragma unroll
      for (int t = 0; t < 3 * PER / 4; ++t) {
        const float4 w = __ldg(p4 + t);
        c[u][4 * t] = w.x; c[u][4 * t + 1] = w.y; c[u][4 * t + 2] = w.z; c[u][4 * t + 3] = w.w;
      }
    } else {
#pragma unroll
      for (int t = 0; t < 3 * PER; ++t)
        c[u][t] = (base[u] + t / 3 < n_rows) ? __ldg(A + 3 * static_cast<size_t>(base[u]) + t) : __int_as_float(0x7fc00000);
    }
  }
#pragma unroll
  for (int u = 0; u < RPW; ++u) {
    const float v = __uint_as_float(static_cast<uint32_t>(key[u] >> 32));
    int first = PER;  // first matching query of this lane
#pragma unroll
    for (int t = PER - 1; t >= 0; --t) {
      const float d = dist_yxz(__fsub_rn(bx[u], c[u][3 * t]), __fsub_rn(by[u], c[u][3 * t + 1]), __fsub_rn(bz[u], c[u][3 * t + 2]));
      first = (d == v) ? t : first;  // NaN padding never matches
    }
    const unsigned mk = __ballot_sync(0xffffffffu, first < PER);
    const int src = mk ? __ffs(mk) - 1 : 0;
    const int found = __shfl_sync(0xffffffffu, base[u] + first, src);
    if (lane == 0 && j0 + u < n_cols) {
      const size_t gw = cloud * n_cols + j0 + u;
      dist[gw] = v;
      idx[gw] = mk ? found : 0;
    }
  }
}

// Group-major variant of the recovery for clouds up to a few thousand points: one CTA per (cloud, QPG-row group).
// Every lane keeps QPG/32 consecutive rows of the group in registers for the whole CTA; the warps sweep the cloud's
// column keys (coalesced), and only the columns whose key names this group are resolved, their coordinates and
// minimum broadcast from the lane that read them.  The warp-per-column kernel above re-reads a whole group
// (12*QPG bytes) per column -- 400 MB of L2 traffic at 128 x 2048^2; here the traffic is the keys once per group
// (8*n_cols bytes per CTA), 10x less, and the rows never leave registers.  Traffic grows with n_rows*n_cols/QPG, so
// the launcher keeps the warp-per-column kernel for large clouds.
template <int QPG>
__global__ void __launch_bounds__(256) chamfer_col_recover_grouped_kernel(const float *__restrict__ rows,
                                                                          const float *__restrict__ cols,
                                                                          const uint64_t *__restrict__ colkeys,
                                                                          int n_rows, int n_cols,
                                                                          float *__restrict__ dist, int *__restrict__ idx) {
  constexpr int PER = QPG / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned g = blockIdx.x;
  const size_t cloud = blockIdx.y;
  const float *__restrict__ A = rows + cloud * n_rows * 3;
  const int base = static_cast<int>(g) * QPG + lane * PER;
  float rx[PER], ry[PER], rz[PER];
  if ((n_rows & 3) == 0 && PER % 4 == 0 && base + PER <= n_rows) {
    const float4 *p4 = reinterpret_cast<const float4 *>(A + 3 * static_cast<size_t>(base));
    float c[3 * PER];
#pragma unroll
    for (int t = 0; t < 3 * PER / 4; ++t) {
      const float4 w = __ldg(p4 + t);
      c[4 * t] = w.x; c[4 * t + 1] = w.y; c[4 * t + 2] = w.z; c[4 * t + 3] = w.w;
    }
#pragma unroll
    for (int t = 0; t < PER; ++t) rx[t] = c[3 * t], ry[t] = c[3 * t + 1], rz[t] = c[3 * t + 2];
  } else {
#pragma unroll
    for (int t = 0; t < PER; ++t) {
      const bool in = base + t < n_rows;  // NaN padding never matches
      rx[t] = in ? __ldg(A + 3 * static_cast<size_t>(base + t)) : __int_as_float(0x7fc00000);
      ry[t] = in ? __ldg(A + 3 * static_cast<size_t>(base + t) + 1) : 0.0f;
      rz[t] = in ? __ldg(A + 3 * static_cast<size_t>(base + t) + 2) : 0.0f;
    }
  }
  const uint64_t *__restrict__ K = colkeys + cloud * n_cols;
  const float *__restrict__ C = cols + cloud * n_cols * 3;
  constexpr int UNR = 4;  // key sweeps in flight per warp: the loop is pure L2 latency otherwise
  for (int j0 = warp * 32; j0 < n_cols; j0 += 256 * UNR) {
    uint64_t key[UNR];
    float cx[UNR], cy[UNR], cz[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {  // keys and coordinates fetched together (coalesced, no dependent second trip)
      const int j = j0 + u * 256 + lane;
      const bool in = j < n_cols;
      key[u] = in ? K[j] : ~0ull;
      cx[u] = in ? __ldg(C + 3 * j) : 0.f, cy[u] = in ? __ldg(C + 3 * j + 1) : 0.f, cz[u] = in ? __ldg(C + 3 * j + 2) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int j = j0 + u * 256 + lane;
      const bool match = static_cast<uint32_t>(key[u]) == g && j < n_cols;
      const float v = __uint_as_float(static_cast<uint32_t>(key[u] >> 32));
      unsigned mk = __ballot_sync(0xffffffffu, match);
      int mine = 0;
      while (mk) {
        const int src = __ffs(mk) - 1;
        mk &= mk - 1;
        const float bx = __shfl_sync(0xffffffffu, cx[u], src), by = __shfl_sync(0xffffffffu, cy[u], src);
        const float bz = __shfl_sync(0xffffffffu, cz[u], src), want = __shfl_sync(0xffffffffu, v, src);
        int first = PER;  // first matching row of this lane
#pragma unroll
        for (int t = PER - 1; t >= 0; --t) {
          const float d = dist_yxz(__fsub_rn(bx, rx[t]), __fsub_rn(by, ry[t]), __fsub_rn(bz, rz[t]));
          first = (d == want) ? t : first;
        }
        const unsigned hit = __ballot_sync(0xffffffffu, first < PER);
        const int found = __shfl_sync(0xffffffffu, base + first, hit ? __ffs(hit) - 1 : 0);
        if (lane == src) mine = hit ? found : 0;
      }
      if (match) {
        dist[cloud * n_cols + j] = v;
        idx[cloud * n_cols + j] = mine;
      }
    }
  }
}

// List variant of the group-major recovery (default for clouds up to 32 groups): the CTA of (cloud, group) stages the
// group's QPG rows in shared memory as planes, every thread tests a strided share of the cloud's column keys and
// appends the columns won by this group to a shared list, and then ONE THREAD PER LISTED COLUMN walks the group's rows
// (broadcast LDS.64 of two rows per plane, packed distance, compare with the recorded minimum bit for bit), back to
// front so that the lowest matching row is the one kept.  No shuffles or ballots in the walk: ~6.5 instructions per
// (column, row) against ~14 in the warp-cooperative kernel above.
constexpr int RECOVER_LIST_MAX = 4096;  // columns examined per pass (list of int in shared memory)

template <int QPG>
__global__ void __launch_bounds__(256) chamfer_col_recover_list_kernel(const float *__restrict__ rows,
                                                                       const float *__restrict__ cols,
                                                                       const uint64_t *__restrict__ colkeys, int n_rows,
                                                                       int n_cols, float *__restrict__ dist,
                                                                       int *__restrict__ idx,
                                                                       const uint64_t *__restrict__ rowkeys,
                                                                       float *__restrict__ drow, int *__restrict__ irow) {
  __shared__ __align__(16) float sx[QPG], sy[QPG], sz[QPG];
  __shared__ int list[RECOVER_LIST_MAX];
  __shared__ int count;
  const unsigned g = blockIdx.x;
  const size_t cloud = blockIdx.y;
  const float *__restrict__ A = rows + cloud * n_rows * 3;
  const uint64_t *__restrict__ K = colkeys + cloud * n_cols;
  const float *__restrict__ C = cols + cloud * n_cols * 3;
  const int tid = threadIdx.x;
  if (tid == 0) count = 0;
  if (rowkeys != nullptr) {  // column-split forward: this CTA's rows leave their merged keys here (no extra launch)
    for (int r = tid; r < QPG; r += 256) {
      const long long row = static_cast<long long>(g) * QPG + r;
      if (row < n_rows) {
        const uint64_t key = rowkeys[cloud * n_rows + row];
        drow[cloud * n_rows + row] = __uint_as_float(static_cast<uint32_t>(key >> 32));
        irow[cloud * n_rows + row] = static_cast<int>(static_cast<uint32_t>(key));
      }
    }
  }
  for (int e = tid; e < 3 * QPG; e += 256) {  // coalesced AoS read of the group's rows; NaN padding never matches
    const int r = e / 3, c = e - 3 * r;
    const long long row = static_cast<long long>(g) * QPG + r;
    const float v = row < n_rows ? __ldg(A + row * 3 + c) : __int_as_float(0x7fc00000);
    (c == 0 ? sx : c == 1 ? sy : sz)[r] = v;
  }
  __syncthreads();
  constexpr int UNR = 8;
  for (int c0 = 0; c0 < n_cols; c0 += RECOVER_LIST_MAX) {  // the list holds one chunk of columns at a time
    const int c1 = c0 + RECOVER_LIST_MAX < n_cols ? c0 + RECOVER_LIST_MAX : n_cols;
    for (int j0 = c0 + tid; j0 < c1; j0 += 256 * UNR) {
      uint32_t gid[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int j = j0 + u * 256;
        gid[u] = j < c1 ? static_cast<uint32_t>(K[j]) : 0xffffffffu;
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u)
        if (gid[u] == g) list[atomicAdd(&count, 1)] = j0 + u * 256;
    }
    __syncthreads();
    const int total = count;
    for (int m = tid; m < total; m += 256) {
      const int j = list[m];
      const float want = __uint_as_float(static_cast<uint32_t>(K[j] >> 32));
      const float bx = __ldg(C + 3 * j), by = __ldg(C + 3 * j + 1), bz = __ldg(C + 3 * j + 2);
      const float2 cx = make_float2(bx, bx), cy = make_float2(by, by), cz = make_float2(bz, bz);
      int found = 0;
#pragma unroll 8
      for (int r = QPG - 2; r >= 0; r -= 2) {
        const float2 d = dist_yxz2(sub2(cx, *reinterpret_cast<const float2 *>(sx + r)),
                                   sub2(cy, *reinterpret_cast<const float2 *>(sy + r)),
                                   sub2(cz, *reinterpret_cast<const float2 *>(sz + r)));
        found = (d.y == want) ? r + 1 : found;
        found = (d.x == want) ? r : found;
      }
      dist[cloud * n_cols + j] = want;
      idx[cloud * n_cols + j] = static_cast<int>(g) * QPG + found;
    }
    __syncthreads();
    if (tid == 0) count = 0;
    __syncthreads();
  }
}

// Recovery for the balanced kernel: its column keys name the winning warp's rows as (first row / 32) << 3 | QT, a
// 32-aligned range of 32*QT <= 128 rows.  One CTA per (cloud, aligned 128-row block g) stages rows
// [128 g, 128 g + 224) -- a range that starts in the block ends before that -- collects the columns whose range starts
// in the block and walks exactly the range, back to front (lowest matching row kept), as the list kernel above.
__global__ void __launch_bounds__(256) chamfer_col_recover_var_kernel(const float *__restrict__ rows,
                                                                      const float *__restrict__ cols,
                                                                      const uint64_t *__restrict__ colkeys, int n_rows,
                                                                      int n_cols, float *__restrict__ dist,
                                                                      int *__restrict__ idx) {
  constexpr int SPAN = 224;
  __shared__ __align__(16) float sx[SPAN], sy[SPAN], sz[SPAN];
  __shared__ int list[RECOVER_LIST_MAX];
  __shared__ int count;
  const unsigned g = blockIdx.x;
  const size_t cloud = blockIdx.y;
  const float *__restrict__ A = rows + cloud * n_rows * 3;
  const uint64_t *__restrict__ K = colkeys + cloud * n_cols;
  const float *__restrict__ C = cols + cloud * n_cols * 3;
  const int tid = threadIdx.x;
  if (tid == 0) count = 0;
  for (int e = tid; e < 3 * SPAN; e += 256) {  // NaN padding (rows past the cloud) never matches
    const int r = e / 3, c = e - 3 * r;
    const long long row = static_cast<long long>(g) * 128 + r;
    const float v = row < n_rows ? __ldg(A + row * 3 + c) : __int_as_float(0x7fc00000);
    (c == 0 ? sx : c == 1 ? sy : sz)[r] = v;
  }
  __syncthreads();
  constexpr int UNR = 8;
  for (int c0 = 0; c0 < n_cols; c0 += RECOVER_LIST_MAX) {
    const int c1 = c0 + RECOVER_LIST_MAX < n_cols ? c0 + RECOVER_LIST_MAX : n_cols;
    for (int j0 = c0 + tid; j0 < c1; j0 += 256 * UNR) {
      uint32_t who[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int j = j0 + u * 256;
        who[u] = j < c1 ? static_cast<uint32_t>(K[j]) : 0xffffffffu;
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u)
        if ((who[u] >> 5) == g) list[atomicAdd(&count, 1)] = j0 + u * 256;  // (first row / 32) / 4 == block
    }
    __syncthreads();
    const int total = count;
    for (int m = tid; m < total; m += 256) {
      const int j = list[m];
      const uint64_t key = K[j];
      const float want = __uint_as_float(static_cast<uint32_t>(key >> 32));
      const int lo = static_cast<int>((static_cast<uint32_t>(key) >> 3) & 3u) * 32;  // range start inside the block
      const int len = static_cast<int>(static_cast<uint32_t>(key) & 7u) * 32;
      const float bx = __ldg(C + 3 * j), by = __ldg(C + 3 * j + 1), bz = __ldg(C + 3 * j + 2);
      const float2 cx = make_float2(bx, bx), cy = make_float2(by, by), cz = make_float2(bz, bz);
      int found = lo;
#pragma unroll 8
      for (int r = lo + len - 2; r >= lo; r -= 2) {
        const float2 d = dist_yxz2(sub2(cx, *reinterpret_cast<const float2 *>(sx + r)),
                                   sub2(cy, *reinterpret_cast<const float2 *>(sy + r)),
                                   sub2(cz, *reinterpret_cast<const float2 *>(sz + r)));
        found = (d.y == want) ? r + 1 : found;
        found = (d.x == want) ? r : found;
      }
      dist[cloud * n_cols + j] = want;
      idx[cloud * n_cols + j] = static_cast<int>(g) * 128 + found;
    }
    __syncthreads();
    if (tid == 0) count = 0;
    __syncthreads();
  }
}

// second half of the symmetric forward: picks the recovery kernel by cloud size (see the comment above)
template <int QPG>
static int launch_col_recover(const float *rows, const float *cols, const uint64_t *ck, int b, int n_rows, int n_cols,
                              float *dcol, int *icol, cudaStream_t st, const uint64_t *rowkeys = nullptr,
                              float *drow = nullptr, int *irow = nullptr) {
  const long long groups = (static_cast<long long>(n_rows) + QPG - 1) / QPG;
  // group-major recovery while its key sweeps (8*groups B per column) stay cheaper than re-reading a group per column
  const int gdefault = chamfer_variant() < 25 ? 128 : 32;
  const int glimit = chamfer_variant() >= 50 ? 0 : (getenv("PDAE_RECOVER_GROUPS") ? atoi(getenv("PDAE_RECOVER_GROUPS")) : gdefault);
  if (groups <= glimit) {  // key + coordinate sweeps (20*groups B per column) cheaper than row re-reads (12*QPG B)
    const dim3 ggrid(static_cast<unsigned>(groups), b);
    if (chamfer_variant() < 25) {
      chamfer_col_recover_list_kernel<QPG><<<ggrid, 256, 0, st>>>(rows, cols, ck, n_rows, n_cols, dcol, icol, rowkeys, drow, irow);
    } else {
      if (rowkeys) return PDAE_E_INVALID;  // only the list kernel unpacks row keys (recover_unpacks_rows)
      chamfer_col_recover_grouped_kernel<QPG><<<ggrid, 256, 0, st>>>(rows, cols, ck, n_rows, n_cols, dcol, icol);
    }
  } else {
    if (rowkeys) return PDAE_E_INVALID;
    // 4 column points per warp, 8 warps per CTA (splitting a warp across points was measured slower)
    const dim3 rgrid(static_cast<unsigned>((n_cols + 31) / 32), b);
    chamfer_col_recover_kernel<QPG, 4><<<rgrid, 256, 0, st>>>(rows, cols, ck, n_rows, n_cols, dcol, icol);
  }
  PDAE_RETURN_IF_LAUNCH_FAILED();
  return 0;
}

// true when the recovery of an (unphased) column-split forward can unpack the row keys itself: default kernel shape
// (the only one that splits) and few enough 128-row groups for the list kernel, which visits every row exactly once
static bool recover_unpacks_rows(int n_rows) {
  const int glimit = getenv("PDAE_RECOVER_GROUPS") ? atoi(getenv("PDAE_RECOVER_GROUPS")) : 128;
  return chamfer_variant() == 0 && (static_cast<long long>(n_rows) + 127) / 128 <= glimit;
}

static int launch_col_recover_for_variant(const float *rows, const float *cols, const uint64_t *ck, int b, int n_rows,
                                          int n_cols, float *dcol, int *icol, cudaStream_t st,
                                          const uint64_t *rowkeys = nullptr, float *drow = nullptr, int *irow = nullptr) {
  const int v = chamfer_variant() % 25;  // queries per warp = 32 * QT of the variant launched
  if (v == 3 || v == 4) return launch_col_recover<64>(rows, cols, ck, b, n_rows, n_cols, dcol, icol, st);
  if ((v >= 8 && v <= 11) || v == 14 || v == 15) return launch_col_recover<256>(rows, cols, ck, b, n_rows, n_cols, dcol, icol, st);
  return launch_col_recover<128>(rows, cols, ck, b, n_rows, n_cols, dcol, icol, st, rowkeys, drow, irow);
}

}  // namespace pdae

using namespace pdae;

// column keys of the smaller cloud + (for column-split units) row keys of the larger one.  A workspace of only the
// first part (b * min(n, m) keys, the size this function returned before the split existed) is still accepted: the
// forward then runs unsplit.
extern "C" size_t pdae_chamfer_fwd_workspace_bytes(int b, int n, int m) {
  if (b <= 0 || n <= 0 || m <= 0) return 0;
  if (n <= SMALL_MAX && m <= SMALL_MAX) return 0;
  return static_cast<size_t>(b) * (static_cast<size_t>(n) + m) * sizeof(uint64_t);
}

// Column chunks per 512-row block for the symmetric forward.  One 4-warp CTA saturates an SM's FMA issue, so an SM
// works through its CTAs at a fixed rate and the kernel ends when the SM with the most CTAs ends: with U equal units
// on S SMs that is ceil(U / S) unit times.  512 row blocks on 148 SMs (128 x 2048^2) make 4 against an average of
// 3.46; cutting every block's column sweep in two makes 7 half-units against 6.92.  A unit costs its tiles plus a
// fixed prologue / epilogue (first tile not overlapped, query loads, 16-point rescan, key posts), put at a quarter
// of a 512-point tile.  PDAE_CHAMFER_SPLIT forces a value (1 = never split).
static int g_chamfer_split = -1;
static int chamfer_split_forced() {
  if (g_chamfer_split < 0) {
    const char *e = getenv("PDAE_CHAMFER_SPLIT");
    g_chamfer_split = e ? atoi(e) : 0;
  }
  return g_chamfer_split;
}
static int chamfer_csplit(long long row_units, int ntiles) {
  const int forced = chamfer_split_forced();
  if (ntiles <= 1) return 1;
  if (forced > 0) return forced < ntiles ? forced : ntiles;
  const long long sms = sm_count();
  int best_nc = 1;
  double best_cost = 0.0;
  for (int nc = 1; nc <= ntiles && nc <= 16; ++nc) {
    const int tpc = (ntiles + nc - 1) / nc;
    const long long chunks = (ntiles + tpc - 1) / tpc;  // chunks that actually hold tiles
    const long long waves = (row_units * chunks + sms - 1) / sms;
    const double cost = static_cast<double>(waves) * (tpc + 0.25);
    if (nc == 1 || cost < best_cost * 0.97) {  // split only for a clear (> 3 %) gain
      best_cost = cost;
      best_nc = nc;
    }
  }
  return best_nc;
}

// phase 0: the whole forward; 1: everything except the column recovery of the symmetric path (the "scan": row
// results final, column minima parked in the workspace); 2: only that recovery.  Paths without a separate recovery
// (tiny clouds, two-scan, empty inputs) do all their work in phase 1 and nothing in phase 2.
static int chamfer_fwd_impl(const float *xyz1, const float *xyz2, int b, int n, int m, float *dist1, float *dist2,
                            int *idx1, int *idx2, void *workspace, size_t workspace_bytes, pdae_stream_t stream,
                            int phase) {
  if (b < 0 || n < 0 || m < 0) return PDAE_E_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t bn = static_cast<size_t>(b) * n, bm = static_cast<size_t>(b) * m;
  if ((bn && (!xyz1 || !dist1 || !idx1)) || (bm && (!xyz2 || !dist2 || !idx2))) return PDAE_E_INVALID;
  if (b == 0) return 0;
  const int nq_max = n > m ? n : m;
  const size_t need = static_cast<size_t>(b) * (n < m ? n : m) * sizeof(uint64_t);  // column keys (the row keys are optional)
  const bool sym = n > 0 && m > 0 && !(n <= SMALL_MAX && m <= SMALL_MAX) && workspace != nullptr &&
                   workspace_bytes >= need && nq_max > 256 && chamfer_variant() < 100 &&
                   b <= 65535;  // the recovery kernels index clouds with gridDim.y; larger batches take the two-scan path
  if (phase == 2 && !sym) return 0;
  if (n == 0 || m == 0) {  // reference: outputs stay at their zero initialisation (chamfer.cu:152-157)
    if (bn) {
      PDAE_CUDA_TRY(cudaMemsetAsync(dist1, 0, bn * sizeof(float), st));
      PDAE_CUDA_TRY(cudaMemsetAsync(idx1, 0, bn * sizeof(int), st));
    }
    if (bm) {
      PDAE_CUDA_TRY(cudaMemsetAsync(dist2, 0, bm * sizeof(float), st));
      PDAE_CUDA_TRY(cudaMemsetAsync(idx2, 0, bm * sizeof(int), st));
    }
    return 0;
  }
  if (n <= SMALL_MAX && m <= SMALL_MAX) {
    const long long warps = 2LL * b;
    const long long grid = (warps + SMALL_WARPS - 1) / SMALL_WARPS;
    if (grid > 0x7fffffffLL) return PDAE_E_UNSUPPORTED;
    chamfer_small_kernel<<<static_cast<unsigned>(grid), SMALL_WARPS * 32, 0, st>>>(xyz1, xyz2, b, n, m, dist1, dist2,
                                                                                  idx1, idx2);
    PDAE_RETURN_IF_LAUNCH_FAILED();
    return 0;
  }
  if (chamfer_tc_applies(b, n, m, workspace ? workspace_bytes : 0)) {  // tensor-core filter + exact evaluation (chamfer_tc.cu)
    if (phase == 2) return 0;
    return chamfer_tc_forward(xyz1, xyz2, b, n, m, dist1, dist2, idx1, idx2, workspace, workspace_bytes, st);
  }
  if (sym) {
    // rows (register-resident queries) = the larger cloud, columns = the smaller one
    const bool swap = m > n;
    const float *rows = swap ? xyz2 : xyz1, *cols = swap ? xyz1 : xyz2;
    const int nr_rows = swap ? m : n, nr_cols = swap ? n : m;
    float *drow = swap ? dist2 : dist1, *dcol = swap ? dist1 : dist2;
    int *irow = swap ? idx2 : idx1, *icol = swap ? idx1 : idx2;
    uint64_t *ck = static_cast<uint64_t *>(workspace);
    const long long ncol = static_cast<long long>(b) * nr_cols;
    const long long slices_per_cloud = (static_cast<long long>(nr_rows) + 127) / 128;
    const bool balanced = chamfer_variant() == 17 && slices_per_cloud <= 128;  // opt-in (measured no faster)
    const long long nrow = static_cast<long long>(b) * nr_rows;
    // column split: default kernel shape only (512-row blocks, 512-point tiles), and only with room for the row keys
    const int nsplit = (!balanced && chamfer_variant() == 0 &&
                        workspace_bytes >= static_cast<size_t>(ncol + nrow) * sizeof(uint64_t))
                           ? chamfer_csplit(static_cast<long long>(b) * ceil_div(nr_rows, 512), ceil_div(nr_cols, 512))
                           : 1;
    // the whole forward in one call: the list recovery writes the rows' results while it is there
    const bool fused_unpack = nsplit > 1 && phase == 0 && recover_unpacks_rows(nr_rows);
    if (phase != 2) {
      const long long nfill = nsplit > 1 ? ncol + nrow : ncol;
      fill_keys_kernel<<<static_cast<unsigned>((nfill + 255) / 256), 256, 0, st>>>(ck, nfill);
      PDAE_RETURN_IF_LAUNCH_FAILED();
      if (balanced) {
        const long long total = slices_per_cloud * b;
        const long long slots = 2LL * sm_count();
        ChamferDir d{rows, cols, drow, irow, nullptr, ck, nr_rows, nr_cols, 0, 0};
        chamfer_min_balanced_kernel<128, 512, 8><<<static_cast<unsigned>(total < slots ? total : slots), 128, 0, st>>>(
            d, static_cast<int>(slices_per_cloud), total);
        PDAE_RETURN_IF_LAUNCH_FAILED();
      } else {
        const int qpc = chamfer_qpc(nr_rows, true);
        ChamferDir d0{rows, cols, drow, irow, nullptr, ck, nr_rows, nr_cols, ceil_div(nr_rows, qpc), 0};
        ChamferDir d1{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0};
        if (nsplit > 1) {  // rows leave as merged keys instead of dist / idx
          d0.csplit = nsplit;
          d0.keys = ck + ncol;
          d0.dist = nullptr;
          d0.idx = nullptr;
        }
        const int rc = launch_min<true>(d0, d1, b, st);
        if (rc) return rc;
        if (nsplit > 1 && !fused_unpack) {  // merged row keys -> dist / idx of the larger cloud
          unpack_keys_kernel<<<static_cast<unsigned>((nrow + 255) / 256), 256, 0, st>>>(ck + ncol, nrow, drow, irow);
          PDAE_RETURN_IF_LAUNCH_FAILED();
        }
      }
    }
    if (phase == 1) return 0;
    if (balanced) {
      const dim3 vgrid(static_cast<unsigned>(slices_per_cloud), b);
      chamfer_col_recover_var_kernel<<<vgrid, 256, 0, st>>>(rows, cols, ck, nr_rows, nr_cols, dcol, icol);
      PDAE_RETURN_IF_LAUNCH_FAILED();
      return 0;
    }
    if (fused_unpack) return launch_col_recover_for_variant(rows, cols, ck, b, nr_rows, nr_cols, dcol, icol, st, ck + ncol, drow, irow);
    return launch_col_recover_for_variant(rows, cols, ck, b, nr_rows, nr_cols, dcol, icol, st);
  }
  const int qpc = chamfer_qpc(nq_max, false);
  ChamferDir d0{xyz1, xyz2, dist1, idx1, nullptr, nullptr, n, m, ceil_div(n, qpc), 0};
  ChamferDir d1{xyz2, xyz1, dist2, idx2, nullptr, nullptr, m, n, ceil_div(m, qpc), 0};
  return launch_min<false>(d0, d1, b, st);
}

// tuning / test hook: force the number of column chunks inside one process (0 = automatic, 1 = never split);
// returns the previous setting, a negative argument only queries.
extern "C" int pdae_tune_chamfer_split(int nc) {
  const int old = chamfer_split_forced();
  if (nc >= 0) g_chamfer_split = nc;
  return old;
}

extern "C" int pdae_chamfer_fwd_f32(const float *xyz1, const float *xyz2, int b, int n, int m, float *dist1,
                                    float *dist2, int *idx1, int *idx2, void *workspace, size_t workspace_bytes,
                                    pdae_stream_t stream) {
  return chamfer_fwd_impl(xyz1, xyz2, b, n, m, dist1, dist2, idx1, idx2, workspace, workspace_bytes, stream, 0);
}

extern "C" int pdae_chamfer_fwd_phase_f32(const float *xyz1, const float *xyz2, int b, int n, int m, float *dist1,
                                          float *dist2, int *idx1, int *idx2, void *workspace, size_t workspace_bytes,
                                          int phase, pdae_stream_t stream) {
  if (phase != 1 && phase != 2) return PDAE_E_INVALID;
  return chamfer_fwd_impl(xyz1, xyz2, b, n, m, dist1, dist2, idx1, idx2, workspace, workspace_bytes, stream, phase);
}

extern "C" int pdae_chamfer_min_keys_u64(const float *queries, const float *refs, int b, int nq, int nr,
                                         int ref_offset, uint64_t *keys, pdae_stream_t stream) {
  if (b < 0 || nq < 0 || nr < 0 || ref_offset < 0) return PDAE_E_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t bq = static_cast<size_t>(b) * nq;
  if (bq == 0) return 0;
  if (!queries || !keys || (nr && !refs)) return PDAE_E_INVALID;
  if (nr == 0) {  // empty slice: identity of MIN
    fill_keys_kernel<<<static_cast<unsigned>((bq + 255) / 256), 256, 0, st>>>(keys, static_cast<long long>(bq));
    PDAE_RETURN_IF_LAUNCH_FAILED();
    return 0;
  }
  const int qpc = chamfer_qpc(nq, false);
  ChamferDir d0{queries, refs, nullptr, nullptr, keys, nullptr, nq, nr, ceil_div(nq, qpc), ref_offset};
  ChamferDir d1{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0};
  const int nsplit = (chamfer_variant() == 0 && nq > 256) ? chamfer_csplit(static_cast<long long>(b) * ceil_div(nq, 512), ceil_div(nr, 512)) : 1;
  if (nsplit > 1) {  // column-split units merge into the output keys
    fill_keys_kernel<<<static_cast<unsigned>((bq + 255) / 256), 256, 0, st>>>(keys, static_cast<long long>(bq));
    PDAE_RETURN_IF_LAUNCH_FAILED();
    d0.csplit = nsplit;
  }
  return launch_min<false>(d0, d1, b, st);
}

// one rank's share of a reference-set-sharded Chamfer forward: all of xyz1 against the local slice of xyz2,
// every pair evaluated once (symmetric kernel): row minima leave as packed keys for the MIN all-reduce, column
// minima are already final for the slice because every rank sees all of xyz1.
extern "C" int pdae_chamfer_sharded_f32(const float *xyz1, const float *xyz2_local, int b, int n, int m_local,
                                        int ref_offset, uint64_t *keys1, float *dist2_local, int *idx2_local,
                                        void *workspace, size_t workspace_bytes, pdae_stream_t stream) {
  if (b < 0 || n < 0 || m_local < 0 || ref_offset < 0) return PDAE_E_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t bn = static_cast<size_t>(b) * n, bm = static_cast<size_t>(b) * m_local;
  if (b == 0) return 0;
  if (bn && (!xyz1 || !keys1)) return PDAE_E_INVALID;
  if (bm && (!xyz2_local || !dist2_local || !idx2_local)) return PDAE_E_INVALID;
  if (b > 65535) return PDAE_E_UNSUPPORTED;
  if (m_local == 0) {
    if (bn) {
      fill_keys_kernel<<<static_cast<unsigned>((bn + 255) / 256), 256, 0, st>>>(keys1, static_cast<long long>(bn));
      PDAE_RETURN_IF_LAUNCH_FAILED();
    }
    return 0;
  }
  if (n == 0) {  // reference: outputs stay zero
    PDAE_CUDA_TRY(cudaMemsetAsync(dist2_local, 0, bm * sizeof(float), st));
    PDAE_CUDA_TRY(cudaMemsetAsync(idx2_local, 0, bm * sizeof(int), st));
    return 0;
  }
  if (!workspace || workspace_bytes < bm * sizeof(uint64_t)) return PDAE_E_WORKSPACE;
  if (chamfer_tc_sharded_applies(b, n, m_local))  // tensor-core filter in column chunks; same keys (chamfer_tc.cu)
    return chamfer_tc_sharded(xyz1, xyz2_local, b, n, m_local, ref_offset, keys1, dist2_local, idx2_local, workspace,
                              workspace_bytes, st);
  uint64_t *ck = static_cast<uint64_t *>(workspace);
  fill_keys_kernel<<<static_cast<unsigned>((bm + 255) / 256), 256, 0, st>>>(ck, static_cast<long long>(bm));
  PDAE_RETURN_IF_LAUNCH_FAILED();
  const int qpc = chamfer_qpc(n, true);
  ChamferDir d0{xyz1, xyz2_local, nullptr, nullptr, keys1, ck, n, m_local, ceil_div(n, qpc), ref_offset};
  ChamferDir d1{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0};
  // a single scene-scale cloud gives few row blocks (100 000 points: 196 for 148 SMs): column-split units, merged
  // through the very keys this call returns
  const int nsplit = chamfer_variant() == 0 ? chamfer_csplit(static_cast<long long>(b) * ceil_div(n, 512), ceil_div(m_local, 512)) : 1;
  if (nsplit > 1) {
    fill_keys_kernel<<<static_cast<unsigned>((bn + 255) / 256), 256, 0, st>>>(keys1, static_cast<long long>(bn));
    PDAE_RETURN_IF_LAUNCH_FAILED();
    d0.csplit = nsplit;
  }
  const int rc = launch_min<true>(d0, d1, b, st);
  if (rc) return rc;
  return launch_col_recover_for_variant(xyz1, xyz2_local, ck, b, n, m_local, dist2_local, idx2_local, st);
}

extern "C" int pdae_chamfer_unpack_keys(const uint64_t *keys, long long count, float *dist, int *idx,
                                        pdae_stream_t stream) {
  if (count < 0) return PDAE_E_INVALID;
  if (count == 0) return 0;
  if (!keys || !dist || !idx) return PDAE_E_INVALID;
  const long long grid = (count + 255) / 256;
  if (grid > 0x7fffffffLL) return PDAE_E_UNSUPPORTED;
  unpack_keys_kernel<<<static_cast<unsigned>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(keys, count, dist, idx);
  PDAE_RETURN_IF_LAUNCH_FAILED();
  return 0;
}

extern "C" int pdae_chamfer_bwd_f32(const float *xyz1, const float *xyz2, const int *idx1, const int *idx2,
                                    const float *gd1, const float *gd2, int b, int n, int m, float *gx1, float *gx2,
                                    pdae_stream_t stream) {
  if (b < 0 || n < 0 || m < 0) return PDAE_E_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long t1 = static_cast<long long>(b) * n, t2 = static_cast<long long>(b) * m;
  if ((t1 && (!xyz1 || !gx1)) || (t2 && (!xyz2 || !gx2))) return PDAE_E_INVALID;
  if (n == 0 || m == 0 || b == 0) {  // reference: the loops never execute, grads stay zero
    if (t1) PDAE_CUDA_TRY(cudaMemsetAsync(gx1, 0, static_cast<size_t>(t1) * 3 * sizeof(float), st));
    if (t2) PDAE_CUDA_TRY(cudaMemsetAsync(gx2, 0, static_cast<size_t>(t2) * 3 * sizeof(float), st));
    return 0;
  }
  if (!idx1 || !idx2 || !gd1 || !gd2) return PDAE_E_INVALID;
  return launch_chamfer_bwd(xyz1, xyz2, idx1, idx2, GradArrays{gd1, gd2}, b, n, m, gx1, gx2, st);
}

extern "C" size_t pdae_chamfer_loss_workspace_bytes(void) { return 2 * LOSS_BLOCKS * sizeof(float); }

extern "C" int pdae_chamfer_loss_f32(const float *dist1, const float *dist2, int b, int n, int m, int l1, float *loss3,
                                     void *workspace, size_t workspace_bytes, pdae_stream_t stream) {
  if (b < 0 || n < 0 || m < 0 || !loss3) return PDAE_E_INVALID;
  if (!workspace || workspace_bytes < pdae_chamfer_loss_workspace_bytes()) return PDAE_E_WORKSPACE;
  const long long c1 = static_cast<long long>(b) * n, c2 = static_cast<long long>(b) * m;
  if ((c1 && !dist1) || (c2 && !dist2)) return PDAE_E_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float *partial = static_cast<float *>(workspace);
  chamfer_loss_partial_kernel<<<LOSS_BLOCKS, 256, 0, st>>>(dist1, dist2, c1, c2, l1 ? 1 : 0, partial);
  PDAE_RETURN_IF_LAUNCH_FAILED();
  chamfer_loss_final_kernel<<<1, 32, 0, st>>>(partial, c1, c2, l1 ? 1 : 0, loss3);
  PDAE_RETURN_IF_LAUNCH_FAILED();
  return 0;
}

extern "C" int pdae_chamfer_loss_bwd_f32(const float *xyz1, const float *xyz2, const int *idx1, const int *idx2,
                                         const float *dist1, const float *dist2, const float *gloss, float w1, float w2,
                                         int b, int n, int m, int l1, float *gx1, float *gx2, pdae_stream_t stream) {
  if (b < 0 || n < 0 || m < 0) return PDAE_E_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long t1 = static_cast<long long>(b) * n, t2 = static_cast<long long>(b) * m;
  if ((t1 && (!xyz1 || !gx1)) || (t2 && (!xyz2 || !gx2))) return PDAE_E_INVALID;
  if (n == 0 || m == 0 || b == 0) {
    if (t1) PDAE_CUDA_TRY(cudaMemsetAsync(gx1, 0, static_cast<size_t>(t1) * 3 * sizeof(float), st));
    if (t2) PDAE_CUDA_TRY(cudaMemsetAsync(gx2, 0, static_cast<size_t>(t2) * 3 * sizeof(float), st));
    return 0;
  }
  if (!idx1 || !idx2 || !gloss || (l1 && (!dist1 || !dist2))) return PDAE_E_INVALID;
  const float n1 = static_cast<float>(t1), n2 = static_cast<float>(t2);
  return launch_chamfer_bwd(xyz1, xyz2, idx1, idx2, GradMean{gloss, dist1, dist2, w1 / n1, w2 / n2, l1 ? 1 : 0}, b, n, m, gx1,
                            gx2, st);
}
