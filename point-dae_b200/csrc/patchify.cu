// patchify.cu -- the whole patchifier of `Group.forward` (models/PointCAE_transformer.py:61-86) in ONE launch:
// utils/misc.py:13-20 `fps` (furthest point sampling + centre gather) and the kNN + gather + centre-subtract tail.
//
// FPS is strictly sequential (one centre per iteration, ~0.27 us each) and leaves the SM it runs on almost idle; the
// kNN of a centre needs nothing but that centre and the cloud -- and FPS iteration j+1 evaluates the distance of EVERY
// point to centre j anyway.  So one CTA per cloud runs both, warp-specialised:
//   * warps 0..3 are the FPS of fps.cu (`fps_reg_kernel<128, P, 2>`: points in registers, rank-ordered staged copy,
//     REDUX + 4-key register tree, one barrier per iteration -- here a NAMED barrier over these four warps only).  Every
//     iteration also drops the distances it has just computed (to the previous centre) into a ring of shared-memory
//     buffers (4 STS.128 per thread), posts the new centre, and thread 0 arrives on the mbarrier of the search task
//     whose last centre now has its distances;
//   * warps 4.. are kNN consumers: consumer w waits for the mbarriers of tasks w, w + NCW, ... (QW centres each).  The
//     FPS distances use the reference FPS rounding order (y product first), KNN_CUDA's use x first, so they serve as a
//     FILTER: both forms are within 3 ulp of the true value, hence within 2^-21.4 of each other; the consumer takes the
//     k-th smallest of 64 segment minima of the FPS distances as threshold, widens it by 2^-20 (twice), records the
//     4-point groups that hold a distance below it, gives the ring buffer back to the FPS warps, and re-evaluates only
//     the recorded groups EXACTLY (KNN_CUDA order) from a planar copy of the cloud kept in the FPS register order.
//     The survivors are ordered by (distance, index) with the tagged 32-bit sort of knn4.cu (exact 64-bit fallback);
//     queries with mass ties, a non-finite or denormal-range threshold are redone with the exact streaming warp-select.
// The search of centre j therefore costs ~40 % of the instructions of the stand-alone kernel and overlaps the FPS
// iterations j+2.., and the launch ends a few microseconds after the last FPS iteration instead of a whole kNN kernel
// later.  Same results as the two-launch path bit for bit (sampling_gpu.cu:72-176 rounding and tie rule; KNN_CUDA's
// sequential `ssd += tmp*tmp`; ascending by (distance, index)).
#include <cstdlib>

#include "fps_rank.cuh"
#include "knn_select.cuh"

namespace pdae {

namespace {

struct PatchArgs {
  const float *data;  // (b, n, 3)
  int *fps_idx;       // (b, g) int32
  float *centers;     // (b, g, 3)
  int64_t *idx;       // optional (b, g, k)
  float *group;       // (b, g, k, 3)
  int raw_group;
  GroupAffine aff;    // AFF instances: the corrupted copy of every patch and centre (corrupt.cu / knn4.cu epilogue)
  int n, g, k;
  int npb;   // ceil(n / 512): points per reference slot
  int nbuf;  // distance buffers in the ring (power of two, >= 2 QW)
  int lg_nbuf;
  int overlap_previous;  // launch attribute only (not read by the kernel): programmatic stream serialization
  long long *trace;  // diagnostics (pdae_patchify_trace): clock64 stamps of CTA 0, else NULL
};

__device__ __forceinline__ uint32_t pf_smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void pf_mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pf_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void pf_mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pf_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool pf_mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(pf_smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
template <int ID>
__device__ __forceinline__ void named_barrier(int threads) {
  asm volatile("bar.sync %0, %1;" ::"n"(ID), "r"(threads) : "memory");
}

constexpr int PF_E = 2;                 // keys per lane in the final sort (k <= 32)
constexpr int PF_CAP = 32 * PF_E;       // keys per query queue
constexpr int PF_LC = 5 * PF_E;         // lane-private step slots per query (+1 that absorbs overflow)
constexpr size_t PF_PQ = (PF_LC + 1) * 64 > PF_CAP * 8 ? (PF_LC + 1) * 64 : PF_CAP * 8;  // step lists, later the key queue
constexpr size_t PF_DENSE = static_cast<size_t>(PF_LC) * 32 * 4;
constexpr int PF_FPS_T = 128;  // FPS threads (one warp per scheduler)

}  // namespace

// P = points per FPS thread (n <= 128 * P), QW = centres per consumer task, NCW = consumer warps, AFF = the epilogue
// also emits the affinely corrupted patches and centres (models/PointCAE_transformer.py:1011-1017)
template <int P, int QW, int NCW, bool AFF>
__global__ void __launch_bounds__(PF_FPS_T + NCW * 32, 1) fps_group_kernel(const PatchArgs a) {
  constexpr int E = PF_E, CAP = PF_CAP, LC = PF_LC;
  constexpr size_t PQ = PF_PQ;
  constexpr int S = 2, LG_BS = 9;  // reference block size 512 (512 <= n), CTA of 128 FPS threads: 4 slots per thread
  constexpr int PG = P >> S;
  constexpr int NP = P * PF_FPS_T;  // register slots of the cloud (>= n); slot f = (q4 * 128 + tid) * 4 + e holds the
                                    // point of register 4 q4 + e of FPS thread tid
  constexpr int NQ4 = NP >> 2;      // 4-slot groups = float4 per plane / per distance buffer
  static_assert(P % 4 == 0, "registers are handed over four at a time");
  static_assert(PG == 4 || PG == 2, "slot -> cloud index shortcut of the rescan");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int n = a.n, m = a.g, k = a.k, npb = a.npb, nbuf = a.nbuf;
  const int ntask = (m + QW - 1) / QW;
  unsigned long long *slots = reinterpret_cast<unsigned long long *>(smem_raw);          // [2][32] FPS reduction
  float4 *sp = reinterpret_cast<float4 *>(smem_raw + 512);                               // [npb * 512] rank-ordered cloud
  float *planes = reinterpret_cast<float *>(sp + (static_cast<size_t>(npb) << LG_BS));   // [3][NP] slot order
  float *ring = planes + 3 * NP;                                                         // [nbuf][NP] FPS distances
  float4 *cen_s = reinterpret_cast<float4 *>(ring + static_cast<size_t>(nbuf) * NP);     // [m] centres as selected
  uint64_t *full = reinterpret_cast<uint64_t *>(cen_s + m);                              // [ntask] centres + distances posted
  uint64_t *empty = full + ntask;                                                        // [nbuf] ring buffer given back
  unsigned char *warp_area = reinterpret_cast<unsigned char *>(empty + nbuf);            // (8-byte aligned: u64 key queues)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cloud_id = blockIdx.x;
  const float *__restrict__ cloud = a.data + static_cast<size_t>(cloud_id) * n * 3;
  const float INF = __int_as_float(0x7f800000);

  for (int i = threadIdx.x; i < ntask + nbuf; i += blockDim.x) pf_mbar_init(full + i, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  // trace (CTA 0 only): [0] start, [1 + j] centre j posted, [1 + g + 8 t + p] task t: 0 centres seen, 1 threshold known,
  // 2 step lists written, 3 candidates queued, 4 patches emitted, 5 (overflow queries redone)
  long long *trace = blockIdx.x == 0 ? a.trace : nullptr;
  if (trace && threadIdx.x == 0) trace[0] = clock64();

  if (warp < PF_FPS_T / 32) {
    // ---- furthest point sampling: fps.cu `fps_reg_kernel<128, P, 2>` -------------------------------------------------
    const int tid = threadIdx.x;
    int *__restrict__ out = a.fps_idx + static_cast<size_t>(cloud_id) * m;
    const unsigned rank_base = (__brev(static_cast<unsigned>(tid)) >> (32 - 7)) << (22 + S);
    auto rank_of_reg = [&](int i) -> unsigned {
      return rank_base + (static_cast<unsigned>(i / PG) << 22) + static_cast<unsigned>(i % PG);
    };
    auto pos_of_rank = [&](unsigned r) -> unsigned { return (r >> 22) * static_cast<unsigned>(npb) + (r & 0x3fffffu); };
    // coordinates as register PAIRS: the distance update runs on the packed fp32 pipe forms (FADD2 / FMUL2 / FFMA2, IEEE
    // per half), 6 instructions per two points instead of 12
    float2 px2[P / 2], py2[P / 2], pz2[P / 2];
    float pt[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const int kk = tid + fps_point_of_reg<S, PG>(p) * PF_FPS_T;
      float x = INF, y = 0.f, z = 0.f, t = -2.0f;  // padding: distance +inf (or NaN) to every centre, never sampled
      if (kk < n) {
        x = __ldg(cloud + static_cast<size_t>(kk) * 3);
        y = __ldg(cloud + static_cast<size_t>(kk) * 3 + 1);
        z = __ldg(cloud + static_cast<size_t>(kk) * 3 + 2);
        const float mag = dist_yxz(x, y, z);
        t = (static_cast<double>(mag) <= 1e-3) ? -2.0f : 1e10f;  // sampling_gpu.cu:103-104
        sp[pos_of_rank(rank_of_reg(p))] = make_float4(x, y, z, __int_as_float(kk));
      }
      (p & 1 ? px2[p / 2].y : px2[p / 2].x) = x;
      (p & 1 ? py2[p / 2].y : py2[p / 2].x) = y;
      (p & 1 ? pz2[p / 2].y : pz2[p / 2].x) = z;
      pt[p] = t;
    }
    named_barrier<1>(PF_FPS_T);
    float4 o = sp[0];  // point 0 has rank 0
    if (tid == 0) {
      out[0] = 0;
      cen_s[0] = o;
      if (trace) trace[1] = clock64();
    }
    // distances of all points to centre c = j - 1 go to ring buffer c % nbuf, once the search of centre c - nbuf has
    // given it back
    auto hand_over = [&](int c, const float (&dd)[P]) {
      const int buf = c & (nbuf - 1);
      if (c >= nbuf) {
        const uint32_t par = static_cast<uint32_t>((c >> a.lg_nbuf) - 1) & 1u;
        while (!pf_mbar_try_wait(empty + buf, par)) {}
      }
      float4 *dst = reinterpret_cast<float4 *>(ring + static_cast<size_t>(buf) * NP) + tid;
#pragma unroll
      for (int q4 = 0; q4 < P / 4; ++q4) dst[q4 * PF_FPS_T] = make_float4(dd[4 * q4], dd[4 * q4 + 1], dd[4 * q4 + 2], dd[4 * q4 + 3]);
    };
    for (int j = 1; j < m; ++j) {
      float v[P], dd[P];
      int vi[P];
      const float2 ox2 = make_float2(o.x, o.x), oy2 = make_float2(o.y, o.y), oz2 = make_float2(o.z, o.z);
#pragma unroll
      for (int h = 0; h < P / 2; ++h) {
        const float2 d2 = dist_yxz2(sub2(px2[h], ox2), sub2(py2[h], oy2), sub2(pz2[h], oz2));
        dd[2 * h] = d2.x, dd[2 * h + 1] = d2.y;
      }
#pragma unroll
      for (int p = 0; p < P; ++p) {
        pt[p] = fminf(dd[p], pt[p]);
        v[p] = pt[p];
        vi[p] = p;
      }
#pragma unroll
      for (int stride = 1; stride < P; stride *= 2) {
#pragma unroll
        for (int i = 0; i + stride < P; i += 2 * stride) {
          const bool take = v[i + stride] > v[i];
          v[i] = take ? v[i + stride] : v[i];
          vi[i] = take ? vi[i + stride] : vi[i];
        }
      }
      const bool any = v[0] > -1.0f;
      const float best = any ? v[0] : -1.0f;
      const unsigned myrank = any ? rank_of_reg(vi[0]) : 0u;
      // block arg-max over the four FPS warps: one 64-bit key per warp, register tree after the named barrier
      const unsigned FULLM = 0xffffffffu;
      const unsigned vb = fps_val_bits(best);
      const unsigned mx = __reduce_max_sync(FULLM, vb);
      const unsigned rr = __reduce_min_sync(FULLM, vb == mx ? myrank : 0xffffffffu);
      unsigned long long *buf = slots + (j & 1) * 32;
      if (lane == 0) buf[warp] = (static_cast<unsigned long long>(mx) << 32) | static_cast<unsigned>(~rr);
      hand_over(j - 1, dd);  // (off the arg-max chain: issued while the REDUX results travel)
      named_barrier<1>(PF_FPS_T);
      const ulonglong2 k01 = *reinterpret_cast<const ulonglong2 *>(buf);
      const ulonglong2 k23 = *reinterpret_cast<const ulonglong2 *>(buf + 2);
      const unsigned long long ka = k01.x > k01.y ? k01.x : k01.y, kb = k23.x > k23.y ? k23.x : k23.y;
      const unsigned r = ~static_cast<unsigned>(ka > kb ? ka : kb);
      o = sp[pos_of_rank(r)];
      if (tid == 0) {
        out[j] = __float_as_int(o.w);
        cen_s[j] = o;
        if (j % QW == 0) pf_mbar_arrive(full + (j - 1) / QW);  // centres up to j - 1 have their distances (barrier above)
        if (trace) trace[1 + j] = clock64();
      }
    }
    {  // the last centre's distances
      float dd[P];
#pragma unroll
      for (int p = 0; p < P; ++p)
        dd[p] = dist_yxz(__fsub_rn(p & 1 ? px2[p / 2].y : px2[p / 2].x, o.x), __fsub_rn(p & 1 ? py2[p / 2].y : py2[p / 2].x, o.y),
                         __fsub_rn(p & 1 ? pz2[p / 2].y : pz2[p / 2].x, o.z));
      hand_over(m - 1, dd);
      named_barrier<1>(PF_FPS_T);
      if (tid == 0) pf_mbar_arrive(full + (m - 1) / QW);
    }
    return;
  }

  // ---- kNN consumers --------------------------------------------------------------------------------------------------
  constexpr int CNT = NCW * 32;
  const int cw = warp - PF_FPS_T / 32, ctid = threadIdx.x - PF_FPS_T;
  const unsigned FULL = 0xffffffffu;
  const unsigned lt_mask = (1u << lane) - 1u;
  // cloud index of slot f
  auto point_of_slot = [&](int f) -> int {
    return ((f >> 2) & (PF_FPS_T - 1)) + fps_point_of_reg_rt<S, PG>(4 * (f >> 9) + (f & 3)) * PF_FPS_T;
  };
  for (int f = ctid; f < NP; f += CNT) {  // planar copy of the cloud in slot order, padding: x = +inf
    const int kk = point_of_slot(f);
    float x = INF, y = 0.f, z = 0.f;
    if (kk < n) x = __ldg(cloud + 3 * kk), y = __ldg(cloud + 3 * kk + 1), z = __ldg(cloud + 3 * kk + 2);
    planes[f] = x, planes[NP + f] = y, planes[2 * NP + f] = z;
  }
  named_barrier<2>(CNT);

  unsigned char *my_area = warp_area + static_cast<size_t>(cw) * (QW * PQ + PF_DENSE);
  uint32_t *dense = reinterpret_cast<uint32_t *>(my_area + static_cast<size_t>(QW) * PQ);  // [32 * LC]
  const uint32_t lbase_s = pf_smem_u32(my_area) + 2 * lane;
  const int klane = k - 1;
  const uint32_t dir_mask = warp_sort_dir_mask(lane);
  const float4 *pl4 = reinterpret_cast<const float4 *>(planes);
  constexpr int nsteps = P;  // 128 slots per step

  for (int task = cw; task < ntask; task += NCW) {
    while (!pf_mbar_try_wait(full + task, 0u)) {}
    long long *tr = trace ? trace + 1 + m + 8 * task : nullptr;
    if (tr && lane == 0) tr[0] = clock64();
    const int qbase = task * QW;
    float q0[QW], q1[QW], q2[QW];
    const float4 *dq[QW];  // the query's FPS distances, this lane's column
#pragma unroll
    for (int qi = 0; qi < QW; ++qi) {
      const int cidx = min(qbase + qi, m - 1);
      const float4 c = cen_s[cidx];
      q0[qi] = c.x, q1[qi] = c.y, q2[qi] = c.z;
      dq[qi] = reinterpret_cast<const float4 *>(ring + static_cast<size_t>(cidx & (nbuf - 1)) * NP) + lane;
    }
    if (lane < 3 * QW && qbase + lane / 3 < m)  // utils/misc.py:18-19: the centre rows
      a.centers[(static_cast<size_t>(cloud_id) * m + qbase) * 3 + lane] =
          reinterpret_cast<const float *>(cen_s + qbase + lane / 3)[lane % 3];

    // pass 1: 64 segment minima of the FPS distances -> their k-th smallest bounds the k-th FPS distance from above
    float mn[QW][E];
#pragma unroll
    for (int qi = 0; qi < QW; ++qi)
#pragma unroll
      for (int e = 0; e < E; ++e) mn[qi][e] = INF;
    float4 keep[QW == 1 ? nsteps : 1];  // one query per task: its distances stay in registers for the second pass
#pragma unroll
    for (int s = 0; s < nsteps; ++s) {
#pragma unroll
      for (int qi = 0; qi < QW; ++qi) {
        const float4 D = dq[qi][s * 32];
        if (QW == 1) keep[s] = D;
        mn[qi][0] = min3(mn[qi][0], D.x, D.y);
        mn[qi][1] = min3(mn[qi][1], D.z, D.w);
      }
    }
    // tau_fps -> thresholds: the exact (KNN_CUDA-order) distance of every true neighbour is <= tau_fps (1 + 2^-20) =: tk,
    // and its FPS-order distance <= tk (1 + 2^-20) =: tf (both orders are within 3 roundings of the real value)
    float tk[QW], tf[QW];
    {
      // the 32nd smallest of the 64 minima (>= the k-th for every k <= 32): both halves sorted across the lanes, then
      // min(A[i], B[31 - i]) are the 32 smallest and their maximum is the bound (15 network stages instead of 21)
      uint32_t sv[2 * QW][1];
#pragma unroll
      for (int qi = 0; qi < QW; ++qi) sv[2 * qi][0] = __float_as_uint(mn[qi][0]), sv[2 * qi + 1][0] = __float_as_uint(mn[qi][1]);
      warp_sort_u32_multi<2 * QW, 1>(sv, lane, dir_mask);
#pragma unroll
      for (int qi = 0; qi < QW; ++qi) {
        const uint32_t rev = __shfl_sync(FULL, sv[2 * qi + 1][0], 31 - lane);
        const float tau = __uint_as_float(__reduce_max_sync(FULL, min(sv[2 * qi][0], rev)));
        tk[qi] = __fmul_rn(tau, 1.00000095367431640625f);
        tf[qi] = __fmul_rn(tk[qi], 1.00000095367431640625f);
        // tau == 0 is exact in both orders (every product is zero); below 1e-30 products may have underflowed and the
        // relative bound is void -> exact fallback (as for a non-finite tau: tf = +inf)
        if (tau > 0.f && tau < 1e-30f) tf[qi] = INF;
      }
    }
    if (tr && lane == 0) tr[1] = clock64();

    // pass 2: which 4-slot steps of the lane hold an FPS distance <= tf (lane-private lists: the running step number is
    // stored unconditionally, the slot only advances on a hit, so a miss is overwritten by the next step)
    uint32_t lp[QW], lend[QW];
#pragma unroll
    for (int qi = 0; qi < QW; ++qi) lp[qi] = lbase_s + qi * static_cast<uint32_t>(PQ), lend[qi] = lp[qi] + LC * 64;
#pragma unroll
    for (int s = 0; s < nsteps; ++s) {
#pragma unroll
      for (int qi = 0; qi < QW; ++qi) {
        const float4 D = QW == 1 ? keep[s] : dq[qi][s * 32];
        const float mm = min3(D.x, D.y, fminf(D.z, D.w));
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(lp[qi]), "r"(s) : "memory");
        if (mm <= tf[qi]) lp[qi] = min(lp[qi] + 64u, lend[qi]);
      }
    }
    __syncwarp();
    if (lane < QW && qbase + lane < m) pf_mbar_arrive(empty + ((qbase + lane) & (nbuf - 1)));  // the ring buffers go back
    if (tr && lane == 0) tr[2] = clock64();

    // rescan: the recorded steps, compacted across the warp, evaluated exactly from the planes (one group per lane and
    // round); points with an exact distance <= tk go to the query's key queue
    int qn[QW];
    bool ovf[QW];
#pragma unroll 1
    for (int qi = 0; qi < QW; ++qi) {
      uint32_t lpq = lp[0];
      float tq = tk[0], tfq = tf[0], f0 = q0[0], f1 = q1[0], f2 = q2[0];
#pragma unroll
      for (int j = 1; j < QW; ++j)
        if (j == qi) lpq = lp[j], tq = tk[j], tfq = tf[j], f0 = q0[j], f1 = q1[j], f2 = q2[j];
      int cnt = 0;
      bool over = !(tfq < INF);
      const int c = static_cast<int>(lpq - lbase_s - qi * static_cast<uint32_t>(PQ)) >> 6;
      over = over || __any_sync(FULL, c >= LC);  // a lane filled its list: it may have dropped steps
      int incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += t;
      }
      const int total = __shfl_sync(FULL, incl, 31);
      if (!over && total > 32 * LC) over = true;
      if (!over && total > 0) {
        const int maxc = __reduce_max_sync(FULL, c);
        const int off = incl - c;
        for (int it = 0; it < maxc; ++it)
          if (it < c)
            dense[off + it] = (static_cast<uint32_t>(reinterpret_cast<const unsigned short *>(my_area + qi * PQ)[it * 32 + lane]) << 5) | lane;
        __syncwarp();
        uint64_t *qq = reinterpret_cast<uint64_t *>(my_area + qi * PQ);
        const float2 g0 = make_float2(f0, f0), g1 = make_float2(f1, f1), g2 = make_float2(f2, f2);
        // two groups per lane and round (the ~40 recorded groups of a query are one round): all six LDS.128 and both
        // distance chains are in flight before the first ballot
        for (int e0 = 0; e0 < total; e0 += 64) {
          bool act[2];
          uint32_t quad[2];
          float dv[2][4];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            act[h] = e0 + 32 * h + lane < total;
            quad[h] = act[h] ? dense[e0 + 32 * h + lane] : 0u;  // = q4 * 128 + FPS thread
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float4 X = pl4[quad[h]], Y = pl4[NQ4 + quad[h]], Z = pl4[2 * NQ4 + quad[h]];
            const float2 dxa = sub2(make_float2(X.x, X.y), g0), dxb = sub2(make_float2(X.z, X.w), g0);
            const float2 dya = sub2(make_float2(Y.x, Y.y), g1), dyb = sub2(make_float2(Y.z, Y.w), g1);
            const float2 dza = sub2(make_float2(Z.x, Z.y), g2), dzb = sub2(make_float2(Z.z, Z.w), g2);
            const float2 dA = fma2(dza, dza, fma2(dya, dya, mul2(dxa, dxa)));
            const float2 dB = fma2(dzb, dzb, fma2(dyb, dyb, mul2(dxb, dxb)));
            dv[h][0] = dA.x, dv[h][1] = dA.y, dv[h][2] = dB.x, dv[h][3] = dB.y;
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            // cloud index of slot e of the group: FPS thread + 128 * fps_point_of_reg(4 q4 + e) = base + a constant per e
            // (PG = 4: brev2(q4) + 4 e;  PG = 2: q4 + 2 (e >> 1) + 4 (e & 1))
            const int q4 = static_cast<int>(quad[h] >> 7);
            const int kbase = static_cast<int>(quad[h] & (PF_FPS_T - 1)) + PF_FPS_T * (PG == 4 ? ((q4 & 1) << 1 | (q4 >> 1)) : q4);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const bool hit = act[h] && dv[h][e] <= tq;
              const unsigned mk = __ballot_sync(FULL, hit);
              const int pos = cnt + __popc(mk & lt_mask);
              if (hit && pos < CAP)
                qq[pos] = pack_key(dv[h][e], static_cast<uint32_t>(kbase + PF_FPS_T * (PG == 4 ? 4 * e : 2 * (e >> 1) + 4 * (e & 1))));
              cnt += __popc(mk);
            }
          }
        }
        __syncwarp();
        over = cnt > CAP;
      }
#pragma unroll
      for (int j = 0; j < QW; ++j)
        if (j == qi) qn[j] = cnt, ovf[j] = over;
    }
    __syncwarp();
    if (tr && lane == 0) tr[3] = clock64();

    auto emit = [&](int qidx, float f0, float f1, float f2, const uint64_t (&keys)[E]) {
      const size_t bq = static_cast<size_t>(cloud_id) * m + qidx;
      const bool raw = a.raw_group != 0;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int p = e * 32 + lane;
        if (p < k) {
          const uint32_t ji = static_cast<uint32_t>(keys[e]);
          if (a.idx) a.idx[bq * k + p] = static_cast<int64_t>(ji);
          // models/PointCAE_transformer.py:84-85: neighbours relative to the centre
          const int reg = fps_reg_of_point_rt<S, PG>(static_cast<int>(ji >> 7));
          const int f = ((reg >> 2) << 9) + ((ji & (PF_FPS_T - 1)) << 2) + (reg & 3);
          const float x = planes[f], y = planes[NP + f], z = planes[2 * NP + f];
          float *g = a.group + (bq * k + p) * 3;
          if constexpr (!AFF) {
            g[0] = raw ? x : __fsub_rn(x, f0);
            g[1] = raw ? y : __fsub_rn(y, f1);
            g[2] = raw ? z : __fsub_rn(z, f2);
          } else {
            // the reference re-adds the centre to the centred patch, transforms patch and centre with the same matrices
            // and subtracts the centres again (same rounding steps as knn4_emit)
            const float *mats = a.aff.mats + static_cast<size_t>(cloud_id) * a.aff.t * 9;
            float ax = __fadd_rn(__fsub_rn(x, f0), f0), ay = __fadd_rn(__fsub_rn(y, f1), f1), az = __fadd_rn(__fsub_rn(z, f2), f2);
            g[0] = __fsub_rn(ax, f0), g[1] = __fsub_rn(ay, f1), g[2] = __fsub_rn(az, f2);
            float cx = f0, cy = f1, cz = f2;
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
    };

    // exact order of the candidates: 32-bit words (distance bits, low TB bits = queue position) for all queries in
    // lockstep; the full 64-bit keys only when two of the first k+1 words agree above the tag
    constexpr int TB = 6;
    {
      uint32_t w[QW][E];
#pragma unroll
      for (int qi = 0; qi < QW; ++qi) {
        const uint64_t *qq = reinterpret_cast<const uint64_t *>(my_area + qi * PQ);
        const int nn = ovf[qi] ? 0 : qn[qi];
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int p = e * 32 + lane;
          w[qi][e] = p < nn ? ((static_cast<uint32_t>(qq[p] >> 32) & ~((1u << TB) - 1u)) | static_cast<uint32_t>(p)) : 0xffffffffu;
        }
      }
      warp_sort_u32_multi<QW, E>(w, lane, dir_mask);
#pragma unroll
      for (int qi = 0; qi < QW; ++qi) {
        const int qidx = qbase + qi;
        if (ovf[qi] || qidx >= m) continue;  // warp-uniform
        const uint64_t *qq = reinterpret_cast<const uint64_t *>(my_area + qi * PQ);
        bool amb = false;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          uint32_t nxt = __shfl_down_sync(FULL, w[qi][e], 1);
          const uint32_t head = __shfl_sync(FULL, w[qi][e + 1 < E ? e + 1 : e], 0);
          if (lane == 31) nxt = e + 1 < E ? head : 0xffffffffu;
          amb |= (e * 32 + lane < k) && (w[qi][e] >> TB) == (nxt >> TB);
        }
        uint64_t keys[E];
        if (!__any_sync(FULL, amb)) {
#pragma unroll
          for (int e = 0; e < E; ++e) keys[e] = w[qi][e] == 0xffffffffu ? KEY_INF : qq[w[qi][e] & ((1u << TB) - 1u)];
        } else {
#pragma unroll
          for (int e = 0; e < E; ++e) keys[e] = (e * 32 + lane) < qn[qi] ? qq[e * 32 + lane] : KEY_INF;
          warp_sort_multi<E>(keys, lane);
        }
        emit(qidx, q0[qi], q1[qi], q2[qi], keys);
      }
    }
    if (tr && lane == 0) tr[4] = clock64();
    // queries whose lists or queue overflowed (mass ties) or whose threshold is unusable: exact streaming warp-select
#pragma unroll 1
    for (int qi = 0; qi < QW; ++qi) {
      bool flagged = ovf[0];
      float f0 = q0[0], f1 = q1[0], f2 = q2[0];
#pragma unroll
      for (int j = 1; j < QW; ++j)
        if (j == qi) flagged = ovf[j], f0 = q0[j], f1 = q1[j], f2 = q2[j];
      if (!flagged || qbase + qi >= m) continue;  // warp-uniform
      __syncwarp();
      WarpSelect<1> sel;
      sel.init();
      uint64_t *wq = reinterpret_cast<uint64_t *>(my_area);  // 64 entries
      for (int f0i = 0; f0i < NP; f0i += 32) {
        const int f = f0i + lane;
        const int kk = point_of_slot(f);
        const bool in = kk < n;
        const float d = dist_seq3(__fsub_rn(planes[f], f0), __fsub_rn(planes[NP + f], f1), __fsub_rn(planes[2 * NP + f], f2));
        const uint64_t key = pack_key(d, static_cast<uint32_t>(kk));
        sel.offer(in && key < sel.tau, key, wq, lane, 0, klane);
      }
      sel.finish(wq, lane);
      uint64_t keys[E] = {sel.L[0], KEY_INF};
      emit(qbase + qi, f0, f1, f2, keys);
    }
    __syncwarp();
    if (tr && lane == 0) tr[5] = clock64();
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------------
static int pf_env_int(const char *name, int dflt) {
  const char *s = std::getenv(name);
  return s && *s ? std::atoi(s) : dflt;
}
struct PatchTune {
  int enabled, qw, ncw;
  long long *trace;
};
static PatchTune &patch_tune() {
  static PatchTune t{pf_env_int("PDAE_PATCHIFY", 1), pf_env_int("PDAE_PATCHIFY_QW", 1), pf_env_int("PDAE_PATCHIFY_NCW", 8), nullptr};
  return t;
}

static size_t patch_smem_bytes(const PatchArgs &a, int p, int qw, int ncw, int nbuf) {
  const int ntask = (a.g + qw - 1) / qw;
  const size_t np = static_cast<size_t>(p) * PF_FPS_T;
  return 512 + (static_cast<size_t>(a.npb) << 9) * 16 + np * 12 + static_cast<size_t>(nbuf) * np * 4 + static_cast<size_t>(a.g) * 16 +
         static_cast<size_t>(ntask + nbuf + ((ntask + nbuf) & 1)) * 8 + static_cast<size_t>(ncw) * (qw * PF_PQ + PF_DENSE);
}

template <int P, int QW, int NCW, bool AFF>
static int patch_launch(PatchArgs a, int b, cudaStream_t st) {
  // ring of distance buffers: as many as fit (a consumer gives a buffer back after its second pass, ~10 FPS iterations
  // after the centre was posted when the SM is busy); at least 2 QW so that the FPS warps never wait for their own task
  size_t smem = 0;
  for (a.nbuf = 16; a.nbuf >= 2 * QW && a.nbuf >= 2; a.nbuf >>= 1) {
    smem = patch_smem_bytes(a, P, QW, NCW, a.nbuf);
    if (smem <= 220 * 1024) break;
  }
  for (a.lg_nbuf = 0; (1 << a.lg_nbuf) < a.nbuf; ++a.lg_nbuf) {}
  if (a.nbuf < 2 * QW || a.nbuf < 2 || smem > 220 * 1024) return PDAE_E_UNSUPPORTED;
  PDAE_CUDA_TRY(cudaFuncSetAttribute(fps_group_kernel<P, QW, NCW, AFF>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  if (a.overlap_previous) {
    // the caller vouches that the previous kernel on this stream does not produce this launch's inputs: the grid may be
    // scheduled as soon as that kernel's CTAs have all started (it triggers `griddepcontrol.launch_dependents`) or exited
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(b), 1, 1);
    cfg.blockDim = dim3(PF_FPS_T + NCW * 32, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PDAE_CUDA_TRY(cudaLaunchKernelEx(&cfg, fps_group_kernel<P, QW, NCW, AFF>, a));
    return 0;
  }
  fps_group_kernel<P, QW, NCW, AFF><<<b, PF_FPS_T + NCW * 32, smem, st>>>(a);
  PDAE_RETURN_IF_LAUNCH_FAILED();
  return 0;
}

template <int P, bool AFF>
static int patch_launch_p(const PatchArgs &a, int b, int qw, int ncw, cudaStream_t st) {
  // instantiated shapes: one or two centres per task; 8 or 12 consumer warps (4 and 6 were measured slower: 40 / 32.5 us)
  if (ncw <= 8) return qw == 2 ? patch_launch<P, 2, 8, AFF>(a, b, st) : patch_launch<P, 1, 8, AFF>(a, b, st);
  return qw == 2 ? patch_launch<P, 2, 12, AFF>(a, b, st) : patch_launch<P, 1, 12, AFF>(a, b, st);
}

// Below ~48 clouds the stand-alone kNN kernel is short (it spreads the few queries over all SMs) and the FPS warps lose
// more to the consumers beside them than the overlap saves (measured, profiles/r02/time_patchify.json: 16 clouds 57 vs
// 48 us, 128 clouds 32 vs 50 us, 256 clouds 61 vs 65 us); enabled = 2 forces the single launch for every eligible shape.
bool patchify_fused_applies(int b, int n, int g, int m) {
  const int mode = patch_tune().enabled;
  return mode != 0 && (b >= 48 || mode == 2) && n >= 512 && n <= 2048 && g >= 1 && g <= 1024 && m >= 1 && m <= 32 && m <= n;
}

int patchify_fused(const float *xyz, int b, int n, int g, int m, int *fps_idx, float *center, int64_t *idx, float *neighborhood,
                   int raw, const GroupAffine *affine, cudaStream_t st, int overlap_previous = 0) {
  PatchArgs a{xyz, fps_idx, center, idx, neighborhood, raw, affine ? *affine : GroupAffine{nullptr, 0, nullptr, nullptr},
              n, g, m, (n + 511) >> 9, 0, 0, overlap_previous, patch_tune().trace};
  const PatchTune &t = patch_tune();
  if (affine) return n <= 1024 ? patch_launch_p<8, true>(a, b, t.qw, t.ncw, st) : patch_launch_p<16, true>(a, b, t.qw, t.ncw, st);
  return n <= 1024 ? patch_launch_p<8, false>(a, b, t.qw, t.ncw, st) : patch_launch_p<16, false>(a, b, t.qw, t.ncw, st);
}

}  // namespace pdae

using namespace pdae;

extern "C" int pdae_tune_patchify(int enabled, int qw, int ncw) {
  PatchTune &t = patch_tune();
  t.enabled = enabled, t.qw = qw, t.ncw = ncw;
  return 0;
}

extern "C" int pdae_patchify_trace(long long *device_buffer) {
  patch_tune().trace = device_buffer;
  return 0;
}

extern "C" size_t pdae_fps_group_workspace_bytes(int b, int n, int g, int m) {
  if (patchify_fused_applies(b, n, g, m)) return 0;
  const size_t f = pdae_fps_workspace_bytes(b, n, g), kq = pdae_knn_workspace_bytes(b, n, g, 3, m);
  return f > kq ? f : kq;
}

extern "C" int pdae_fps_group_ex_f32(const float *xyz, int b, int n, int g, int m, int *fps_idx, float *center, int64_t *idx,
                                     float *neighborhood, void *workspace, size_t workspace_bytes, unsigned flags,
                                     pdae_stream_t stream) {
  if (b < 0 || n < 0 || g < 0 || m <= 0 || (flags & ~static_cast<unsigned>(PDAE_LAUNCH_OVERLAP_PREVIOUS))) return PDAE_E_INVALID;
  if (b == 0 || g == 0) return 0;
  if (n == 0 || m > n) return PDAE_E_INVALID;
  if (!xyz || !fps_idx || !center || !neighborhood) return PDAE_E_INVALID;
  if (patchify_fused_applies(b, n, g, m))
    return patchify_fused(xyz, b, n, g, m, fps_idx, center, idx, neighborhood, 0, nullptr, static_cast<cudaStream_t>(stream),
                          (flags & PDAE_LAUNCH_OVERLAP_PREVIOUS) ? 1 : 0);
  const int rc = pdae_fps_gather_f32(xyz, b, n, 3, g, fps_idx, center, workspace, workspace_bytes, stream);
  if (rc) return rc;
  return pdae_group_ws_f32(xyz, center, b, n, g, m, idx, neighborhood, workspace, workspace_bytes, stream);
}

extern "C" int pdae_fps_group_f32(const float *xyz, int b, int n, int g, int m, int *fps_idx, float *center, int64_t *idx,
                                  float *neighborhood, void *workspace, size_t workspace_bytes, pdae_stream_t stream) {
  return pdae_fps_group_ex_f32(xyz, b, n, g, m, fps_idx, center, idx, neighborhood, workspace, workspace_bytes, 0u, stream);
}

extern "C" int pdae_fps_group_affine_f32(const float *xyz, const float *mats, int b, int n, int g, int m, int t, int *fps_idx,
                                         float *center, int64_t *idx, float *neighborhood, float *t_neighborhood, float *t_center,
                                         void *workspace, size_t workspace_bytes, pdae_stream_t stream) {
  if (b < 0 || n < 0 || g < 0 || m <= 0 || t < 0 || t > PDAE_AFFINE_MAX_CHAIN) return PDAE_E_INVALID;
  if (b == 0 || g == 0) return 0;
  if (n == 0 || m > n) return PDAE_E_INVALID;
  if (!xyz || !fps_idx || !center || !neighborhood || !t_neighborhood || !t_center || (t > 0 && !mats)) return PDAE_E_INVALID;
  if (patchify_fused_applies(b, n, g, m)) {
    // t == 0 is the identity chain: point at any valid address, never dereferenced
    const GroupAffine aff{mats ? mats : xyz, t, t_neighborhood, t_center};
    return patchify_fused(xyz, b, n, g, m, fps_idx, center, idx, neighborhood, 0, &aff, static_cast<cudaStream_t>(stream));
  }
  const int rc = pdae_fps_gather_f32(xyz, b, n, 3, g, fps_idx, center, workspace, workspace_bytes, stream);
  if (rc) return rc;
  return pdae_group_affine_f32(xyz, center, mats, b, n, g, m, t, idx, neighborhood, t_neighborhood, t_center, stream);
}
