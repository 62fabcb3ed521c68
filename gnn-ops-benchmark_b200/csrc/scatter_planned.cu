// torch_scatter.scatter with a FULL-SHAPE index, on a cached plan: atomic-free and deterministic
// for every dtype and reduce.
//
// scatter_onchip.cu needs no preparation but pays one shared-memory atomic per element, and the
// LSU retires those at about two cycles per lane: 45 M elements cannot finish under ~0.28 ms on
// 148 SMs whatever the memory system does (profiles/r2b_ops_c1.jsonl: 0.28-0.71 ms for the
// reference's (6708, 6708) fp16 case, 12-30 % of the HBM roofline).  The reference scripts call the
// op in a loop on the SAME index tensor (op_bm_scripts/benchmark_scatter_add.py:97-118,
// timeit(100)), and so does a GNN layer, so — exactly as for a 1-D index — the index is sorted
// once into a plan that is cached on the tensor's identity.
//
// Plan layout (r2u): BLOCKED by the CTA that consumes it.  src / index are [B, E, K], out is
// [B, N, K]; CTA (b, cb) owns the KB = 1 << kb_shift adjacent columns k0 = cb*KB ... and the
// blocked output id of (b, n, k = k0 + kk) is
//     ob = (((b*ncb + cb)*N + n) << kb_shift) + kk          ncb = ceil(K / KB)
//     ptr   [B*ncb*N*KB + 1] int32   output ob owns order[ptr[ob] : ptr[ob+1])
//     order [B*E*K]  int16 (E <= 32768) or int32: position e along the scatter dim, elements in
//           ascending (ob, e)  (a stable sort of ob)
// so a CTA reads ONE contiguous slice of ptr and ONE contiguous slice of order (the first version
// kept both in natural (b, n, k) order: a CTA's share was 16-byte strips, one per output row, and
// half of every sector it fetched belonged to a neighbour).  Per call a CTA stages its KB source
// columns src[b, :, k0:k0+KB] in shared memory with 8/16-byte loads, and every thread reduces the
// segments of kPlU outputs at a time from shared memory: no atomics, fp32 accumulation in ascending
// e (the order of upstream's sequential CPU loop), strict compares, so MIN/MAX ties resolve to the
// lowest position.  One launch.
//
// Roofline: HBM.  Bytes per element (N = E): s (src) + 2|4 (order) + 4 (ptr) + s (out) [+ 8 arg].
#include "common.cuh"

namespace gno {

struct PlannedParams {
  const void* src;
  const void* order;
  const int32_t* ptr;
  void* out;
  int64_t* arg;
  int64_t B, E, K, N;
  int kb_shift;  // columns per CTA = 1 << kb_shift
  int vec;       // bytes per staging load (16, 8 or 0 = element-wise)
  int reduce;
  int accumulate;
};

// kPlThreads = 512, two CTAs per SM (tiles up to 110 KB), or 1024 with one CTA per SM (a tile up to
// 220 KB: twice the columns, so twice as wide strips in src / out / arg)
template <typename T, typename OT, int RED, int kPlThreads>
__global__ void __launch_bounds__(kPlThreads, 2048 / kPlThreads / 2) scatter_planned_kernel(const PlannedParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* tile = reinterpret_cast<T*>(smem_raw);  // [E][KB]
  const int KB = 1 << p.kb_shift;
  const int64_t ncb = (p.K + KB - 1) >> p.kb_shift;
  const int64_t b = blockIdx.x / ncb, cb = blockIdx.x - b * ncb;
  const int64_t k0 = cb << p.kb_shift;
  const int kw = (int)imin64(KB, p.K - k0);
  const T* s = static_cast<const T*>(p.src) + b * p.E * p.K;
  const OT* order = static_cast<const OT*>(p.order);

  T* o = static_cast<T*>(p.out) + b * p.N * p.K;
  int64_t* a = p.arg ? p.arg + b * p.N * p.K : nullptr;
  const int64_t n_out = p.N << p.kb_shift;             // blocked outputs of this CTA (dead kk >= kw included)
  const int32_t* ptr = p.ptr + (int64_t)blockIdx.x * n_out;
  const bool ptr_vec = (reinterpret_cast<uintptr_t>(ptr) & 15) == 0;
  constexpr int kPlU = 8;
  // pointers of the kPlU consecutive outputs starting at i0 (and the end of the last one)
  auto load_ptrs = [&](int64_t i0, int32_t (&pp)[kPlU + 1]) {
    if (ptr_vec && i0 + kPlU <= n_out) {
      const int4 v0 = __ldg(reinterpret_cast<const int4*>(ptr + i0));
      const int4 v1 = __ldg(reinterpret_cast<const int4*>(ptr + i0) + 1);
      pp[0] = v0.x; pp[1] = v0.y; pp[2] = v0.z; pp[3] = v0.w;
      pp[4] = v1.x; pp[5] = v1.y; pp[6] = v1.z; pp[7] = v1.w;
      pp[8] = __ldg(ptr + i0 + kPlU);
    } else if (i0 < n_out) {
#pragma unroll
      for (int u = 0; u <= kPlU; ++u) pp[u] = __ldg(ptr + imin64(i0 + u, n_out));
    } else {
#pragma unroll
      for (int u = 0; u <= kPlU; ++u) pp[u] = 0;
    }
  };
  // the first two iterations' pointers do not depend on the tile: in flight while it is staged
  const int64_t istep = (int64_t)kPlThreads * kPlU;
  int32_t cur[kPlU + 1], nxt[kPlU + 1];
  load_ptrs((int64_t)threadIdx.x * kPlU, cur);
  load_ptrs((int64_t)threadIdx.x * kPlU + istep, nxt);

  // stage the column block: row e of the tile = src[b, e, k0 : k0+KB); kStU loads in flight per thread
  constexpr int kStU = 8, kStU16 = 4;
  if (p.vec == 16 && kw == KB) {
    const int wsh = p.kb_shift + (sizeof(T) == 4 ? 2 : 1) - 4;  // log2(16-byte words per tile row)
    const int64_t n_w = p.E << wsh;
    for (int64_t i = threadIdx.x; i < n_w; i += (int64_t)kPlThreads * kStU16) {
      uint4 v[kStU16];
#pragma unroll
      for (int u = 0; u < kStU16; ++u) {
        const int64_t iu = i + (int64_t)u * kPlThreads;
        if (iu < n_w) v[u] = __ldg(reinterpret_cast<const uint4*>(s + (iu >> wsh) * p.K + k0) + (iu & ((1 << wsh) - 1)));
      }
#pragma unroll
      for (int u = 0; u < kStU16; ++u) {
        const int64_t iu = i + (int64_t)u * kPlThreads;
        if (iu < n_w) reinterpret_cast<uint4*>(tile)[iu] = v[u];
      }
    }
  } else if (p.vec >= 8 && kw == KB) {
    const int wsh = p.kb_shift + (sizeof(T) == 4 ? 2 : 1) - 3;  // log2(8-byte words per tile row)
    const int64_t n_w = p.E << wsh;
    for (int64_t i = threadIdx.x; i < n_w; i += (int64_t)kPlThreads * kStU) {
      uint2 v[kStU];
#pragma unroll
      for (int u = 0; u < kStU; ++u) {
        const int64_t iu = i + (int64_t)u * kPlThreads;
        if (iu < n_w) v[u] = __ldg(reinterpret_cast<const uint2*>(s + (iu >> wsh) * p.K + k0) + (iu & ((1 << wsh) - 1)));
      }
#pragma unroll
      for (int u = 0; u < kStU; ++u) {
        const int64_t iu = i + (int64_t)u * kPlThreads;
        if (iu < n_w) reinterpret_cast<uint2*>(tile)[iu] = v[u];
      }
    }
  } else {
    const int64_t n_tile = p.E << p.kb_shift;
    for (int64_t i = threadIdx.x; i < n_tile; i += (int64_t)kPlThreads * kStU) {
      T v[kStU];
#pragma unroll
      for (int u = 0; u < kStU; ++u) {
        const int64_t iu = i + (int64_t)u * kPlThreads;
        if (iu < n_tile && (int)(iu & (KB - 1)) < kw) v[u] = s[(iu >> p.kb_shift) * p.K + k0 + (iu & (KB - 1))];
      }
#pragma unroll
      for (int u = 0; u < kStU; ++u) {
        const int64_t iu = i + (int64_t)u * kPlThreads;
        if (iu < n_tile && (int)(iu & (KB - 1)) < kw) tile[iu] = v[u];
      }
    }
  }
  __syncthreads();

  // Every thread owns kPlU CONSECUTIVE blocked outputs, hence one contiguous run of `order`
  // (segments hold ~1 element when N = E): its 9 pointers are two 16-byte loads and one scalar
  // (issued two iterations ahead), the run's first element misses to DRAM once and the rest hit
  // the same L1 sector, and with KB >= 8 the outputs are one 16-byte strip of an output row.
  // (The first version gave a thread outputs 512 apart and walked the r-th element of four
  // segments at a time: a warp looped to the LONGEST of its 128 segments, one dependent DRAM
  // latency per round — ncu showed 47 % warps active at 10 long-scoreboard stalls per issue and
  // 1.2 TB/s, profiles/r2t_planned_ncu.txt.)
  for (int64_t i0 = (int64_t)threadIdx.x * kPlU; i0 < n_out; i0 += istep) {
    int32_t pp[kPlU + 1];
#pragma unroll
    for (int u = 0; u <= kPlU; ++u) { pp[u] = cur[u]; cur[u] = nxt[u]; }
    load_ptrs(i0 + 2 * istep, nxt);
    const int64_t n0 = i0 >> p.kb_shift;
    const int kk0 = (int)(i0 & (KB - 1));
    const int64_t q0 = n0 * p.K + k0 + kk0;
    // (K == KB: one column block spans the whole row, blocked order = natural order)
    const bool strip = i0 + kPlU <= n_out && ((KB >= kPlU && kk0 + kPlU <= kw) || p.K == KB);
    if (sizeof(T) == 2 && strip && !p.accumulate &&
        (reinterpret_cast<uintptr_t>(o + q0) & 7) == 0) {
      // Fast path (16-bit data, a full strip of live outputs).  Segments hold c = 0, 1, 2, ...
      // elements (Poisson(1) when N = E), so the first kPlInl elements of every segment are
      // handled by straight-line predicated code — all lanes active, the element loads of four
      // outputs issued together — and only longer segments (8 % at kPlInl = 2, 2 % at 3) enter a
      // loop.  A loop per output ran, warp-wide, to the longest of 32 segments with a quarter of
      // the lanes active (ncu: 7.9 threads per warp, 109 thread instructions per output,
      // profiles/r2v_planned_ncu.txt); one flat loop over the thread's elements with bit-mask
      // bookkeeping was no better (64-bit variable shifts: ~100 instructions per element, r2w).
      constexpr int kPlInl = 3;
      const unsigned short* t16 = reinterpret_cast<const unsigned short*>(tile);
      float res[kPlU];
      int win[kPlU];
      unsigned short wb[kPlU];
#pragma unroll
      for (int hh = 0; hh < kPlU; hh += 4) {
        int e[4][kPlInl];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int c = pp[hh + v + 1] - pp[hh + v];
#pragma unroll
          for (int k = 0; k < kPlInl; ++k) e[v][k] = (c > k) ? (int)__ldg(order + pp[hh + v] + k) : -1;
        }
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int u = hh + v;
          const int kk = (kk0 + u) & (KB - 1);
          float acc;
          if (RED == GNO_SUM || RED == GNO_MEAN) acc = 0.f;
          else if (RED == GNO_MUL) acc = 1.f;
          else acc = RED == GNO_MAX ? DType<T>::lowest() : DType<T>::highest();
          int w1 = -1;
          unsigned short wbits = 0;
          // branch-free for the inline elements (absent ones read tile row 0 and are discarded by
          // selects: the compiler turned `if (present) ...` into divergent branches, 1400 SASS
          // instructions per iteration, r2x)
#pragma unroll
          for (int k = 0; k < kPlInl; ++k) {
            const int ee = e[v][k];
            const bool have = ee >= 0;
            const unsigned short tb = t16[((have ? ee : 0) << p.kb_shift) + kk];
            T tv;
            *reinterpret_cast<unsigned short*>(&tv) = tb;
            const float val = DType<T>::to_f(tv);
            if (RED == GNO_SUM || RED == GNO_MEAN) acc += have ? val : 0.f;
            else if (RED == GNO_MUL) acc *= have ? val : 1.f;
            else {
              const bool better = have && (RED == GNO_MAX ? (val > acc) : (val < acc));
              acc = better ? val : acc;
              w1 = better ? ee : w1;
              wbits = better ? tb : wbits;
            }
          }
          if (pp[u + 1] - pp[u] > kPlInl) {
            for (int32_t j = pp[u] + kPlInl; j < pp[u + 1]; ++j) {
              const int ee = (int)__ldg(order + j);
              const unsigned short tb = t16[(ee << p.kb_shift) + kk];
              T tv;
              *reinterpret_cast<unsigned short*>(&tv) = tb;
              const float val = DType<T>::to_f(tv);
              if (RED == GNO_SUM || RED == GNO_MEAN) acc += val;
              else if (RED == GNO_MUL) acc *= val;
              else if (RED == GNO_MAX) { if (val > acc) { acc = val; w1 = ee; wbits = tb; } }
              else { if (val < acc) { acc = val; w1 = ee; wbits = tb; } }
            }
          }
          res[u] = acc;
          win[u] = w1;
          wb[u] = wbits;
        }
      }
      unsigned long long w[2] = {0ull, 0ull};   // outputs 0-3, 4-7
#pragma unroll
      for (int u = 0; u < kPlU; ++u) {
        unsigned short rb;
        if (RED == GNO_MIN || RED == GNO_MAX) {
          rb = wb[u];   // the winner's own bits; empty or all-NaN segments: 0, arg = E
        } else {
          float r = res[u];
          if (RED == GNO_MEAN) {
            const int c = pp[u + 1] - pp[u];
            if (c > 1) r = r / (float)c;
          }
          const T rt = DType<T>::from_f(r);
          rb = *reinterpret_cast<const unsigned short*>(&rt);
        }
        w[u >> 2] |= (unsigned long long)rb << ((u & 3) * 16);
      }
      unsigned long long* dst = reinterpret_cast<unsigned long long*>(o + q0);
      dst[0] = w[0];
      dst[1] = w[1];
      if ((RED == GNO_MIN || RED == GNO_MAX) && a) {
        if ((reinterpret_cast<uintptr_t>(a + q0) & 15) == 0) {
#pragma unroll
          for (int u = 0; u < kPlU; u += 2) {
            longlong2 av;
            av.x = win[u] >= 0 ? (int64_t)win[u] : p.E;
            av.y = win[u + 1] >= 0 ? (int64_t)win[u + 1] : p.E;
            *reinterpret_cast<longlong2*>(a + q0 + u) = av;
          }
        } else {
#pragma unroll
          for (int u = 0; u < kPlU; ++u) a[q0 + u] = win[u] >= 0 ? (int64_t)win[u] : p.E;
        }
      }
      continue;
    }
    float res[kPlU];
    int32_t win[kPlU];
    T wv[kPlU];
#pragma unroll
    for (int u = 0; u < kPlU; ++u) {
      const int kk = (int)((i0 + u) & (KB - 1));
      float acc;
      if (RED == GNO_SUM || RED == GNO_MEAN) acc = 0.f;
      else if (RED == GNO_MUL) acc = 1.f;
      else acc = RED == GNO_MAX ? DType<T>::lowest() : DType<T>::highest();
      win[u] = -1;
      wv[u] = DType<T>::from_f(0.f);
      for (int32_t j = pp[u]; j < pp[u + 1]; ++j) {
        const int32_t e = (int32_t)__ldg(order + j);
        const T tv = tile[((int64_t)e << p.kb_shift) + kk];
        const float v = DType<T>::to_f(tv);
        if (RED == GNO_SUM || RED == GNO_MEAN) acc += v;
        else if (RED == GNO_MUL) acc *= v;
        else if (RED == GNO_MAX) { if (v > acc) { acc = v; win[u] = e; wv[u] = tv; } }
        else { if (v < acc) { acc = v; win[u] = e; wv[u] = tv; } }
      }
      res[u] = acc;
    }
#pragma unroll
    for (int u = 0; u < kPlU; ++u) {
      const int64_t i = i0 + u;
      const int kk = (int)(i & (KB - 1));
      if (i >= n_out || kk >= kw) continue;
      const int64_t q = (i >> p.kb_shift) * p.K + k0 + kk;
      if (RED == GNO_SUM || RED == GNO_MEAN) {
        float r = res[u];
        if (p.accumulate) r += DType<T>::to_f(o[q]);
        if (RED == GNO_MEAN) {
          const int c = pp[u + 1] - pp[u];
          r = r / (float)(c > 1 ? c : 1);
        }
        o[q] = DType<T>::from_f(r);
      } else if (RED == GNO_MUL) {
        float r = res[u];
        if (p.accumulate) r *= DType<T>::to_f(o[q]);
        o[q] = DType<T>::from_f(r);
      } else if (p.accumulate) {
        // out= form: the existing value survives (arg = E) unless an element beats it strictly
        const float prev = DType<T>::to_f(o[q]);
        const bool beat = win[u] >= 0 && ((RED == GNO_MAX) ? (res[u] > prev) : (res[u] < prev));
        if (beat) o[q] = wv[u];
        if (a) a[q] = beat ? (int64_t)win[u] : p.E;
      } else {
        o[q] = win[u] >= 0 ? wv[u] : DType<T>::from_f(0.f);  // the winner's own bits (-0.0 stays -0.0)
        if (a) a[q] = win[u] >= 0 ? (int64_t)win[u] : p.E;
      }
    }
  }
}

// tile budget: two CTAs of 512 threads per SM (latency hiding matters more than tile width)
constexpr int64_t kPlannedSmemSmall = 110 * 1024, kPlannedSmemBig = 220 * 1024;
static int64_t planned_smem_budget() {
  // default: a tile of up to 220 KB (one 1024-thread CTA per SM when it exceeds 110 KB) — on the
  // reference's (6708, 6708) fp16 dim-0 call 16 columns per CTA instead of 8: 0.238 vs 0.291 ms
  // (sum), 0.440 vs 0.578 ms (max+arg), profiles/r2z_ops_c1_wide.jsonl.  GNO_PLANNED_WIDE=0: 110 KB.
  static const int64_t v = (getenv("GNO_PLANNED_WIDE") && !atoi(getenv("GNO_PLANNED_WIDE"))) ? kPlannedSmemSmall
                                                                                              : kPlannedSmemBig;
  return v;
}

template <typename T, typename OT>
static int planned_dispatch2(const PlannedParams& p, int64_t blocks, size_t smem, cudaStream_t s) {
#define GNO_PLANNED(R)                                                                              \
  {                                                                                                 \
    if (smem > (size_t)kPlannedSmemSmall) {                                                         \
      auto k = scatter_planned_kernel<T, OT, R, 1024>;                                              \
      GNO_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPlannedSmemBig)); \
      k<<<(unsigned)blocks, 1024, smem, s>>>(p);                                                    \
    } else {                                                                                        \
      auto k = scatter_planned_kernel<T, OT, R, 512>;                                               \
      if (smem > 48 * 1024)                                                                         \
        GNO_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPlannedSmemSmall)); \
      k<<<(unsigned)blocks, 512, smem, s>>>(p);                                                     \
    }                                                                                               \
  }                                                                                                 \
  break;
  switch (p.reduce) {
    case GNO_SUM: GNO_PLANNED(GNO_SUM)
    case GNO_MEAN: GNO_PLANNED(GNO_MEAN)
    case GNO_MUL: GNO_PLANNED(GNO_MUL)
    case GNO_MIN: GNO_PLANNED(GNO_MIN)
    case GNO_MAX: GNO_PLANNED(GNO_MAX)
    default: return fail(GNO_ERR_INVALID, "gno_scatter_planned: unknown reduce %d", p.reduce);
  }
#undef GNO_PLANNED
  GNO_LAUNCHED("scatter_planned_kernel");
  return GNO_OK;
}

template <typename T>
static int planned_dispatch(const PlannedParams& p, int order_bytes, int64_t blocks, size_t smem, cudaStream_t s) {
  return order_bytes == 2 ? planned_dispatch2<T, int16_t>(p, blocks, smem, s)
                          : planned_dispatch2<T, int32_t>(p, blocks, smem, s);
}

// columns per CTA for a [B, E, K] source of es-byte elements; -1 when one column does not fit
static int planned_kb_shift(int64_t E, int64_t K, int es) {
  int sh = 4;
  const int64_t budget = planned_smem_budget();
  while (sh > 0 && ((int64_t(1) << sh) > K * 2 - 1 || (E << sh) * es > budget)) --sh;
  return (E << sh) * es > budget ? -1 : sh;
}

}  // namespace gno

using namespace gno;

extern "C" {

int gno_scatter_planned_layout(int64_t B, int64_t E, int64_t K, int64_t N, int dtype, int* kb_shift,
                               int* order_bytes) {
  if (B <= 0 || E <= 0 || K <= 0 || N <= 0) return 0;
  const int sh = planned_kb_shift(E, K, dtype == GNO_F32 ? 4 : 2);
  if (sh < 0) return 0;
  const int64_t ncb = (K + (int64_t(1) << sh) - 1) >> sh;
  if (B * E * K >= (int64_t(1) << 31) || ((B * ncb * N) << sh) >= (int64_t(1) << 31) - 1) return 0;
  if (kb_shift) *kb_shift = sh;
  if (order_bytes) *order_bytes = E <= 32768 ? 2 : 4;
  return 1;
}

int gno_scatter_planned_ok(int64_t B, int64_t E, int64_t K, int64_t N, int dtype) {
  return gno_scatter_planned_layout(B, E, K, N, dtype, nullptr, nullptr);
}

int gno_scatter_planned(const void* src, const void* order, const int32_t* ptr, int64_t B, int64_t E,
                        int64_t K, void* out, int64_t* arg, int64_t N, int dtype, int reduce,
                        int accumulate, gno_stream_t stream) {
  GNO_CHECK_ARG(B >= 0 && E >= 0 && K >= 0 && N >= 0, "gno_scatter_planned: negative size");
  GNO_CHECK_ARG(reduce >= GNO_SUM && reduce <= GNO_MAX, "gno_scatter_planned: unknown reduce %d", reduce);
  GNO_CHECK_ARG(dtype == GNO_F32 || dtype == GNO_F16 || dtype == GNO_BF16, "gno_scatter_planned: unknown dtype %d", dtype);
  GNO_CHECK_ARG(arg == nullptr || reduce == GNO_MIN || reduce == GNO_MAX, "gno_scatter_planned: arg output only for MIN/MAX");
  if (B * N * K == 0) return GNO_OK;
  GNO_CHECK_ARG(out && ptr && (B * E * K == 0 || (src && order)), "gno_scatter_planned: NULL buffer");
  int kb_shift = 0, order_bytes = 4;
  if (!gno_scatter_planned_layout(B, E > 0 ? E : 1, K, N, dtype, &kb_shift, &order_bytes))
    return fail(GNO_ERR_UNSUPPORTED, "gno_scatter_planned: one source column (E=%lld) does not fit in shared memory",
                (long long)E);
  const int es = dtype == GNO_F32 ? 4 : 2;
  PlannedParams p;
  p.src = src;
  p.order = order;
  p.ptr = ptr;
  p.out = out;
  p.arg = arg;
  p.B = B;
  p.E = E;
  p.K = K;
  p.N = N;
  p.kb_shift = kb_shift;
  p.reduce = reduce;
  p.accumulate = accumulate ? 1 : 0;
  // widest staging load every tile row start is aligned to: rows start at (e*K + k0)*es bytes
  const int64_t row_b = K * es, blk_b = (int64_t(1) << kb_shift) * es;
  const uintptr_t a0 = reinterpret_cast<uintptr_t>(src);
  p.vec = (row_b % 16 == 0 && blk_b % 16 == 0 && a0 % 16 == 0) ? 16
          : (row_b % 8 == 0 && blk_b % 8 == 0 && a0 % 8 == 0) ? 8 : 0;
  const int64_t ncb = (K + (int64_t(1) << p.kb_shift) - 1) >> p.kb_shift;
  const size_t smem = (size_t)((E << p.kb_shift) * es);
  cudaStream_t s = (cudaStream_t)stream;
  switch (dtype) {
    case GNO_F32: return planned_dispatch<float>(p, order_bytes, B * ncb, smem, s);
    case GNO_F16: return planned_dispatch<__half>(p, order_bytes, B * ncb, smem, s);
    default: return planned_dispatch<__nv_bfloat16>(p, order_bytes, B * ncb, smem, s);
  }
}

}  // extern "C"
