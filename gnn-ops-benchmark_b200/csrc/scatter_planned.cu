// torch_scatter.scatter with a FULL-SHAPE index, on a cached plan: atomic-free and deterministic
// for every dtype and reduce.
//
// scatter_onchip.cu needs no preparation but pays one shared-memory atomic per element, and the
// LSU retires those at about two cycles per lane: 45 M elements cannot finish under ~0.28 ms on
// 148 SMs whatever the memory system does (profiles/r2b_ops_c1.jsonl: 0.28-0.71 ms for the
// reference's (6708, 6708) fp16 case, 12-30 % of the HBM roofline).  The reference scripts call the
// op in a loop on the SAME index tensor (op_bm_scripts/benchmark_scatter_add.py:97-118,
// timeit(100)), and so does a GNN layer, so — exactly as for a 1-D index — the index is sorted
// once into a plan that is cached on the tensor's identity:
//     order [B*E*K] int32   position e (along the scatter dim) of the j-th element in OUTPUT order
//     ptr   [B*N*K+1] int32 CSR offsets: output element o = (b*N+n)*K+k owns order[ptr[o]:ptr[o+1])
// (a stable sort, so each segment lists its elements by ascending e).  Per call a CTA stages the KB
// adjacent source columns src[b, :, k0:k0+KB] in shared memory (coalesced), and every thread
// reduces the segments of its outputs sequentially from shared memory: no atomics, fp32
// accumulation in ascending e (the order of upstream's sequential CPU loop), strict compares, so
// MIN/MAX ties resolve to the lowest position.  One launch.
//
// Roofline: HBM.  Bytes per element (N = E): s (src) + 4 (order) + 4 (ptr) + s (out) [+ 8 arg].
#include "common.cuh"

namespace gno {

struct PlannedParams {
  const void* src;
  const int32_t* order;
  const int32_t* ptr;
  void* out;
  int64_t* arg;
  int64_t B, E, K, N;
  int kb_shift;  // columns per CTA = 1 << kb_shift
  int reduce;
  int accumulate;
};

template <typename T, int RED>
__global__ void __launch_bounds__(512) scatter_planned_kernel(const PlannedParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* tile = reinterpret_cast<T*>(smem_raw);  // [E][KB]
  const int KB = 1 << p.kb_shift;
  const int64_t ncb = (p.K + KB - 1) >> p.kb_shift;
  const int64_t b = blockIdx.x / ncb, cb = blockIdx.x - b * ncb;
  const int64_t k0 = cb << p.kb_shift;
  const int kw = (int)imin64(KB, p.K - k0);
  const T* s = static_cast<const T*>(p.src) + b * p.E * p.K;

  // stage the column block (rows of kw contiguous elements)
  const int64_t n_tile = p.E << p.kb_shift;
  for (int64_t i = threadIdx.x; i < n_tile; i += blockDim.x) {
    const int kk = (int)(i & (KB - 1));
    const int64_t e = i >> p.kb_shift;
    if (kk < kw) tile[i] = s[e * p.K + k0 + kk];
  }
  __syncthreads();

  T* o = static_cast<T*>(p.out) + b * p.N * p.K;
  int64_t* a = p.arg ? p.arg + b * p.N * p.K : nullptr;
  const int32_t* ptr = p.ptr + b * p.N * p.K;
  const int64_t n_out = p.N << p.kb_shift;
  // kPlU outputs per thread at a time: their pointer pairs, then the r-th element of each
  // segment, are independent loads — the chain ptr -> order -> shared memory -> store would
  // otherwise run once per output at full global latency (segments hold ~1 element when N = E)
  constexpr int kPlU = 4;
  for (int64_t base = threadIdx.x; base < n_out; base += (int64_t)blockDim.x * kPlU) {
    int32_t lo[kPlU], hi[kPlU], win[kPlU];
    int64_t oo[kPlU];
    float acc[kPlU];
    T wv[kPlU];
    int kk[kPlU];
    int32_t longest = 0;
#pragma unroll
    for (int u = 0; u < kPlU; ++u) {
      const int64_t i = base + (int64_t)u * blockDim.x;
      kk[u] = (int)(i & (KB - 1));
      const bool live = i < n_out && kk[u] < kw;
      oo[u] = (i >> p.kb_shift) * p.K + k0 + kk[u];
      lo[u] = live ? __ldg(ptr + oo[u]) : 0;
      hi[u] = live ? __ldg(ptr + oo[u] + 1) : -1;   // hi < lo marks a dead slot
      if (RED == GNO_SUM || RED == GNO_MEAN) acc[u] = 0.f;
      else if (RED == GNO_MUL) acc[u] = 1.f;
      else acc[u] = RED == GNO_MAX ? DType<T>::lowest() : DType<T>::highest();
      win[u] = -1;
      wv[u] = DType<T>::from_f(0.f);
    }
#pragma unroll
    for (int u = 0; u < kPlU; ++u) longest = max(longest, hi[u] - lo[u]);
    for (int32_t r = 0; r < longest; ++r) {
      int32_t e[kPlU];
#pragma unroll
      for (int u = 0; u < kPlU; ++u) e[u] = (lo[u] + r < hi[u]) ? __ldg(p.order + lo[u] + r) : -1;
#pragma unroll
      for (int u = 0; u < kPlU; ++u) {
        if (e[u] < 0) continue;
        const T tv = tile[((int64_t)e[u] << p.kb_shift) + kk[u]];
        const float v = DType<T>::to_f(tv);
        if (RED == GNO_SUM || RED == GNO_MEAN) acc[u] += v;
        else if (RED == GNO_MUL) acc[u] *= v;
        else if (RED == GNO_MAX) { if (v > acc[u]) { acc[u] = v; win[u] = e[u]; wv[u] = tv; } }
        else { if (v < acc[u]) { acc[u] = v; win[u] = e[u]; wv[u] = tv; } }
      }
    }
#pragma unroll
    for (int u = 0; u < kPlU; ++u) {
      if (hi[u] < lo[u]) continue;
      const int64_t q = oo[u];
      if (RED == GNO_SUM || RED == GNO_MEAN) {
        float r = acc[u];
        if (p.accumulate) r += DType<T>::to_f(o[q]);
        if (RED == GNO_MEAN) r = r / (float)(hi[u] - lo[u] > 1 ? hi[u] - lo[u] : 1);
        o[q] = DType<T>::from_f(r);
      } else if (RED == GNO_MUL) {
        float r = acc[u];
        if (p.accumulate) r *= DType<T>::to_f(o[q]);
        o[q] = DType<T>::from_f(r);
      } else if (p.accumulate) {
        // out= form: the existing value survives (arg = E) unless an element beats it strictly
        const float prev = DType<T>::to_f(o[q]);
        const bool beat = win[u] >= 0 && ((RED == GNO_MAX) ? (acc[u] > prev) : (acc[u] < prev));
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
constexpr int64_t kPlannedSmem = 110 * 1024;

template <typename T>
static int planned_dispatch(const PlannedParams& p, int64_t blocks, size_t smem, cudaStream_t s) {
#define GNO_PLANNED(R)                                                                              \
  {                                                                                                 \
    auto k = scatter_planned_kernel<T, R>;                                                          \
    if (smem > 48 * 1024)                                                                           \
      GNO_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPlannedSmem)); \
    k<<<(unsigned)blocks, 512, smem, s>>>(p);                                                       \
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

// columns per CTA for a [B, E, K] source of es-byte elements; -1 when one column does not fit
static int planned_kb_shift(int64_t E, int64_t K, int es) {
  int sh = 4;
  while (sh > 0 && ((int64_t(1) << sh) > K * 2 - 1 || (E << sh) * es > kPlannedSmem)) --sh;
  return (E << sh) * es > kPlannedSmem ? -1 : sh;
}

}  // namespace gno

using namespace gno;

extern "C" {

int gno_scatter_planned_ok(int64_t B, int64_t E, int64_t K, int64_t N, int dtype) {
  if (B <= 0 || E <= 0 || K <= 0 || N <= 0) return 0;
  if (B * E * K >= (int64_t(1) << 31) || B * N * K >= (int64_t(1) << 31) - 1) return 0;
  return planned_kb_shift(E, K, dtype == GNO_F32 ? 4 : 2) >= 0 ? 1 : 0;
}

int gno_scatter_planned(const void* src, const int32_t* order, const int32_t* ptr, int64_t B, int64_t E,
                        int64_t K, void* out, int64_t* arg, int64_t N, int dtype, int reduce,
                        int accumulate, gno_stream_t stream) {
  GNO_CHECK_ARG(B >= 0 && E >= 0 && K >= 0 && N >= 0, "gno_scatter_planned: negative size");
  GNO_CHECK_ARG(reduce >= GNO_SUM && reduce <= GNO_MAX, "gno_scatter_planned: unknown reduce %d", reduce);
  GNO_CHECK_ARG(dtype == GNO_F32 || dtype == GNO_F16 || dtype == GNO_BF16, "gno_scatter_planned: unknown dtype %d", dtype);
  GNO_CHECK_ARG(arg == nullptr || reduce == GNO_MIN || reduce == GNO_MAX, "gno_scatter_planned: arg output only for MIN/MAX");
  if (B * N * K == 0) return GNO_OK;
  GNO_CHECK_ARG(out && ptr && (B * E * K == 0 || (src && order)), "gno_scatter_planned: NULL buffer");
  if (!gno_scatter_planned_ok(B, E > 0 ? E : 1, K, N, dtype))
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
  p.kb_shift = planned_kb_shift(E > 0 ? E : 1, K, es);
  p.reduce = reduce;
  p.accumulate = accumulate ? 1 : 0;
  const int64_t ncb = (K + (int64_t(1) << p.kb_shift) - 1) >> p.kb_shift;
  const size_t smem = (size_t)((E << p.kb_shift) * es);
  cudaStream_t s = (cudaStream_t)stream;
  switch (dtype) {
    case GNO_F32: return planned_dispatch<float>(p, B * ncb, smem, s);
    case GNO_F16: return planned_dispatch<__half>(p, B * ncb, smem, s);
    default: return planned_dispatch<__nv_bfloat16>(p, B * ncb, smem, s);
  }
}

}  // extern "C"
