// Integer segment reduce over a dst-sorted plan: torch_scatter.scatter on int32 / int64 values.
//
// PyG calls scatter_add on integer tensors for bookkeeping — TopKPooling / to_dense_batch count
// nodes per graph with scatter_add(batch.new_ones(n), batch, dim=0)
// (graph_benchmark/models/ptg_models.py:165-172 -> GraphUNet -> TopKPooling) — so the sum path
// needs an exact integer instantiation: int64 accumulation, no float round trip.  These inputs are
// small index vectors, not feature matrices: one warp per destination row when the row is a
// scalar (K == 1; lanes stride the row's edges, shuffle-reduce), one thread per (row, column)
// otherwise.  Integer arithmetic is associative, so the result is exact and deterministic in any
// order; MIN/MAX report the lowest edge position among equal values like the float path; MEAN is
// the floor division upstream applies to integer tensors (div_(count, rounding_mode='floor')).
#include <climits>

#include "common.cuh"

namespace gno {

template <typename T>
struct IntLimits;
template <>
struct IntLimits<int32_t> {
  __device__ static long long lo() { return INT_MIN; }
  __device__ static long long hi() { return INT_MAX; }
};
template <>
struct IntLimits<int64_t> {
  __device__ static long long lo() { return LLONG_MIN; }
  __device__ static long long hi() { return LLONG_MAX; }
};

__device__ __forceinline__ long long floor_div(long long a, long long b) {  // b > 0
  long long q = a / b;
  if ((a % b != 0) && (a < 0)) --q;
  return q;
}

struct IntSegParams {
  const int64_t* rowptr;
  const int32_t* gidx;  // row of x for sorted edge k (NULL = k)
  const int32_t* eid;   // position reported by arg (NULL = k)
  const void* x;
  void* out;
  int64_t* arg;
  int64_t N, K, ldx, ldo, arg_fill;
  int reduce, accumulate;
};

template <typename T, int RED>
__device__ __forceinline__ void int_combine(long long& a, long long& ae, long long v, long long e) {
  if (RED == GNO_SUM || RED == GNO_MEAN) a += v;
  else if (RED == GNO_MUL) a *= v;
  else if (RED == GNO_MAX) { if (v > a || (v == a && e < ae)) { a = v; ae = e; } }
  else { if (v < a || (v == a && e < ae)) { a = v; ae = e; } }
}

template <typename T, int RED>
__device__ __forceinline__ long long int_init() {
  if (RED == GNO_SUM || RED == GNO_MEAN) return 0;
  if (RED == GNO_MUL) return 1;
  return RED == GNO_MAX ? IntLimits<T>::lo() : IntLimits<T>::hi();
}

template <typename T, int RED>
__device__ __forceinline__ void int_store(const IntSegParams& p, int64_t row, int64_t col, long long a,
                                          long long ae, int64_t len) {
  T* o = static_cast<T*>(p.out) + row * p.ldo + col;
  if (RED == GNO_SUM) {
    if (p.accumulate) a += (long long)*o;
  } else if (RED == GNO_MEAN) {
    if (p.accumulate) a += (long long)*o;
    a = floor_div(a, len > 1 ? len : 1);
  } else if (RED == GNO_MUL) {
    if (p.accumulate) a *= (long long)*o;
  } else {
    // strict compare against the init, as upstream: a value equal to the init never wins
    const bool none = (ae == LLONG_MAX) || a == int_init<T, RED>();
    if (none) a = 0;
    if (p.arg) p.arg[row * p.K + col] = none ? p.arg_fill : ae;
  }
  *o = (T)a;
}

// K == 1: one warp per row.
template <typename T, int RED>
__global__ void __launch_bounds__(256) int_seg_rowwarp_kernel(const IntSegParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); row < p.N; row += warps) {
    const int64_t kb = p.rowptr[row], ke = p.rowptr[row + 1];
    long long a = int_init<T, RED>(), ae = LLONG_MAX;
    for (int64_t k = kb + lane; k < ke; k += 32) {
      const int64_t g = p.gidx ? (int64_t)p.gidx[k] : k;
      const long long v = (long long)static_cast<const T*>(p.x)[g * p.ldx];
      const long long e = p.eid ? (long long)p.eid[k] : (long long)k;
      if ((RED == GNO_MAX || RED == GNO_MIN) && v == int_init<T, RED>()) continue;
      int_combine<T, RED>(a, ae, v, e);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const long long v2 = __shfl_xor_sync(0xffffffffu, a, o);
      const long long e2 = __shfl_xor_sync(0xffffffffu, ae, o);
      if (RED == GNO_SUM || RED == GNO_MEAN) a += v2;
      else if (RED == GNO_MUL) a *= v2;
      else if (e2 != LLONG_MAX) int_combine<T, RED>(a, ae, v2, e2);
    }
    if (lane == 0) int_store<T, RED>(p, row, 0, a, ae, ke - kb);
  }
}

// K > 1: one thread per (row, column), edges in sorted (= ascending position) order.
template <typename T, int RED>
__global__ void __launch_bounds__(256) int_seg_elem_kernel(const IntSegParams p) {
  const int64_t total = p.N * p.K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / p.K, col = i - row * p.K;
    const int64_t kb = p.rowptr[row], ke = p.rowptr[row + 1];
    long long a = int_init<T, RED>(), ae = LLONG_MAX;
    for (int64_t k = kb; k < ke; ++k) {
      const int64_t g = p.gidx ? (int64_t)p.gidx[k] : k;
      const long long v = (long long)static_cast<const T*>(p.x)[g * p.ldx + col];
      const long long e = p.eid ? (long long)p.eid[k] : (long long)k;
      if ((RED == GNO_MAX || RED == GNO_MIN) && v == int_init<T, RED>()) continue;
      int_combine<T, RED>(a, ae, v, e);
    }
    int_store<T, RED>(p, row, col, a, ae, ke - kb);
  }
}

template <typename T, int RED>
static int int_launch(const IntSegParams& p, cudaStream_t s) {
  if (p.K == 1) {
    const int64_t blocks = imin64(ceil_div(p.N * 32, 256), (int64_t)kNumSMs * 16);
    int_seg_rowwarp_kernel<T, RED><<<(unsigned)blocks, 256, 0, s>>>(p);
  } else {
    const int64_t blocks = imin64(ceil_div(p.N * p.K, 256), (int64_t)kNumSMs * 16);
    int_seg_elem_kernel<T, RED><<<(unsigned)blocks, 256, 0, s>>>(p);
  }
  GNO_LAUNCHED("int_seg_kernel");
  return GNO_OK;
}

template <typename T>
static int int_dispatch(const IntSegParams& p, cudaStream_t s) {
  switch (p.reduce) {
    case GNO_SUM: return int_launch<T, GNO_SUM>(p, s);
    case GNO_MEAN: return int_launch<T, GNO_MEAN>(p, s);
    case GNO_MUL: return int_launch<T, GNO_MUL>(p, s);
    case GNO_MIN: return int_launch<T, GNO_MIN>(p, s);
    case GNO_MAX: return int_launch<T, GNO_MAX>(p, s);
  }
  return fail(GNO_ERR_INVALID, "gno_segment_reduce_int: unknown reduce %d", p.reduce);
}

}  // namespace gno

using namespace gno;

extern "C" {

int gno_segment_reduce_int(const gno_csr* g, const void* x, int64_t ldx, void* out, int64_t ldo,
                           int64_t* arg, int64_t arg_fill, int64_t K, int elem_bytes, int reduce,
                           int accumulate, gno_stream_t stream) {
  GNO_CHECK_ARG(g != nullptr, "gno_segment_reduce_int: graph is NULL");
  GNO_CHECK_ARG(elem_bytes == 4 || elem_bytes == 8, "gno_segment_reduce_int: elem_bytes must be 4 (int32) or 8 (int64)");
  GNO_CHECK_ARG(reduce >= GNO_SUM && reduce <= GNO_MAX, "gno_segment_reduce_int: unknown reduce %d", reduce);
  GNO_CHECK_ARG(K >= 0 && ldx >= K && ldo >= K, "gno_segment_reduce_int: bad sizes");
  GNO_CHECK_ARG(arg == nullptr || reduce == GNO_MIN || reduce == GNO_MAX,
                "gno_segment_reduce_int: arg output only for MIN/MAX");
  GNO_CHECK_ARG(!accumulate || reduce == GNO_SUM || reduce == GNO_MUL || reduce == GNO_MEAN,
                "gno_segment_reduce_int: accumulate only for SUM/MEAN/MUL");
  if (g->N == 0 || K == 0) return GNO_OK;
  GNO_CHECK_ARG(g->rowptr && out && (g->E == 0 || x), "gno_segment_reduce_int: NULL buffer");
  IntSegParams p;
  p.rowptr = g->rowptr;
  p.gidx = g->gidx;
  p.eid = g->eid;
  p.x = x;
  p.out = out;
  p.arg = arg;
  p.N = g->N;
  p.K = K;
  p.ldx = ldx;
  p.ldo = ldo;
  p.arg_fill = arg_fill;
  p.reduce = reduce;
  p.accumulate = accumulate ? 1 : 0;
  return elem_bytes == 4 ? int_dispatch<int32_t>(p, (cudaStream_t)stream)
                         : int_dispatch<int64_t>(p, (cudaStream_t)stream);
}

}  // extern "C"
