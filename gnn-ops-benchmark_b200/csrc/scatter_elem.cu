// torch_scatter.scatter with a FULL-SHAPE index (index.shape == src.shape):
// the form op_bm_scripts/benchmark_scatter_{add,max,min,mean}.py actually
// pass.  Every element goes to its own destination, so there is no row
// structure to exploit: this is the compatibility path (L2 atomics), not
// the roofline path.  Determinism: MIN/MAX values and args are exact and
// order-independent (arg = lowest position among equal winners, value =
// src[arg], like the sequential upstream CPU loop); SUM/MEAN/MUL accumulate
// in fp32 and round once.
#include <cstring>

#include "common.cuh"

namespace gno {

// scatter_onchip.cu: one-launch shared-memory path (returns -1 when the shape does not fit it)
int scatter_onchip(const void* src, const int64_t* index, int64_t B, int64_t E, int64_t K, void* out,
                   int64_t* arg, int64_t N, int dtype, int reduce, int accumulate, cudaStream_t s);
bool scatter_onchip_ok(int64_t B, int64_t E, int64_t K, int64_t N, int dtype, int reduce);

static unsigned grid_for_elems(int64_t n) {
  int64_t b = ceil_div(n, 256);
  int64_t cap = (int64_t)kNumSMs * 32;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

#define GNO_GRID_STRIDE(i, n)                                                   \
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n);     \
       i += (int64_t)gridDim.x * blockDim.x)

template <typename V>
__global__ void fill_kernel(V* __restrict__ p, int64_t n, V v) {
  GNO_GRID_STRIDE(i, n) p[i] = v;
}

__device__ __forceinline__ uint32_t enc_f(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_f(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
static uint32_t enc_f_host(float f) {
  uint32_t b;
  memcpy(&b, &f, 4);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

struct ElemShape {
  int64_t B, E, K, N;
  int64_t KB;  // column-block width of the traversal (== K: plain row-major order)
};

// Traversal order.  With dim 0 of a wide matrix (K large) consecutive rows hit destinations
// N*K*4 bytes apart, far beyond L2.  Visiting the elements column block by column block keeps
// the live slice of the accumulator (N x KB x 4 bytes) L2-resident, so the atomics stay on
// chip.  Maps the loop index j to the element's position in src / index.
__device__ __forceinline__ int64_t elem_order(const ElemShape& sh, int64_t j) {
  if (sh.KB >= sh.K) return j;
  const int64_t ek = sh.E * sh.K;
  const int64_t b = j / ek, r = j - b * ek;
  const int64_t full = (sh.K / sh.KB) * sh.E * sh.KB;  // elements in the full-width blocks
  int64_t e, k;
  if (r < full) {
    const int64_t cb = r / (sh.E * sh.KB), r2 = r - cb * sh.E * sh.KB;
    e = r2 / sh.KB;
    k = cb * sh.KB + (r2 - e * sh.KB);
  } else {
    const int64_t w = sh.K - (sh.K / sh.KB) * sh.KB, r2 = r - full;
    e = r2 / w;
    k = (sh.K / sh.KB) * sh.KB + (r2 - e * w);
  }
  return (b * sh.E + e) * sh.K + k;
}

// i over [B,E,K] → flat destination in [B,N,K] (or -1 when index is out of range)
__device__ __forceinline__ int64_t elem_target(const ElemShape& sh, int64_t i, int64_t idx,
                                               int64_t* e_out) {
  const int64_t ek = sh.E * sh.K;
  const int64_t b = i / ek, rem = i - b * ek;
  const int64_t e = rem / sh.K, k = rem - e * sh.K;
  *e_out = e;
  if (idx < 0 || idx >= sh.N) return -1;
  return (b * sh.N + idx) * sh.K + k;
}

template <typename T, bool COUNT>
__global__ void __launch_bounds__(256)
    elem_add_kernel(const T* __restrict__ src, const int64_t* __restrict__ index, ElemShape sh,
                    float* __restrict__ acc, float* __restrict__ cnt) {
  const int64_t n = sh.B * sh.E * sh.K;
  GNO_GRID_STRIDE(j, n) {
    const int64_t i = elem_order(sh, j);
    int64_t e;
    const int64_t t = elem_target(sh, i, index[i], &e);
    if (t < 0) continue;
    atomicAdd(acc + t, DType<T>::to_f(src[i]));
    if (COUNT) atomicAdd(cnt + t, 1.0f);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
    elem_mul_kernel(const T* __restrict__ src, const int64_t* __restrict__ index, ElemShape sh,
                    float* __restrict__ acc) {
  const int64_t n = sh.B * sh.E * sh.K;
  GNO_GRID_STRIDE(j, n) {
    const int64_t i = elem_order(sh, j);
    int64_t e;
    const int64_t t = elem_target(sh, i, index[i], &e);
    if (t < 0) continue;
    const float f = DType<T>::to_f(src[i]);
    unsigned* addr = reinterpret_cast<unsigned*>(acc + t);
    unsigned old = *addr, assumed;
    do {
      assumed = old;
      old = atomicCAS(addr, assumed, __float_as_uint(__uint_as_float(assumed) * f));
    } while (old != assumed);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
    elem_finish_sum_kernel(const float* __restrict__ acc, const float* __restrict__ cnt,
                           T* __restrict__ out, int64_t n) {
  GNO_GRID_STRIDE(i, n) {
    float a = acc[i];
    if (cnt) {
      const float c = cnt[i];
      a = a / (c < 1.f ? 1.f : c);
    }
    out[i] = DType<T>::from_f(a);
  }
}

template <typename T, bool IS_MAX>
__global__ void __launch_bounds__(256)
    elem_minmax_kernel(const T* __restrict__ src, const int64_t* __restrict__ index, ElemShape sh,
                       uint32_t* __restrict__ enc) {
  const int64_t n = sh.B * sh.E * sh.K;
  GNO_GRID_STRIDE(j, n) {
    const int64_t i = elem_order(sh, j);
    int64_t e;
    const int64_t t = elem_target(sh, i, index[i], &e);
    if (t < 0) continue;
    const float f = DType<T>::to_f(src[i]);
    if (f != f) continue;  // NaN never wins (strict compare upstream)
    if (IS_MAX) atomicMax(enc + t, enc_f(f));
    else atomicMin(enc + t, enc_f(f));
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
    elem_arg_kernel(const T* __restrict__ src, const int64_t* __restrict__ index, ElemShape sh,
                    const uint32_t* __restrict__ enc, uint32_t enc_init,
                    long long* __restrict__ arg) {
  const int64_t n = sh.B * sh.E * sh.K;
  GNO_GRID_STRIDE(j, n) {
    const int64_t i = elem_order(sh, j);
    int64_t e;
    const int64_t t = elem_target(sh, i, index[i], &e);
    if (t < 0) continue;
    const uint32_t w = enc[t];
    if (w == enc_init) continue;  // nothing beat the initial value
    if (DType<T>::to_f(src[i]) == dec_f(w)) atomicMin(arg + t, (long long)e);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
    elem_finish_minmax_kernel(const T* __restrict__ src, ElemShape sh,
                              const long long* __restrict__ arg, T* __restrict__ out) {
  const int64_t n = sh.B * sh.N * sh.K;
  const int64_t nk = sh.N * sh.K;
  GNO_GRID_STRIDE(t, n) {
    const long long a = arg[t];
    if (a >= sh.E) {
      out[t] = DType<T>::from_f(0.f);
    } else {
      const int64_t b = t / nk, k = (t - b * nk) % sh.K;
      out[t] = src[(b * sh.E + a) * sh.K + k];
    }
  }
}

template <typename T>
static int scatter_elem_impl(const T* src, const int64_t* index, ElemShape sh, T* out, int64_t* arg,
                             int reduce, void* wsp, size_t ws_bytes, cudaStream_t s) {
  const int64_t n_in = sh.B * sh.E * sh.K;
  const int64_t n_out = sh.B * sh.N * sh.K;
  if (n_out == 0) return GNO_OK;
  Workspace ws(wsp, ws_bytes);
  constexpr bool kIsF32 = sizeof(T) == 4;
  if (reduce == GNO_SUM || reduce == GNO_MEAN || reduce == GNO_MUL) {
    float* acc = kIsF32 ? reinterpret_cast<float*>(out) : ws.take<float>((size_t)n_out);
    float* cnt = (reduce == GNO_MEAN) ? ws.take<float>((size_t)n_out) : nullptr;
    if (!ws.ok() || (ws.off > 0 && wsp == nullptr))
      return fail(GNO_ERR_WORKSPACE, "gno_scatter_elementwise: workspace too small (%zu < %zu)", ws_bytes, ws.off);
    fill_kernel<float><<<grid_for_elems(n_out), 256, 0, s>>>(acc, n_out, reduce == GNO_MUL ? 1.f : 0.f);
    GNO_LAUNCHED("fill_kernel");
    if (cnt) {
      fill_kernel<float><<<grid_for_elems(n_out), 256, 0, s>>>(cnt, n_out, 0.f);
      GNO_LAUNCHED("fill_kernel");
    }
    if (n_in > 0) {
      if (reduce == GNO_MUL)
        elem_mul_kernel<T><<<grid_for_elems(n_in), 256, 0, s>>>(src, index, sh, acc);
      else if (cnt)
        elem_add_kernel<T, true><<<grid_for_elems(n_in), 256, 0, s>>>(src, index, sh, acc, cnt);
      else
        elem_add_kernel<T, false><<<grid_for_elems(n_in), 256, 0, s>>>(src, index, sh, acc, nullptr);
      GNO_LAUNCHED("elem_add_kernel");
    }
    if (!kIsF32 || cnt) {
      elem_finish_sum_kernel<T><<<grid_for_elems(n_out), 256, 0, s>>>(acc, cnt, out, n_out);
      GNO_LAUNCHED("elem_finish_sum_kernel");
    }
    return GNO_OK;
  }
  // MIN / MAX
  uint32_t* enc = ws.take<uint32_t>((size_t)n_out);
  long long* argp = arg ? reinterpret_cast<long long*>(arg) : ws.take<long long>((size_t)n_out);
  if (wsp == nullptr || !ws.ok())
    return fail(GNO_ERR_WORKSPACE, "gno_scatter_elementwise: workspace too small (%zu < %zu)", ws_bytes, ws.off);
  const bool is_max = (reduce == GNO_MAX);
  const uint32_t enc_init = enc_f_host(is_max ? DType<T>::lowest() : DType<T>::highest());
  fill_kernel<uint32_t><<<grid_for_elems(n_out), 256, 0, s>>>(enc, n_out, enc_init);
  GNO_LAUNCHED("fill_kernel");
  fill_kernel<long long><<<grid_for_elems(n_out), 256, 0, s>>>(argp, n_out, (long long)sh.E);
  GNO_LAUNCHED("fill_kernel");
  if (n_in > 0) {
    if (is_max)
      elem_minmax_kernel<T, true><<<grid_for_elems(n_in), 256, 0, s>>>(src, index, sh, enc);
    else
      elem_minmax_kernel<T, false><<<grid_for_elems(n_in), 256, 0, s>>>(src, index, sh, enc);
    GNO_LAUNCHED("elem_minmax_kernel");
    elem_arg_kernel<T><<<grid_for_elems(n_in), 256, 0, s>>>(src, index, sh, enc, enc_init, argp);
    GNO_LAUNCHED("elem_arg_kernel");
  }
  elem_finish_minmax_kernel<T><<<grid_for_elems(n_out), 256, 0, s>>>(src, sh, argp, out);
  GNO_LAUNCHED("elem_finish_minmax_kernel");
  return GNO_OK;
}

}  // namespace gno

using namespace gno;

extern "C" {

int gno_scatter_elementwise_workspace(int64_t B, int64_t E, int64_t N, int64_t K, int dtype, int reduce,
                                      size_t* bytes) {
  GNO_CHECK_ARG(bytes && B >= 0 && E >= 0 && N >= 0 && K >= 0, "gno_scatter_elementwise_workspace: bad argument");
  const int64_t n_out = B * N * K;
  WorkspaceSizer sz;
  if (scatter_onchip_ok(B, E, K, N, dtype, reduce)) {  // bins live in shared memory: no workspace
    *bytes = 0;
    return GNO_OK;
  }
  if (reduce == GNO_MIN || reduce == GNO_MAX) {
    sz.take<uint32_t>((size_t)n_out);
    sz.take<long long>((size_t)n_out);
  } else {
    if (dtype != GNO_F32) sz.take<float>((size_t)n_out);
    if (reduce == GNO_MEAN) sz.take<float>((size_t)n_out);
  }
  *bytes = sz.total();
  return GNO_OK;
}

int gno_scatter_elementwise(const void* src, const int64_t* index, int64_t B, int64_t E, int64_t K,
                            void* out, int64_t* arg, int64_t N, int dtype, int reduce, int accumulate,
                            void* ws, size_t ws_bytes, gno_stream_t stream) {
  GNO_CHECK_ARG(B >= 0 && E >= 0 && K >= 0 && N >= 0, "gno_scatter_elementwise: negative size");
  GNO_CHECK_ARG(reduce >= GNO_SUM && reduce <= GNO_MAX, "gno_scatter_elementwise: unknown reduce %d", reduce);
  GNO_CHECK_ARG(arg == nullptr || reduce == GNO_MIN || reduce == GNO_MAX,
                "gno_scatter_elementwise: arg output only for MIN/MAX");
  if (B * N * K == 0) return GNO_OK;
  GNO_CHECK_ARG(out && (B * E * K == 0 || (src && index)), "gno_scatter_elementwise: NULL buffer");
  GNO_CHECK_ARG(dtype == GNO_F32 || dtype == GNO_F16 || dtype == GNO_BF16,
                "gno_scatter_elementwise: unknown dtype %d", dtype);
  if (B * E * K > 0) {
    const int rc = scatter_onchip(src, index, B, E, K, out, arg, N, dtype, reduce, accumulate,
                                  (cudaStream_t)stream);
    if (rc >= 0) return rc;
  }
  if (accumulate)
    return fail(GNO_ERR_UNSUPPORTED, "gno_scatter_elementwise: accumulate needs the on-chip path "
                "(N too large for shared-memory bins, or an empty input)");
  ElemShape sh{B, E, K, N, K};
  // column-blocked traversal when the accumulator slice of one row-major sweep exceeds ~24 MB
  if (K >= 64 && N * K * 4 > (int64_t(24) << 20)) {
    int64_t kb = (int64_t(24) << 20) / (N * 4) / 32 * 32;
    sh.KB = kb < 32 ? 32 : (kb > K ? K : kb);
  }
  cudaStream_t s = (cudaStream_t)stream;
  switch (dtype) {
    case GNO_F32: return scatter_elem_impl<float>((const float*)src, index, sh, (float*)out, arg, reduce, ws, ws_bytes, s);
    case GNO_F16: return scatter_elem_impl<__half>((const __half*)src, index, sh, (__half*)out, arg, reduce, ws, ws_bytes, s);
    case GNO_BF16: return scatter_elem_impl<__nv_bfloat16>((const __nv_bfloat16*)src, index, sh, (__nv_bfloat16*)out, arg, reduce, ws, ws_bytes, s);
  }
  return fail(GNO_ERR_INVALID, "gno_scatter_elementwise: unknown dtype %d", dtype);
}

}  // extern "C"
