// torch_scatter.scatter with a FULL-SHAPE index, accumulated ON CHIP.
//
// src / index are viewed as [B, E, K], out as [B, N, K]:
//     out[b, index[b,e,k], k] = reduce_e src[b,e,k]
// which is what op_bm_scripts/benchmark_scatter_{add,max,min,mean}.py:60-84 pass (fp16 [L,L],
// dims 0 and 1).  Every (b, k) column is an independent 1-D scatter of E elements into N bins, so
// a CTA owns KB adjacent columns of one b and keeps their N*KB bins in shared memory: the input is
// streamed once (coalesced: KB adjacent columns of a row are contiguous), every update is a
// shared-memory atomic, the bins are written out once.  ONE launch, no global atomics, no
// workspace, no fill or finish kernels.
//
// Determinism.
//   MIN/MAX (+arg): the bin is a packed key (ordered value bits, then the complemented position)
//     maximised with one atomic, so the winner is the extreme value at the LOWEST position e —
//     exactly the sequential upstream loop's strict compare; out = src[arg] restores the winner's
//     own bits (-0.0 stays -0.0).  16-bit data with E <= 65536 packs into 32 bits (native
//     ATOMS.MAX); otherwise 64 bits.  NaNs and values that do not beat the dtype's finite
//     lowest()/max() never win (torch_scatter's init + masked_fill).
//   SUM/MEAN over fp16: every finite fp16 is an integer multiple of 2^-24 below 2^16, so the sum of
//     E <= 2^22 of them is exact in a 64-bit fixed-point bin: integer atomics, order-independent,
//     ONE rounding at the end (more accurate than any float accumulation order).  A non-finite
//     input makes the CTA redo its columns with fp32 bins.
//   SUM/MEAN over fp32/bf16 and MUL: fp32 bins updated by compare-and-swap; the order of the
//     updates (hence the last bits of the rounding) is not fixed — documented, DESIGN.md §4.
//
// Roofline: HBM.  Algorithmic bytes per element = s (src) + 8 (int64 index) + s (out, N = E).
#include <type_traits>

#include "common.cuh"

namespace gno {

enum { OC_FIX64 = 0, OC_F32 = 1, OC_KEY32 = 2, OC_KEY64 = 3 };
enum { OC_SUM = 0, OC_MUL = 1, OC_MIN = 2, OC_MAX = 3 };

struct OnchipParams {
  const void* src;
  const int64_t* index;
  void* out;
  int64_t* arg;
  int64_t B, E, K, N;
  int kb_shift;    // columns per CTA = 1 << kb_shift
  int mean;        // divide by max(count, 1)
  int accumulate;  // combine with the values already in out (torch_scatter's out= form)
};

__device__ __forceinline__ uint32_t oc_enc32(float f) {  // order-preserving, -0.0 == +0.0
  uint32_t b = __float_as_uint(f);
  if (b == 0x80000000u) b = 0u;
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ uint32_t oc_enc16(uint32_t h) {  // raw 16 bits of an fp16 / bf16
  if (h == 0x8000u) h = 0u;
  return ((h & 0x8000u) ? ~h : (h | 0x8000u)) & 0xffffu;
}
template <typename T>
__device__ __forceinline__ uint32_t raw16(T v);
template <>
__device__ __forceinline__ uint32_t raw16<__half>(__half v) { return __half_as_ushort(v); }
template <>
__device__ __forceinline__ uint32_t raw16<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat16_as_ushort(v); }
template <>
__device__ __forceinline__ uint32_t raw16<float>(float) { return 0u; }

__device__ __forceinline__ void smem_mul_f32(float* addr, float f) {
  unsigned* a = reinterpret_cast<unsigned*>(addr);
  unsigned old = *a, assumed;
  do {
    assumed = old;
    old = atomicCAS(a, assumed, __float_as_uint(__uint_as_float(assumed) * f));
  } while (old != assumed);
}

template <typename T>
__device__ __forceinline__ T oc_from_double(double d);
template <>
__device__ __forceinline__ float oc_from_double<float>(double d) { return (float)d; }
template <>
__device__ __forceinline__ __half oc_from_double<__half>(double d) { return __double2half(d); }
template <>
__device__ __forceinline__ __nv_bfloat16 oc_from_double<__nv_bfloat16>(double d) {
  return __float2bfloat16_rn((float)d);
}

constexpr int kOcUnroll = 4;

// One pass over this CTA's columns with bins of the given MODE.  Returns true when a FIX64 pass
// met a non-finite input (the caller then redoes the columns with fp32 bins).
template <typename T, int MODE, int RED>
__device__ __forceinline__ bool oc_accumulate(const OnchipParams& p, const T* s, const int64_t* ix,
                                              int64_t k0, int kw, void* bins_raw, int* cnt) {
  const int KB = 1 << p.kb_shift;
  const int kk = threadIdx.x & (KB - 1);
  const int64_t r0 = threadIdx.x >> p.kb_shift;
  const int64_t rstep = blockDim.x >> p.kb_shift;
  bool special = false;
  if (kk >= kw) return false;
  auto update = [&](int64_t i, T tv, int64_t e) {
    if ((uint64_t)i >= (uint64_t)p.N) return;  // out-of-range destinations are dropped
    const int64_t bin = (i << p.kb_shift) + kk;
    const float v = DType<T>::to_f(tv);
    if constexpr (RED == OC_SUM) {
      if constexpr (MODE == OC_FIX64) {
        if (fabsf(v) <= 65504.f) {  // finite (NaN fails the compare)
          const long long fx = __float2ll_rn(v * 16777216.f);  // exact: v is a multiple of 2^-24
          atomicAdd(reinterpret_cast<unsigned long long*>(bins_raw) + bin, (unsigned long long)fx);
        } else {
          special = true;
        }
      } else {
        atomicAdd(reinterpret_cast<float*>(bins_raw) + bin, v);
      }
      if (cnt) atomicAdd(cnt + bin, 1);
    } else if constexpr (RED == OC_MUL) {
      smem_mul_f32(reinterpret_cast<float*>(bins_raw) + bin, v);
    } else {
      // values that do not beat the finite init never win (NaN fails both compares)
      const bool cand = (RED == OC_MAX) ? (v > DType<T>::lowest()) : (v < DType<T>::highest());
      if (!cand) return;
      if constexpr (MODE == OC_KEY32) {
        uint32_t key = oc_enc16(raw16<T>(tv));
        if (RED == OC_MIN) key = 0xffffu - key;
        atomicMax(reinterpret_cast<uint32_t*>(bins_raw) + bin, (key << 16) | (0xffffu - (uint32_t)e));
      } else {
        uint32_t key = oc_enc32(v);
        if (RED == OC_MIN) key = ~key;
        atomicMax(reinterpret_cast<unsigned long long*>(bins_raw) + bin,
                  ((unsigned long long)key << 32) | (unsigned long long)(0xffffffffu - (uint32_t)e));
      }
    }
  };
  int64_t e = r0;
  for (; e + (kOcUnroll - 1) * rstep < p.E; e += kOcUnroll * rstep) {
    int64_t iv[kOcUnroll];
    T sv[kOcUnroll];
#pragma unroll
    for (int u = 0; u < kOcUnroll; ++u) {
      const int64_t off = (e + u * rstep) * p.K + k0 + kk;
      iv[u] = __ldg(ix + off);
      sv[u] = s[off];
    }
#pragma unroll
    for (int u = 0; u < kOcUnroll; ++u) update(iv[u], sv[u], e + u * rstep);
  }
  for (; e < p.E; e += rstep) {
    const int64_t off = e * p.K + k0 + kk;
    update(__ldg(ix + off), s[off], e);
  }
  return special;
}

template <typename T, int MODE, int RED>
__global__ void __launch_bounds__(1024) scatter_onchip_kernel(const OnchipParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_special;
  const int KB = 1 << p.kb_shift;
  const int64_t ncb = (p.K + KB - 1) >> p.kb_shift;
  const int64_t b = blockIdx.x / ncb, cb = blockIdx.x - b * ncb;
  const int64_t k0 = cb << p.kb_shift;
  const int kw = (int)imin64(KB, p.K - k0);
  const int64_t nbins = p.N << p.kb_shift;
  constexpr int BIN = (MODE == OC_FIX64 || MODE == OC_KEY64) ? 8 : 4;
  int* cnt = p.mean ? reinterpret_cast<int*>(smem_raw + nbins * BIN) : nullptr;
  const T* s = static_cast<const T*>(p.src) + b * p.E * p.K;
  const int64_t* ix = p.index + b * p.E * p.K;

  // ---- init ----
  if (threadIdx.x == 0) s_special = 0;
  if constexpr (BIN == 8) {
    unsigned long long* q = reinterpret_cast<unsigned long long*>(smem_raw);
    for (int64_t j = threadIdx.x; j < nbins; j += blockDim.x) q[j] = 0ull;
  } else {
    uint32_t* q = reinterpret_cast<uint32_t*>(smem_raw);
    const uint32_t init = (RED == OC_MUL) ? __float_as_uint(1.f) : 0u;
    for (int64_t j = threadIdx.x; j < nbins; j += blockDim.x) q[j] = init;
  }
  if (cnt)
    for (int64_t j = threadIdx.x; j < nbins; j += blockDim.x) cnt[j] = 0;
  __syncthreads();

  // ---- accumulate ----
  bool fix_ok = true;
  if (oc_accumulate<T, MODE, RED>(p, s, ix, k0, kw, smem_raw, cnt)) s_special = 1;
  __syncthreads();
  if constexpr (MODE == OC_FIX64) {
    if (s_special) {  // a non-finite input: redo these columns with fp32 bins (IEEE inf/NaN sums)
      fix_ok = false;
      float* q = reinterpret_cast<float*>(smem_raw);
      for (int64_t j = threadIdx.x; j < nbins; j += blockDim.x) q[j] = 0.f;
      if (cnt)
        for (int64_t j = threadIdx.x; j < nbins; j += blockDim.x) cnt[j] = 0;
      __syncthreads();
      oc_accumulate<T, OC_F32, RED>(p, s, ix, k0, kw, smem_raw, cnt);
      __syncthreads();
    }
  }

  // ---- write the bins out ----
  T* o = static_cast<T*>(p.out) + b * p.N * p.K;
  int64_t* a = p.arg ? p.arg + b * p.N * p.K : nullptr;
  for (int64_t j = threadIdx.x; j < nbins; j += blockDim.x) {
    const int kk = (int)(j & (KB - 1));
    if (kk >= kw) continue;
    const int64_t n = j >> p.kb_shift;
    const int64_t oo = n * p.K + k0 + kk;
    if constexpr (RED == OC_SUM) {
      double v;
      if (MODE == OC_FIX64 && fix_ok)
        v = (double)reinterpret_cast<const long long*>(smem_raw)[j] * (1.0 / 16777216.0);
      else
        v = (double)reinterpret_cast<const float*>(smem_raw)[j];
      if (p.accumulate) v += (double)DType<T>::to_f(o[oo]);
      if (cnt) {
        const int c = cnt[j];
        v = v / (double)(c < 1 ? 1 : c);
      }
      o[oo] = oc_from_double<T>(v);
    } else if constexpr (RED == OC_MUL) {
      float v = reinterpret_cast<const float*>(smem_raw)[j];
      if (p.accumulate) v *= DType<T>::to_f(o[oo]);
      o[oo] = DType<T>::from_f(v);
    } else {
      int64_t e;
      bool none;
      if constexpr (MODE == OC_KEY32) {
        const uint32_t key = reinterpret_cast<const uint32_t*>(smem_raw)[j];
        none = key == 0u;
        e = (int64_t)(0xffffu - (key & 0xffffu));
      } else {
        const unsigned long long key = reinterpret_cast<const unsigned long long*>(smem_raw)[j];
        none = key == 0ull;
        e = (int64_t)(0xffffffffu - (uint32_t)(key & 0xffffffffull));
      }
      if (p.accumulate) {
        // out= form: the existing value is the starting point and is kept unless an element
        // beats it strictly (arg then stays at the sentinel, as in the sequential loop)
        if (!none) {
          const T w = s[e * p.K + k0 + kk];
          const float wf = DType<T>::to_f(w), pf = DType<T>::to_f(o[oo]);
          const bool win = (RED == OC_MAX) ? (wf > pf) : (wf < pf);
          if (win) o[oo] = w;
          if (a) a[oo] = win ? e : p.E;
        } else if (a) {
          a[oo] = p.E;
        }
      } else {
        o[oo] = none ? DType<T>::from_f(0.f) : s[e * p.K + k0 + kk];
        if (a) a[oo] = none ? p.E : e;
      }
    }
  }
}

constexpr int64_t kOcSmemBudget = 200 * 1024;

struct OcPlan {
  bool ok;
  int mode, red, kb_shift, threads;
  size_t smem;
  int64_t blocks;
};

static OcPlan oc_plan(int64_t B, int64_t E, int64_t K, int64_t N, int dtype, int reduce) {
  OcPlan pl{};
  pl.ok = false;
  if (B <= 0 || E <= 0 || K <= 0 || N <= 0 || E >= (int64_t(1) << 32)) return pl;
  const bool is16 = dtype != GNO_F32;
  if (reduce == GNO_SUM || reduce == GNO_MEAN) {
    pl.red = OC_SUM;
    pl.mode = (dtype == GNO_F16 && E <= (int64_t(1) << 22)) ? OC_FIX64 : OC_F32;
  } else if (reduce == GNO_MUL) {
    pl.red = OC_MUL;
    pl.mode = OC_F32;
  } else {
    pl.red = reduce == GNO_MIN ? OC_MIN : OC_MAX;
    pl.mode = (is16 && E <= 65536) ? OC_KEY32 : OC_KEY64;
  }
  const int64_t per_bin = ((pl.mode == OC_FIX64 || pl.mode == OC_KEY64) ? 8 : 4) + (reduce == GNO_MEAN ? 4 : 0);
  int sh = 4;  // up to 16 columns per CTA
  while (sh > 0 && ((int64_t(1) << sh) > K * 2 - 1 || (N << sh) * per_bin > kOcSmemBudget)) --sh;
  if ((N << sh) * per_bin > kOcSmemBudget) return pl;  // bins of one column do not fit on chip
  pl.kb_shift = sh;
  const int64_t ncb = (K + (int64_t(1) << sh) - 1) >> sh;
  pl.blocks = B * ncb;
  if (pl.blocks >= (int64_t(1) << 31)) return pl;
  // too few CTAs for a long input: the L2-atomic path spreads the elements over the whole chip
  if (pl.blocks < 64 && (E << sh) > (int64_t(1) << 17)) return pl;
  pl.smem = (size_t)((N << sh) * per_bin);
  pl.threads = (E << sh) >= 4096 ? 1024 : 256;
  pl.ok = true;
  return pl;
}

template <typename T, int MODE, int RED>
static int oc_launch(const OnchipParams& p, const OcPlan& pl, cudaStream_t s) {
  auto k = scatter_onchip_kernel<T, MODE, RED>;
  if (pl.smem > 48 * 1024)
    GNO_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kOcSmemBudget));
  k<<<(unsigned)pl.blocks, pl.threads, pl.smem, s>>>(p);
  GNO_LAUNCHED("scatter_onchip_kernel");
  return GNO_OK;
}

template <typename T>
static int oc_dispatch(const OnchipParams& p, const OcPlan& pl, cudaStream_t s) {
  constexpr bool k16 = sizeof(T) == 2;
  switch (pl.red) {
    case OC_SUM:
      if constexpr (std::is_same<T, __half>::value)
        if (pl.mode == OC_FIX64) return oc_launch<T, OC_FIX64, OC_SUM>(p, pl, s);
      return oc_launch<T, OC_F32, OC_SUM>(p, pl, s);
    case OC_MUL: return oc_launch<T, OC_F32, OC_MUL>(p, pl, s);
    case OC_MIN:
      if constexpr (k16)
        if (pl.mode == OC_KEY32) return oc_launch<T, OC_KEY32, OC_MIN>(p, pl, s);
      return oc_launch<T, OC_KEY64, OC_MIN>(p, pl, s);
    case OC_MAX:
      if constexpr (k16)
        if (pl.mode == OC_KEY32) return oc_launch<T, OC_KEY32, OC_MAX>(p, pl, s);
      return oc_launch<T, OC_KEY64, OC_MAX>(p, pl, s);
  }
  return fail(GNO_ERR_INVALID, "gno_scatter_elementwise: bad on-chip plan");
}

// Returns GNO_OK after launching, or -1 when the shape does not fit the on-chip path.
int scatter_onchip(const void* src, const int64_t* index, int64_t B, int64_t E, int64_t K, void* out,
                   int64_t* arg, int64_t N, int dtype, int reduce, int accumulate, cudaStream_t s) {
  const OcPlan pl = oc_plan(B, E, K, N, dtype, reduce);
  if (!pl.ok) return -1;
  OnchipParams p;
  p.src = src;
  p.index = index;
  p.out = out;
  p.arg = arg;
  p.B = B;
  p.E = E;
  p.K = K;
  p.N = N;
  p.kb_shift = pl.kb_shift;
  p.mean = reduce == GNO_MEAN;
  p.accumulate = accumulate ? 1 : 0;
  switch (dtype) {
    case GNO_F32: return oc_dispatch<float>(p, pl, s);
    case GNO_F16: return oc_dispatch<__half>(p, pl, s);
    case GNO_BF16: return oc_dispatch<__nv_bfloat16>(p, pl, s);
  }
  return fail(GNO_ERR_INVALID, "gno_scatter_elementwise: unknown dtype %d", dtype);
}

bool scatter_onchip_ok(int64_t B, int64_t E, int64_t K, int64_t N, int dtype, int reduce) {
  return oc_plan(B, E, K, N, dtype, reduce).ok;
}

}  // namespace gno
