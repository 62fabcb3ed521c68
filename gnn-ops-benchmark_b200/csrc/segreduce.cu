// Segment reduce over a dst-sorted CSR: the hot kernel of the path.
//
//   out[i, :] = reduce_{k in [rowptr[i], rowptr[i+1])}  w[k] * x[gidx[k], :]
//
// Work item = (destination row, 32-vector column tile) → one warp.  A warp
// loads 32 sorted-edge indices with one coalesced request, broadcasts them by
// shuffle, and gathers the feature rows with VB-byte vector loads (16/8/4/2
// chosen from the row alignment), U independent rows in flight per lane.
// Rows narrower than 32 vectors are processed P = 32/G edges at a time by G-lane
// groups and combined by shuffle.  Rows longer than split_len are cut into
// fixed chunks whose partials (fp32 + arg) are combined in chunk order by a
// second kernel — no atomics anywhere, so results are bit-reproducible, and
// arg ties resolve to the lowest original edge position exactly like the
// sequential upstream CPU loop.
//
// Roofline: HBM.  Algorithmic bytes per edge = F*s (feature row) + 4 (index)
// [+4 eid for arg, +s weight]; per row = F*s_out [+8F arg] + 8 (rowptr).
#include <climits>

#include "common.cuh"

namespace gno {

constexpr int kSegThreads = 256;
constexpr int kSegWarps = kSegThreads / 32;
constexpr int kNoArg = INT_MAX;

struct SegParams {
  const int64_t* rowptr;
  const int32_t* gidx;
  const int32_t* eid;
  const void* x;
  const void* w;
  void* out;
  int64_t* arg;
  float* part_val;    // [n_chunks, F] fp32 partials of split rows
  int32_t* part_arg;  // [n_chunks, F]
  const int32_t* hrow;
  const int64_t* hcptr;
  int64_t N, n_heavy, n_chunks, split_len;
  int64_t F;
  int64_t ldx_bytes, ldo_bytes;
  int64_t arg_fill;
  int nvec;      // vectors per row
  int ncoltiles; // column tiles of 32 vectors
  int G;         // lanes per edge group (power of two; 32 when ncoltiles > 1)
  int mean;      // divide by max(row length, 1)
  int accumulate;
};

template <int VB>
struct Words {
  static constexpr int n = VB >= 4 ? VB / 4 : 1;
  uint32_t w[n];
};

template <int VB>
__device__ __forceinline__ Words<VB> ld_vec(const char* p) {
  Words<VB> r;
  if constexpr (VB == 16) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
    r.w[0] = q.x; r.w[1] = q.y; r.w[2] = q.z; r.w[3] = q.w;
  } else if constexpr (VB == 8) {
    const uint2 q = __ldg(reinterpret_cast<const uint2*>(p));
    r.w[0] = q.x; r.w[1] = q.y;
  } else if constexpr (VB == 4) {
    r.w[0] = __ldg(reinterpret_cast<const uint32_t*>(p));
  } else {
    r.w[0] = __ldg(reinterpret_cast<const unsigned short*>(p));
  }
  return r;
}
template <int VB>
__device__ __forceinline__ void st_vec(char* p, const Words<VB>& r) {
  if constexpr (VB == 16) {
    *reinterpret_cast<uint4*>(p) = make_uint4(r.w[0], r.w[1], r.w[2], r.w[3]);
  } else if constexpr (VB == 8) {
    *reinterpret_cast<uint2*>(p) = make_uint2(r.w[0], r.w[1]);
  } else if constexpr (VB == 4) {
    *reinterpret_cast<uint32_t*>(p) = r.w[0];
  } else {
    *reinterpret_cast<unsigned short*>(p) = (unsigned short)r.w[0];
  }
}

template <typename T>
__device__ __forceinline__ float bits_to_f(uint32_t b);
template <>
__device__ __forceinline__ float bits_to_f<float>(uint32_t b) { return __uint_as_float(b); }
template <>
__device__ __forceinline__ float bits_to_f<__half>(uint32_t b) {
  return __half2float(__ushort_as_half((unsigned short)b));
}
template <>
__device__ __forceinline__ float bits_to_f<__nv_bfloat16>(uint32_t b) {
  return __uint_as_float(b << 16);
}
template <typename T>
__device__ __forceinline__ uint32_t f_to_bits(float f);
template <>
__device__ __forceinline__ uint32_t f_to_bits<float>(float f) { return __float_as_uint(f); }
template <>
__device__ __forceinline__ uint32_t f_to_bits<__half>(float f) {
  return __half_as_ushort(__float2half_rn(f));
}
template <>
__device__ __forceinline__ uint32_t f_to_bits<__nv_bfloat16>(float f) {
  return __bfloat16_as_ushort(__float2bfloat16_rn(f));
}

template <typename T, int VB>
__device__ __forceinline__ float elem(const Words<VB>& r, int i) {
  if constexpr (sizeof(T) == 4) {
    return bits_to_f<T>(r.w[i]);
  } else {
    return bits_to_f<T>((r.w[i >> 1] >> ((i & 1) * 16)) & 0xffffu);
  }
}
template <typename T, int VB>
__device__ __forceinline__ void set_elem(Words<VB>& r, int i, float f) {
  if constexpr (sizeof(T) == 4) {
    r.w[i] = f_to_bits<T>(f);
  } else {
    const uint32_t b = f_to_bits<T>(f) & 0xffffu;
    if (i & 1)
      r.w[i >> 1] |= b << 16;
    else
      r.w[i >> 1] = b;
  }
}

template <typename T, int RED>
__device__ __forceinline__ float red_init() {
  if constexpr (RED == GNO_SUM) return 0.f;
  if constexpr (RED == GNO_MUL) return 1.f;
  if constexpr (RED == GNO_MAX) return DType<T>::lowest();
  return DType<T>::highest();
}

// Strict "better than" of the upstream CPU loop: NaNs never win.
template <int RED>
__device__ __forceinline__ bool better(float v, float cur) {
  if constexpr (RED == GNO_MAX) return v > cur;
  return v < cur;
}

// Final element transform shared by the direct and the combine path.
// cnt > 0 selects the mean (divide the sum by the row length, clamped to 1).
template <typename T, int RED>
__device__ __forceinline__ float finalize(float a, int e, float prev, float cnt, int accumulate,
                                          int64_t arg_fill, int64_t* arg_out) {
  if constexpr (RED == GNO_SUM) {
    if (cnt > 0.f) a = a / cnt;
    if (accumulate) a += prev;
  } else if constexpr (RED == GNO_MUL) {
    if (accumulate) a *= prev;
  } else {
    if (a == red_init<T, RED>()) a = 0.f;  // no element ever won (torch_scatter masked_fill_)
    if (arg_out) *arg_out = (e == kNoArg) ? arg_fill : (int64_t)e;
  }
  return a;
}

template <typename T, int VB, int RED, bool ARG, bool HAS_W, int U>
__global__ void __launch_bounds__(kSegThreads) segreduce_kernel(const SegParams p) {
  constexpr int EPV = VB / (int)sizeof(T);
  const int lane = threadIdx.x & 31;
  const int64_t wid = (int64_t)blockIdx.x * kSegWarps + (threadIdx.x >> 5);
  const int64_t n_light = p.N * p.ncoltiles;

  int64_t row, kbeg, kend, chunk = -1;
  int ct;
  if (wid < n_light) {
    row = wid / p.ncoltiles;
    ct = (int)(wid - row * p.ncoltiles);
    kbeg = p.rowptr[row];
    kend = p.rowptr[row + 1];
    if (p.split_len > 0 && kend - kbeg > p.split_len) return;  // split rows: chunk items below
  } else {
    const int64_t hw = wid - n_light;
    chunk = hw / p.ncoltiles;
    ct = (int)(hw - chunk * p.ncoltiles);
    if (chunk >= p.n_chunks) return;
    // largest h with hcptr[h] <= chunk
    int64_t lo = 0, hi = p.n_heavy;
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (p.hcptr[mid] <= chunk) lo = mid; else hi = mid;
    }
    row = p.hrow[lo];
    kbeg = p.rowptr[row] + (chunk - p.hcptr[lo]) * p.split_len;
    kend = min(kbeg + p.split_len, p.rowptr[row + 1]);
  }

  const int G = p.G;
  const int P = 32 / G;
  const int grp = lane / G;
  const int v = ct * 32 + (lane & (G - 1));
  const bool vact = v < p.nvec;
  const char* xcol = static_cast<const char*>(p.x) + (int64_t)v * VB;

  float acc[EPV];
  int ae[ARG ? EPV : 1];
#pragma unroll
  for (int i = 0; i < EPV; ++i) acc[i] = red_init<T, RED>();
  if constexpr (ARG) {
#pragma unroll
    for (int i = 0; i < EPV; ++i) ae[i] = kNoArg;
  }

  const bool stream_idx = (p.ncoltiles == 1);
  const uint64_t pol_stream = l2_policy_evict_first();
  for (int64_t kb = kbeg; kb < kend; kb += 32) {
    const int cnt = (int)gno::imin64(32, kend - kb);
    int my_idx = 0, my_e = 0;
    float my_w = 0.f;
    if (lane < cnt) {
      const int64_t k = kb + lane;
      if (p.gidx)
        my_idx = stream_idx ? ld_stream_i32(p.gidx + k, pol_stream) : __ldg(p.gidx + k);
      else
        my_idx = (int)k;
      if constexpr (ARG) {
        if (p.eid == p.gidx)
          my_e = my_idx;
        else if (p.eid)
          my_e = stream_idx ? ld_stream_i32(p.eid + k, pol_stream) : __ldg(p.eid + k);
        else
          my_e = (int)k;
      }
      if constexpr (HAS_W) my_w = DType<T>::to_f(static_cast<const T*>(p.w)[k]);
    }
    for (int j = 0; j < cnt; j += P * U) {
      Words<VB> val[U];
      int e_u[ARG ? U : 1];
      float w_u[HAS_W ? U : 1];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int slot = j + u * P + grp;
        const int idx = __shfl_sync(0xffffffffu, my_idx, slot & 31);
        if constexpr (ARG) e_u[u] = __shfl_sync(0xffffffffu, my_e, slot & 31);
        if constexpr (HAS_W) w_u[u] = __shfl_sync(0xffffffffu, my_w, slot & 31);
        ok[u] = vact && (slot < cnt);
        if (ok[u]) val[u] = ld_vec<VB>(xcol + (int64_t)idx * p.ldx_bytes);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (ok[u]) {
#pragma unroll
          for (int i = 0; i < EPV; ++i) {
            const float f = elem<T, VB>(val[u], i);
            if constexpr (RED == GNO_SUM) {
              if constexpr (HAS_W) acc[i] = fmaf(w_u[u], f, acc[i]);
              else acc[i] += f;
            } else if constexpr (RED == GNO_MUL) {
              acc[i] *= f;
            } else {
              if (better<RED>(f, acc[i])) {
                acc[i] = f;
                if constexpr (ARG) ae[i] = e_u[u];
              }
            }
          }
        }
      }
    }
  }

  // Combine the P edge groups (fixed butterfly order → deterministic).
  for (int o = G; o < 32; o <<= 1) {
#pragma unroll
    for (int i = 0; i < EPV; ++i) {
      const float ov = __shfl_xor_sync(0xffffffffu, acc[i], o);
      if constexpr (RED == GNO_SUM) {
        acc[i] += ov;
      } else if constexpr (RED == GNO_MUL) {
        acc[i] *= ov;
      } else if constexpr (ARG) {
        const int oe = __shfl_xor_sync(0xffffffffu, ae[i], o);
        if (better<RED>(ov, acc[i]) || (ov == acc[i] && oe < ae[i])) {
          acc[i] = ov;
          ae[i] = oe;
        }
      } else {
        if (better<RED>(ov, acc[i])) acc[i] = ov;
      }
    }
  }
  if (grp != 0 || !vact) return;

  if (chunk >= 0) {  // split row: fp32 partial (+arg) for the combine kernel
    float* pv = p.part_val + chunk * p.F + (int64_t)v * EPV;
#pragma unroll
    for (int i = 0; i < EPV; ++i) pv[i] = acc[i];
    if constexpr (ARG) {
      int32_t* pa = p.part_arg + chunk * p.F + (int64_t)v * EPV;
#pragma unroll
      for (int i = 0; i < EPV; ++i) pa[i] = ae[i];
    }
    return;
  }

  char* optr = static_cast<char*>(p.out) + row * p.ldo_bytes + (int64_t)v * VB;
  Words<VB> prev;
  if (p.accumulate) prev = ld_vec<VB>(optr);
  const float cnt = p.mean ? (float)gno::imax64(kend - kbeg, 1) : 0.f;
  Words<VB> o;
#pragma unroll
  for (int i = 0; i < EPV; ++i) {
    int64_t* ap = nullptr;
    int e = kNoArg;
    if constexpr (ARG) {
      if (p.arg) ap = p.arg + row * p.F + (int64_t)v * EPV + i;
      e = ae[i];
    }
    const float pf = p.accumulate ? elem<T, VB>(prev, i) : 0.f;
    set_elem<T, VB>(o, i, finalize<T, RED>(acc[i], e, pf, cnt, p.accumulate, p.arg_fill, ap));
  }
  st_vec<VB>(optr, o);
}

// Combine the chunk partials of split rows in chunk order: one thread per
// (heavy row, feature).
template <typename T, int RED, bool ARG>
__global__ void __launch_bounds__(256) segcombine_kernel(const SegParams p) {
  const int64_t total = p.n_heavy * p.F;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t h = i / p.F, f = i - h * p.F;
    const int64_t row = p.hrow[h];
    float a = red_init<T, RED>();
    int e = kNoArg;
    for (int64_t c = p.hcptr[h]; c < p.hcptr[h + 1]; ++c) {
      const float pv = p.part_val[c * p.F + f];
      if constexpr (RED == GNO_SUM) {
        a += pv;
      } else if constexpr (RED == GNO_MUL) {
        a *= pv;
      } else {
        // chunks ascend in edge position, so strict compare keeps the lowest
        if (better<RED>(pv, a)) {
          a = pv;
          if constexpr (ARG) e = p.part_arg[c * p.F + f];
        }
      }
    }
    T* op = reinterpret_cast<T*>(static_cast<char*>(p.out) + row * p.ldo_bytes) + f;
    const float prev = p.accumulate ? DType<T>::to_f(*op) : 0.f;
    float cnt = 0.f;
    if (p.mean) cnt = (float)gno::imax64(p.rowptr[row + 1] - p.rowptr[row], 1);
    const float r = finalize<T, RED>(a, e, prev, cnt, p.accumulate, p.arg_fill,
                                     (ARG && p.arg) ? p.arg + row * p.F + f : nullptr);
    *op = DType<T>::from_f(r);
  }
}

// ------------------------------------------------------------- dispatch --
template <typename T, int VB, int RED, bool ARG, bool HAS_W>
static int launch_seg(const SegParams& p, int64_t warps, cudaStream_t s) {
  constexpr int U = VB >= 16 ? 8 : 16;
  const int64_t blocks = ceil_div(warps, kSegWarps);
  if (blocks > 0) {
    segreduce_kernel<T, VB, RED, ARG, HAS_W, U><<<(unsigned)blocks, kSegThreads, 0, s>>>(p);
    GNO_LAUNCHED("segreduce_kernel");
  }
  if (p.n_heavy > 0) {
    const int64_t total = p.n_heavy * p.F;
    const int grid = (int)gno::imin64(ceil_div(total, 256), (int64_t)kNumSMs * 16);
    segcombine_kernel<T, RED, ARG><<<grid, 256, 0, s>>>(p);
    GNO_LAUNCHED("segcombine_kernel");
  }
  return GNO_OK;
}

template <typename T, int VB>
static int dispatch_red(const SegParams& p, int reduce, bool /*with_arg*/, bool has_w,
                        int64_t warps, cudaStream_t s) {
  switch (reduce) {
    case GNO_SUM:
    case GNO_MEAN:
      return has_w ? launch_seg<T, VB, GNO_SUM, false, true>(p, warps, s)
                   : launch_seg<T, VB, GNO_SUM, false, false>(p, warps, s);
    case GNO_MUL:
      return launch_seg<T, VB, GNO_MUL, false, false>(p, warps, s);
    // MIN/MAX always track the winner's position (even with arg == NULL) so
    // that ties between +0.0 and -0.0 resolve to the first occurrence, which
    // keeps the VALUES bit-identical to the sequential upstream loop.
    case GNO_MIN:
      return launch_seg<T, VB, GNO_MIN, true, false>(p, warps, s);
    case GNO_MAX:
      return launch_seg<T, VB, GNO_MAX, true, false>(p, warps, s);
  }
  return fail(GNO_ERR_INVALID, "gno_segment_reduce: unknown reduce %d", reduce);
}

template <typename T>
static int dispatch_vb(const SegParams& p, int vb, int reduce, bool with_arg, bool has_w,
                       int64_t warps, cudaStream_t s) {
  switch (vb) {
    case 16: return dispatch_red<T, 16>(p, reduce, with_arg, has_w, warps, s);
    case 8: return dispatch_red<T, 8>(p, reduce, with_arg, has_w, warps, s);
    case 4: return dispatch_red<T, 4>(p, reduce, with_arg, has_w, warps, s);
    case 2:
      if constexpr (sizeof(T) == 2) return dispatch_red<T, 2>(p, reduce, with_arg, has_w, warps, s);
  }
  return fail(GNO_ERR_INVALID, "gno_segment_reduce: bad vector width %d", vb);
}

static int dtype_size(int dtype) { return dtype == GNO_F32 ? 4 : 2; }

}  // namespace gno

using namespace gno;

extern "C" {

int gno_segment_reduce_workspace(const gno_csr* g, int64_t F, int dtype, int reduce, int with_arg,
                                 size_t* bytes) {
  GNO_CHECK_ARG(g && bytes, "gno_segment_reduce_workspace: NULL argument");
  (void)dtype;
  WorkspaceSizer sz;
  if (g->n_chunks > 0) {
    sz.take<float>((size_t)(g->n_chunks * F));
    if (with_arg || reduce == GNO_MIN || reduce == GNO_MAX)
      sz.take<int32_t>((size_t)(g->n_chunks * F));
  }
  *bytes = sz.total();
  return GNO_OK;
}

int gno_segment_reduce(const gno_csr* g, const void* x, int64_t x_rows, int64_t ldx, const void* w,
                       void* out, int64_t ldo, int64_t* arg, int64_t arg_fill, int64_t F, int dtype,
                       int reduce, int accumulate, void* wsp, size_t ws_bytes,
                       gno_stream_t stream) {
  cudaStream_t s = (cudaStream_t)stream;
  GNO_CHECK_ARG(g != nullptr, "gno_segment_reduce: graph is NULL");
  GNO_CHECK_ARG(dtype == GNO_F32 || dtype == GNO_F16 || dtype == GNO_BF16,
                "gno_segment_reduce: unknown dtype %d", dtype);
  GNO_CHECK_ARG(reduce >= GNO_SUM && reduce <= GNO_MAX, "gno_segment_reduce: unknown reduce %d", reduce);
  GNO_CHECK_ARG(g->N >= 0 && g->E >= 0 && F >= 0 && ldx >= F && ldo >= F,
                "gno_segment_reduce: bad sizes N=%lld E=%lld F=%lld ldx=%lld ldo=%lld",
                (long long)g->N, (long long)g->E, (long long)F, (long long)ldx, (long long)ldo);
  if (g->N == 0 || F == 0) return GNO_OK;
  GNO_CHECK_ARG(g->rowptr && out && (x || g->E == 0), "gno_segment_reduce: NULL buffer");
  GNO_CHECK_ARG(x_rows >= 0, "gno_segment_reduce: x_rows < 0");
  const bool with_arg = (arg != nullptr);
  GNO_CHECK_ARG(!with_arg || reduce == GNO_MIN || reduce == GNO_MAX,
                "gno_segment_reduce: arg output only for MIN/MAX");
  GNO_CHECK_ARG(!accumulate || reduce == GNO_SUM || reduce == GNO_MUL,
                "gno_segment_reduce: accumulate only for SUM/MUL");
  if (w != nullptr && !(reduce == GNO_SUM || reduce == GNO_MEAN))
    return fail(GNO_ERR_UNSUPPORTED, "gno_segment_reduce: edge weights only with SUM/MEAN");
  GNO_CHECK_ARG(g->n_heavy == 0 || (g->hrow && g->hcptr && g->split_len >= 32),
                "gno_segment_reduce: split table missing");

  const int es = dtype_size(dtype);
  // Widest vector every row start and the row length allow.
  const uintptr_t a = (uintptr_t)x | (uintptr_t)out | (uintptr_t)(ldx * es) | (uintptr_t)(ldo * es) |
                      (uintptr_t)(F * es);
  int vb = 16;
  while (vb > es && (a % vb) != 0) vb >>= 1;
  GNO_CHECK_ARG(a % es == 0, "gno_segment_reduce: buffers not aligned to the element size");

  SegParams p;
  p.rowptr = g->rowptr;
  p.gidx = g->gidx;
  p.eid = g->eid;
  p.x = x;
  p.w = w;
  p.out = out;
  p.arg = arg;
  p.hrow = g->hrow;
  p.hcptr = g->hcptr;
  p.N = g->N;
  p.n_heavy = g->split_len > 0 ? g->n_heavy : 0;
  p.n_chunks = g->split_len > 0 ? g->n_chunks : 0;
  p.split_len = g->split_len;
  p.F = F;
  p.ldx_bytes = ldx * es;
  p.ldo_bytes = ldo * es;
  p.arg_fill = arg_fill;
  p.nvec = (int)(F * es / vb);
  p.ncoltiles = (p.nvec + 31) / 32;
  int G = 1;
  while (G < p.nvec && G < 32) G <<= 1;
  p.G = G;
  p.mean = (reduce == GNO_MEAN);
  p.accumulate = accumulate ? 1 : 0;
  p.part_val = nullptr;
  p.part_arg = nullptr;
  if (p.n_chunks > 0) {
    if (wsp == nullptr) return fail(GNO_ERR_WORKSPACE, "gno_segment_reduce: workspace is NULL");
    Workspace ws(wsp, ws_bytes);
    p.part_val = ws.take<float>((size_t)(p.n_chunks * F));
    if (reduce == GNO_MIN || reduce == GNO_MAX)
      p.part_arg = ws.take<int32_t>((size_t)(p.n_chunks * F));
    if (!ws.ok())
      return fail(GNO_ERR_WORKSPACE, "gno_segment_reduce: workspace too small (%zu < %zu)", ws_bytes, ws.off);
  }
  const int64_t warps = (p.N + p.n_chunks) * p.ncoltiles;
  const bool has_w = (w != nullptr);
  switch (dtype) {
    case GNO_F32: return dispatch_vb<float>(p, vb, reduce, with_arg, has_w, warps, s);
    case GNO_F16: return dispatch_vb<__half>(p, vb, reduce, with_arg, has_w, warps, s);
    case GNO_BF16: return dispatch_vb<__nv_bfloat16>(p, vb, reduce, with_arg, has_w, warps, s);
  }
  return fail(GNO_ERR_INVALID, "gno_segment_reduce: unknown dtype %d", dtype);
}

}  // extern "C"
