// torch.sort(input, dim, stable=True) for fp32 on top of the radix sort.
//
// Keys are the usual order-preserving float→uint32 transform with torch's
// comparison semantics folded in: every NaN maps to the largest key (NaNs
// last, in input order) and -0.0 maps to +0.0's key (they compare equal, so
// their input order is kept).  Sorting along a dim of [outer, len, inner] is
// one global sort of a 64-bit key (segment id << 32 | float key); the payload
// is the position along `len`.  Values whose key was canonicalised are
// re-read from the input so the output is bit-identical to torch's.
#include "common.cuh"

namespace gno {

int sort_pairs(const void* keys_in, void* keys_out, const void* vals_in, void* vals_out, int64_t n,
               int key_bytes, int val_bytes, int begin_bit, int end_bit, void* ws, size_t ws_bytes,
               cudaStream_t s);
size_t sort_pairs_workspace(int64_t n, int key_bytes, int val_bytes);

#define GNO_GS(i, n)                                                            \
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n);     \
       i += (int64_t)gridDim.x * blockDim.x)

constexpr uint32_t kKeyNaN = 0xffffffffu;
constexpr uint32_t kKeyZero = 0x80000000u;

__device__ __forceinline__ uint32_t float_key(float f) {
  uint32_t b = __float_as_uint(f);
  if (f != f) return kKeyNaN;
  if (b == 0x80000000u) b = 0u;  // -0.0 ≡ +0.0
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

static unsigned sgrid(int64_t n) {
  int64_t b = ceil_div(n, 256);
  int64_t cap = (int64_t)kNumSMs * 32;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}
static int bits_for_u64(uint64_t v) {
  int b = 0;
  while (v) {
    ++b;
    v >>= 1;
  }
  return b;
}

// 1-D: 32-bit keys, payload = position.
__global__ void fkeys32_kernel(const float* __restrict__ in, uint32_t* __restrict__ keys, int64_t n,
                               uint32_t flip) {
  GNO_GS(i, n) keys[i] = float_key(in[i]) ^ flip;
}
__global__ void fout32_kernel(const float* __restrict__ in, const uint32_t* __restrict__ skeys,
                              const uint32_t* __restrict__ sidx, float* __restrict__ out_values,
                              int64_t* __restrict__ out_index, int64_t n, uint32_t flip) {
  GNO_GS(i, n) {
    const uint32_t k = skeys[i] ^ flip;
    const uint32_t j = sidx[i];
    out_index[i] = (int64_t)j;
    out_values[i] = (k == kKeyNaN || k == kKeyZero) ? in[j] : key_float(k);
  }
}

// Segmented: element (o, j, i) of [outer, len, inner]; segment = o*inner + i.
__global__ void fkeys64_kernel(const float* __restrict__ in, uint64_t* __restrict__ keys,
                               uint32_t* __restrict__ pos, int64_t outer, int64_t len,
                               int64_t inner, uint32_t flip) {
  const int64_t n = outer * len * inner;
  GNO_GS(t, n) {
    const int64_t o = t / (len * inner), rem = t - o * len * inner;
    const int64_t j = rem / inner, i = rem - j * inner;
    keys[t] = ((uint64_t)(o * inner + i) << 32) | (uint64_t)(float_key(in[t]) ^ flip);
    pos[t] = (uint32_t)j;
  }
}
__global__ void fout64_kernel(const float* __restrict__ in, const uint64_t* __restrict__ skeys,
                              const uint32_t* __restrict__ spos, float* __restrict__ out_values,
                              int64_t* __restrict__ out_index, int64_t outer, int64_t len,
                              int64_t inner, uint32_t flip) {
  const int64_t n = outer * len * inner;
  GNO_GS(t, n) {  // t = segment * len + rank
    const int64_t seg = t / len, r = t - seg * len;
    const int64_t o = seg / inner, i = seg - o * inner;
    const uint32_t k = (uint32_t)skeys[t] ^ flip;
    const int64_t j = spos[t];
    const int64_t dst = (o * len + r) * inner + i;
    out_index[dst] = j;
    out_values[dst] = (k == kKeyNaN || k == kKeyZero) ? in[(o * len + j) * inner + i] : key_float(k);
  }
}

struct SortF32Ws {
  void* keys;
  void* skeys;
  uint32_t* pos;
  uint32_t* spos;
  char* sort_ws;
  size_t sort_bytes;
};
template <typename W>
static SortF32Ws sortf32_layout(W& ws, int64_t n, bool seg) {
  SortF32Ws c;
  const size_t n1 = (size_t)(n > 0 ? n : 1);
  const int kb = seg ? 8 : 4;
  c.keys = ws.template take<char>(n1 * kb);
  c.skeys = ws.template take<char>(n1 * kb);
  c.pos = seg ? ws.template take<uint32_t>(n1) : nullptr;
  c.spos = ws.template take<uint32_t>(n1);
  c.sort_bytes = sort_pairs_workspace((int64_t)n1, kb, 4);
  c.sort_ws = ws.template take<char>(c.sort_bytes);
  return c;
}
struct SizerShim2 {
  WorkspaceSizer sz;
  template <typename T>
  T* take(size_t n) {
    sz.take<T>(n);
    return nullptr;
  }
};

}  // namespace gno

using namespace gno;

extern "C" {

int gno_sort_f32_workspace(int64_t outer, int64_t len, int64_t inner, size_t* bytes) {
  GNO_CHECK_ARG(bytes && outer >= 0 && len >= 0 && inner >= 0, "gno_sort_f32_workspace: bad argument");
  const int64_t n = outer * len * inner;
  GNO_CHECK_ARG(n < (int64_t(1) << 31), "gno_sort_f32: %lld elements, must be < 2^31", (long long)n);
  SizerShim2 sh;
  sortf32_layout(sh, n, !(outer == 1 && inner == 1));
  *bytes = sh.sz.total();
  return GNO_OK;
}

int gno_sort_f32(const float* in, float* out_values, int64_t* out_index, int64_t outer, int64_t len,
                 int64_t inner, int descending, void* wsp, size_t ws_bytes, gno_stream_t stream) {
  cudaStream_t s = (cudaStream_t)stream;
  GNO_CHECK_ARG(outer >= 0 && len >= 0 && inner >= 0, "gno_sort_f32: negative size");
  const int64_t n = outer * len * inner;
  GNO_CHECK_ARG(n < (int64_t(1) << 31), "gno_sort_f32: %lld elements, must be < 2^31", (long long)n);
  if (n == 0) return GNO_OK;
  GNO_CHECK_ARG(in && out_values && out_index, "gno_sort_f32: NULL buffer");
  if (wsp == nullptr) return fail(GNO_ERR_WORKSPACE, "gno_sort_f32: workspace is NULL");
  const bool seg = !(outer == 1 && inner == 1);
  Workspace ws(wsp, ws_bytes);
  SortF32Ws c = sortf32_layout(ws, n, seg);
  if (!ws.ok()) return fail(GNO_ERR_WORKSPACE, "gno_sort_f32: workspace too small (%zu < %zu)", ws_bytes, ws.off);
  const uint32_t flip = descending ? 0xffffffffu : 0u;
  int rc;
  if (!seg) {
    fkeys32_kernel<<<sgrid(n), 256, 0, s>>>(in, (uint32_t*)c.keys, n, flip);
    GNO_LAUNCHED("fkeys32_kernel");
    rc = sort_pairs(c.keys, c.skeys, nullptr, c.spos, n, 4, 4, 0, 32, c.sort_ws, c.sort_bytes, s);
    if (rc) return rc;
    fout32_kernel<<<sgrid(n), 256, 0, s>>>(in, (const uint32_t*)c.skeys, c.spos, out_values, out_index, n, flip);
    GNO_LAUNCHED("fout32_kernel");
  } else {
    fkeys64_kernel<<<sgrid(n), 256, 0, s>>>(in, (uint64_t*)c.keys, c.pos, outer, len, inner, flip);
    GNO_LAUNCHED("fkeys64_kernel");
    const int end_bit = 32 + bits_for_u64((uint64_t)(outer * inner - 1));
    rc = sort_pairs(c.keys, c.skeys, c.pos, c.spos, n, 8, 4, 0, end_bit, c.sort_ws, c.sort_bytes, s);
    if (rc) return rc;
    fout64_kernel<<<sgrid(n), 256, 0, s>>>(in, (const uint64_t*)c.skeys, c.spos, out_values, out_index, outer, len, inner, flip);
    GNO_LAUNCHED("fout64_kernel");
  }
  return GNO_OK;
}

}  // extern "C"
