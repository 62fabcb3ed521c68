// torch_sparse.coalesce / transpose: radix-sort COO entries by (row, col),
// flag run heads, scan, and emit unique indices with their merged values.
//
// Generic path: 64-bit key row*n+col, sorted over exactly the bits m*n needs.
// Fast path (flags bit0, input already ascending in `col`, e.g. the transpose
// of a coalesced matrix): a stable sort on the 32-bit `row` alone is enough —
// the counting-sort-style transpose.
// HBM-bound integer work; merged values accumulate in fp32 in sorted order
// (stable sort ⇒ duplicates are summed in input order: deterministic).
#include "common.cuh"

namespace gno {

int sort_pairs(const void* keys_in, void* keys_out, const void* vals_in, void* vals_out, int64_t n,
               int key_bytes, int val_bytes, int begin_bit, int end_bit, void* ws, size_t ws_bytes,
               cudaStream_t s);
size_t sort_pairs_workspace(int64_t n, int key_bytes, int val_bytes);

static unsigned cgrid(int64_t n) {
  int64_t b = ceil_div(n, 256);
  int64_t cap = (int64_t)kNumSMs * 32;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}
static int bits_for_u64(uint64_t max_value) {  // bits needed to hold [0, max_value]
  int b = 0;
  while (max_value) {
    ++b;
    max_value >>= 1;
  }
  return b;
}

#define GNO_GS(i, n)                                                            \
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n);     \
       i += (int64_t)gridDim.x * blockDim.x)

__global__ void make_keys64_kernel(const int64_t* __restrict__ row, const int64_t* __restrict__ col,
                                   uint64_t* __restrict__ keys, int64_t E, int64_t n) {
  GNO_GS(e, E) keys[e] = (uint64_t)row[e] * (uint64_t)n + (uint64_t)col[e];
}
__global__ void make_keys32_kernel(const int64_t* __restrict__ row, uint32_t* __restrict__ keys,
                                   int64_t E) {
  GNO_GS(e, E) keys[e] = (uint32_t)row[e];
}
__global__ void gather_i64_kernel(const int64_t* __restrict__ src, const uint32_t* __restrict__ perm,
                                  int64_t* __restrict__ out, int64_t E) {
  GNO_GS(k, E) out[k] = src[perm[k]];
}

template <bool FAST>
__global__ void head_kernel(const void* __restrict__ skeys, const int64_t* __restrict__ scol,
                            int32_t* __restrict__ head, int64_t E, int keep_all) {
  GNO_GS(k, E) {
    bool h = (k == 0) || keep_all;  // keep_all: sort only, duplicates stay separate entries
    if (!h) {
      if (FAST) {
        const uint32_t* r = static_cast<const uint32_t*>(skeys);
        h = (r[k] != r[k - 1]) || (scol[k] != scol[k - 1]);
      } else {
        const uint64_t* q = static_cast<const uint64_t*>(skeys);
        h = q[k] != q[k - 1];
      }
    }
    head[k] = h ? 1 : 0;
  }
}

template <bool FAST>
__global__ void emit_index_kernel(const void* __restrict__ skeys, const int64_t* __restrict__ scol,
                                  const int32_t* __restrict__ head, const int32_t* __restrict__ pos,
                                  int64_t E, int64_t n, int64_t* __restrict__ out_row,
                                  int64_t* __restrict__ out_col, int32_t* __restrict__ ustart,
                                  int64_t* __restrict__ nnz_out) {
  GNO_GS(k, E) {
    const int h = head[k];
    const int u = pos[k];
    if (h) {
      if (FAST) {
        out_row[u] = (int64_t) static_cast<const uint32_t*>(skeys)[k];
        out_col[u] = scol[k];
      } else {
        const uint64_t q = static_cast<const uint64_t*>(skeys)[k];
        out_row[u] = (int64_t)(q / (uint64_t)n);
        out_col[u] = (int64_t)(q % (uint64_t)n);
      }
      ustart[u] = (int32_t)k;
    }
    if (k == E - 1) {
      *nnz_out = (int64_t)u + h;
      ustart[u + h] = (int32_t)E;
    }
  }
}

template <typename T, int RED>
__global__ void merge_values_kernel(const T* __restrict__ value, const uint32_t* __restrict__ perm,
                                    const int32_t* __restrict__ ustart,
                                    const int64_t* __restrict__ nnz_dev, int64_t K,
                                    T* __restrict__ out_value, int mean) {
  const int64_t total = (*nnz_dev) * K;
  GNO_GS(i, total) {
    const int64_t u = i / K, kk = i - u * K;
    const int32_t kb = ustart[u], ke = ustart[u + 1];
    float a = DType<T>::to_f(value[(int64_t)perm[kb] * K + kk]);
    for (int32_t k = kb + 1; k < ke; ++k) {
      const float f = DType<T>::to_f(value[(int64_t)perm[k] * K + kk]);
      if (RED == GNO_SUM) a += f;
      else if (RED == GNO_MUL) a *= f;
      else if (RED == GNO_MAX) a = (f > a) ? f : a;
      else a = (f < a) ? f : a;
    }
    if (mean) a = a / (float)(ke - kb);
    out_value[i] = DType<T>::from_f(a);
  }
}

__global__ void order_check_kernel(const int64_t* __restrict__ row, const int64_t* __restrict__ col,
                                   int64_t E, int64_t* __restrict__ status) {
  int inv = 0, dup = 0;
  GNO_GS(k, E) {
    if (k == 0) continue;
    const int64_t r0 = row[k - 1], r1 = row[k], c0 = col[k - 1], c1 = col[k];
    if (r1 < r0 || (r1 == r0 && c1 < c0)) ++inv;
    if (r1 == r0 && c1 == c0) ++dup;
  }
  inv = __reduce_add_sync(0xffffffffu, inv);
  dup = __reduce_add_sync(0xffffffffu, dup);
  if (lane_id() == 0) {
    if (inv) atomicAdd((unsigned long long*)&status[0], (unsigned long long)inv);
    if (dup) atomicAdd((unsigned long long*)&status[1], (unsigned long long)dup);
  }
}

struct CoalesceWs {
  void* keys;
  void* skeys;
  uint32_t* perm;
  char* sort_ws;
  size_t sort_bytes;
  int32_t* head;
  int32_t* pos;
  int32_t* scan_ws;
  int32_t* ustart;
  int64_t* scol;
};
template <typename W>
static CoalesceWs coalesce_layout(W& ws, int64_t E) {
  CoalesceWs c;
  const size_t e1 = (size_t)(E > 0 ? E : 1);
  c.keys = ws.template take<uint64_t>(e1);
  c.skeys = ws.template take<uint64_t>(e1);
  c.perm = ws.template take<uint32_t>(e1);
  c.sort_bytes = sort_pairs_workspace((int64_t)e1, 8, 4);
  c.sort_ws = ws.template take<char>(c.sort_bytes);
  c.head = ws.template take<int32_t>(e1);
  c.pos = ws.template take<int32_t>(e1);
  c.scan_ws = ws.template take<int32_t>(scan_workspace_elems((int64_t)e1));
  c.ustart = ws.template take<int32_t>(e1 + 1);
  c.scol = ws.template take<int64_t>(e1);
  return c;
}
struct SizerShim {  // WorkspaceSizer with take() returning a typed null
  WorkspaceSizer sz;
  template <typename T>
  T* take(size_t n) {
    sz.take<T>(n);
    return nullptr;
  }
};

template <typename T>
static int merge_dispatch(int reduce, const T* value, const uint32_t* perm, const int32_t* ustart,
                          const int64_t* nnz_dev, int64_t K, T* out_value, int64_t E,
                          cudaStream_t s) {
  const unsigned g = cgrid(E * K);
  switch (reduce) {
    case GNO_SUM: merge_values_kernel<T, GNO_SUM><<<g, 256, 0, s>>>(value, perm, ustart, nnz_dev, K, out_value, 0); break;
    case GNO_MEAN: merge_values_kernel<T, GNO_SUM><<<g, 256, 0, s>>>(value, perm, ustart, nnz_dev, K, out_value, 1); break;
    case GNO_MUL: merge_values_kernel<T, GNO_MUL><<<g, 256, 0, s>>>(value, perm, ustart, nnz_dev, K, out_value, 0); break;
    case GNO_MIN: merge_values_kernel<T, GNO_MIN><<<g, 256, 0, s>>>(value, perm, ustart, nnz_dev, K, out_value, 0); break;
    case GNO_MAX: merge_values_kernel<T, GNO_MAX><<<g, 256, 0, s>>>(value, perm, ustart, nnz_dev, K, out_value, 0); break;
    default: return fail(GNO_ERR_INVALID, "gno_coalesce: unknown reduce %d", reduce);
  }
  GNO_LAUNCHED("merge_values_kernel");
  return GNO_OK;
}

}  // namespace gno

using namespace gno;

extern "C" {

int gno_coalesce_workspace(int64_t E, int64_t m, int64_t n, int64_t K, int dtype, size_t* bytes) {
  GNO_CHECK_ARG(bytes != nullptr, "gno_coalesce_workspace: bytes is NULL");
  GNO_CHECK_ARG(E >= 0 && E < (int64_t(1) << 31), "gno_coalesce: E=%lld outside [0, 2^31)", (long long)E);
  (void)m; (void)n; (void)K; (void)dtype;
  SizerShim sh;
  coalesce_layout(sh, E);
  *bytes = sh.sz.total();
  return GNO_OK;
}

int gno_coalesce(const int64_t* row, const int64_t* col, const void* value, int64_t K, int dtype,
                 int64_t E, int64_t m, int64_t n, int reduce, int flags, int64_t* out_row,
                 int64_t* out_col, void* out_value, int64_t* nnz_out, void* wsp, size_t ws_bytes,
                 gno_stream_t stream) {
  cudaStream_t s = (cudaStream_t)stream;
  GNO_CHECK_ARG(E >= 0 && E < (int64_t(1) << 31), "gno_coalesce: E=%lld outside [0, 2^31)", (long long)E);
  GNO_CHECK_ARG(m >= 0 && n >= 0 && K >= 0, "gno_coalesce: negative size");
  GNO_CHECK_ARG(nnz_out != nullptr, "gno_coalesce: nnz_out is NULL");
  if (E == 0) {
    GNO_CUDA(cudaMemsetAsync(nnz_out, 0, sizeof(int64_t), s));
    return GNO_OK;
  }
  GNO_CHECK_ARG(row && col && out_row && out_col, "gno_coalesce: NULL index buffer");
  GNO_CHECK_ARG(m > 0 && n > 0, "gno_coalesce: empty shape with E > 0");
  GNO_CHECK_ARG(value == nullptr || (out_value != nullptr && K > 0), "gno_coalesce: value without out_value/K");
  const bool fast = (flags & 1) != 0;
  const int keep_all = (flags & 2) ? 1 : 0;
  if (fast) {
    GNO_CHECK_ARG(m <= (int64_t(1) << 32), "gno_coalesce: fast path needs m <= 2^32");
  } else {
    GNO_CHECK_ARG((unsigned __int128)m * (unsigned __int128)n <= (unsigned __int128)UINT64_MAX,
                  "gno_coalesce: m*n overflows the 64-bit sort key");
  }
  if (wsp == nullptr) return fail(GNO_ERR_WORKSPACE, "gno_coalesce: workspace is NULL");
  Workspace ws(wsp, ws_bytes);
  CoalesceWs c = coalesce_layout(ws, E);
  if (!ws.ok()) return fail(GNO_ERR_WORKSPACE, "gno_coalesce: workspace too small (%zu < %zu)", ws_bytes, ws.off);

  int rc;
  if (fast) {
    make_keys32_kernel<<<cgrid(E), 256, 0, s>>>(row, (uint32_t*)c.keys, E);
    GNO_LAUNCHED("make_keys32_kernel");
    rc = sort_pairs(c.keys, c.skeys, nullptr, c.perm, E, 4, 4, 0, bits_for_u64((uint64_t)(m - 1)),
                    c.sort_ws, c.sort_bytes, s);
    if (rc) return rc;
    gather_i64_kernel<<<cgrid(E), 256, 0, s>>>(col, c.perm, c.scol, E);
    GNO_LAUNCHED("gather_i64_kernel");
    head_kernel<true><<<cgrid(E), 256, 0, s>>>(c.skeys, c.scol, c.head, E, keep_all);
  } else {
    make_keys64_kernel<<<cgrid(E), 256, 0, s>>>(row, col, (uint64_t*)c.keys, E, n);
    GNO_LAUNCHED("make_keys64_kernel");
    const uint64_t maxkey = (uint64_t)m * (uint64_t)n - 1;
    rc = sort_pairs(c.keys, c.skeys, nullptr, c.perm, E, 8, 4, 0, bits_for_u64(maxkey), c.sort_ws,
                    c.sort_bytes, s);
    if (rc) return rc;
    head_kernel<false><<<cgrid(E), 256, 0, s>>>(c.skeys, nullptr, c.head, E, keep_all);
  }
  GNO_LAUNCHED("head_kernel");
  rc = exclusive_scan_i32(c.head, c.pos, E, c.scan_ws, s);
  if (rc) return rc;
  if (fast)
    emit_index_kernel<true><<<cgrid(E), 256, 0, s>>>(c.skeys, c.scol, c.head, c.pos, E, n, out_row, out_col, c.ustart, nnz_out);
  else
    emit_index_kernel<false><<<cgrid(E), 256, 0, s>>>(c.skeys, nullptr, c.head, c.pos, E, n, out_row, out_col, c.ustart, nnz_out);
  GNO_LAUNCHED("emit_index_kernel");
  if (value != nullptr) {
    switch (dtype) {
      case GNO_F32: return merge_dispatch<float>(reduce, (const float*)value, c.perm, c.ustart, nnz_out, K, (float*)out_value, E, s);
      case GNO_F16: return merge_dispatch<__half>(reduce, (const __half*)value, c.perm, c.ustart, nnz_out, K, (__half*)out_value, E, s);
      case GNO_BF16: return merge_dispatch<__nv_bfloat16>(reduce, (const __nv_bfloat16*)value, c.perm, c.ustart, nnz_out, K, (__nv_bfloat16*)out_value, E, s);
      default: return fail(GNO_ERR_INVALID, "gno_coalesce: unknown dtype %d", dtype);
    }
  }
  return GNO_OK;
}

int gno_coo_order_check(const int64_t* row, const int64_t* col, int64_t E, int64_t n,
                        int64_t* status, gno_stream_t stream) {
  cudaStream_t s = (cudaStream_t)stream;
  (void)n;
  GNO_CHECK_ARG(status != nullptr && E >= 0, "gno_coo_order_check: bad argument");
  GNO_CUDA(cudaMemsetAsync(status, 0, 2 * sizeof(int64_t), s));
  if (E <= 1) return GNO_OK;
  GNO_CHECK_ARG(row && col, "gno_coo_order_check: NULL buffer");
  order_check_kernel<<<cgrid(E), 256, 0, s>>>(row, col, E, status);
  GNO_LAUNCHED("order_check_kernel");
  return GNO_OK;
}

}  // extern "C"
