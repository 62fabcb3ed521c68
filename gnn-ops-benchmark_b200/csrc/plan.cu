// Plan builder: dst-sorted CSR (rowptr + perm + per-edge row) of a 1-D index
// vector, plus the lists of rows the finish pass must touch (rows cut by a
// chunk boundary, empty rows).
#include "common.cuh"

namespace gno {

int sort_pairs(const void* keys_in, void* keys_out, const void* vals_in, void* vals_out, int64_t n,
               int key_bytes, int val_bytes, int begin_bit, int end_bit, void* ws, size_t ws_bytes,
               cudaStream_t s);
size_t sort_pairs_workspace(int64_t n, int key_bytes, int val_bytes);

static int grid_for(int64_t n, int threads = 256) {
  int64_t b = ceil_div(n, threads);
  int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

static int bits_for(int64_t max_value) {  // bits needed to hold values in [0, max_value]
  int b = 0;
  while (b < 63 && (int64_t(1) << b) <= max_value) ++b;
  return b;
}

// key[e] = index[e] if 0 <= index[e] < N else N (sorts last, outside every row).
__global__ void narrow_keys_kernel(const int64_t* __restrict__ index, uint32_t* __restrict__ keys,
                                   int64_t E, int64_t N, int64_t* __restrict__ info) {
  int bad = 0;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = index[e];
    const bool ok = (v >= 0) && (v < N);
    keys[e] = ok ? (uint32_t)v : (uint32_t)N;
    bad += ok ? 0 : 1;
  }
  bad = __reduce_add_sync(0xffffffffu, bad);
  if (lane_id() == 0 && bad) atomicAdd((unsigned long long*)&info[0], (unsigned long long)bad);
}

// rowptr[r] = first sorted position whose key >= r, for r in [0, N].
__global__ void rowptr_kernel(const uint32_t* __restrict__ keys, int64_t* __restrict__ rowptr,
                              int64_t E, int64_t N) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k <= E;
       k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t prev = (k == 0) ? -1 : (int64_t)keys[k - 1];
    const int64_t cur = (k == E) ? N : (int64_t)keys[k];
    for (int64_t r = prev + 1; r <= cur; ++r) rowptr[r] = k;
  }
}

// Row classes for chunk_len-edge chunks: a row "spans" when a chunk boundary
// cuts it (its partial sums are combined by the finish pass); empty rows are
// zero-filled by the finish pass.  MODE 0: count into info[1..3];
// MODE 1: write flags for the compaction scans.
template <int MODE>
__global__ void row_class_kernel(const int64_t* __restrict__ rowptr, int64_t N, int64_t chunk_len,
                                 int32_t* __restrict__ span_flag, int32_t* __restrict__ empty_flag,
                                 int64_t* __restrict__ info) {
  long long mx = 0;
  int n_span = 0, n_empty = 0;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < N;
       r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t kb = rowptr[r], ke = rowptr[r + 1];
    const bool empty = (ke == kb);
    const bool span = !empty && (kb / chunk_len != (ke - 1) / chunk_len);
    if (MODE == 0) {
      mx = (ke - kb) > mx ? (ke - kb) : mx;
      n_span += span;
      n_empty += empty;
    } else {
      span_flag[r] = span;
      empty_flag[r] = empty;
    }
  }
  if (MODE == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      long long u = __shfl_xor_sync(0xffffffffu, mx, o);
      mx = u > mx ? u : mx;
    }
    n_span = __reduce_add_sync(0xffffffffu, n_span);
    n_empty = __reduce_add_sync(0xffffffffu, n_empty);
    if (lane_id() == 0) {
      if (mx > 0) atomicMax((long long*)&info[1], mx);
      if (n_span) atomicAdd((unsigned long long*)&info[2], (unsigned long long)n_span);
      if (n_empty) atomicAdd((unsigned long long*)&info[3], (unsigned long long)n_empty);
    }
  }
}

__global__ void row_compact_kernel(const int32_t* __restrict__ flag, const int32_t* __restrict__ pos,
                                   int64_t N, int32_t* __restrict__ list) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < N;
       r += (int64_t)gridDim.x * blockDim.x)
    if (flag[r]) list[pos[r]] = (int32_t)r;
}

// erow[k] = row containing sorted position k (binary search in rowptr).
__global__ void expand_rows_kernel(const int64_t* __restrict__ rowptr, int64_t N, int64_t E,
                                   int32_t* __restrict__ erow) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < E;
       k += (int64_t)gridDim.x * blockDim.x) {
    int64_t lo = 0, hi = N;  // largest r with rowptr[r] <= k
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (rowptr[mid] <= k) lo = mid; else hi = mid;
    }
    erow[k] = (int32_t)lo;
  }
}

__global__ void permute_i64_to_i32_kernel(const int64_t* __restrict__ src,
                                          const int32_t* __restrict__ perm,
                                          int32_t* __restrict__ out, int64_t E) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < E;
       k += (int64_t)gridDim.x * blockDim.x)
    out[k] = (int32_t)src[perm[k]];
}
__global__ void narrow_i64_to_i32_kernel(const int64_t* __restrict__ src, int32_t* __restrict__ out,
                                         int64_t E) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < E;
       k += (int64_t)gridDim.x * blockDim.x)
    out[k] = (int32_t)src[k];
}

template <typename V>
__global__ void permute_rows_kernel(const V* __restrict__ src, const int32_t* __restrict__ perm,
                                    V* __restrict__ out, int64_t E, int64_t vpr) {
  const int64_t total = E * vpr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = i / vpr, j = i - k * vpr;
    out[i] = src[(int64_t)perm[k] * vpr + j];
  }
}

}  // namespace gno

using namespace gno;

extern "C" {

int gno_plan_workspace(int64_t E, int64_t N, size_t* bytes) {
  GNO_CHECK_ARG(bytes != nullptr, "gno_plan_workspace: bytes is NULL");
  GNO_CHECK_ARG(E >= 0 && E < (int64_t(1) << 31) && N >= 0 && N < (int64_t(1) << 31) - 1,
                "gno_plan: E=%lld, N=%lld must be < 2^31", (long long)E, (long long)N);
  (void)N;
  WorkspaceSizer sz;
  const int64_t e1 = E > 0 ? E : 1;
  sz.take<uint32_t>((size_t)e1);  // narrowed keys
  sz.take<char>(sort_pairs_workspace(e1, 4, 4));
  *bytes = sz.total();
  return GNO_OK;
}

int gno_plan_build(const int64_t* index, int64_t E, int64_t N, int64_t chunk_len, int64_t* rowptr,
                   int32_t* perm, int32_t* erow, int64_t* info, void* wsp, size_t ws_bytes,
                   gno_stream_t stream) {
  cudaStream_t s = (cudaStream_t)stream;
  GNO_CHECK_ARG(E >= 0 && E < (int64_t(1) << 31) && N >= 0 && N < (int64_t(1) << 31) - 1,
                "gno_plan_build: E=%lld, N=%lld must be < 2^31", (long long)E, (long long)N);
  GNO_CHECK_ARG(chunk_len >= 32 && chunk_len % 32 == 0, "gno_plan_build: chunk_len must be a multiple of 32");
  GNO_CHECK_ARG(rowptr && info && (E == 0 || (index && perm && erow)), "gno_plan_build: NULL buffer");
  if (wsp == nullptr) return fail(GNO_ERR_WORKSPACE, "gno_plan_build: workspace is NULL");
  Workspace ws(wsp, ws_bytes);
  const int64_t e1 = E > 0 ? E : 1;
  uint32_t* keys = ws.take<uint32_t>((size_t)e1);
  const size_t sort_bytes = sort_pairs_workspace(e1, 4, 4);
  char* sort_ws = ws.take<char>(sort_bytes);
  if (!ws.ok()) return fail(GNO_ERR_WORKSPACE, "gno_plan_build: workspace too small (%zu < %zu)", ws_bytes, ws.off);
  GNO_CUDA(cudaMemsetAsync(info, 0, 4 * sizeof(int64_t), s));
  if (E > 0) {
    narrow_keys_kernel<<<grid_for(E), 256, 0, s>>>(index, keys, E, N, info);
    GNO_LAUNCHED("narrow_keys_kernel");
    int rc = sort_pairs(keys, erow, nullptr, perm, E, 4, 4, 0, bits_for(N), sort_ws, sort_bytes, s);
    if (rc) return rc;
  }
  rowptr_kernel<<<grid_for(E + 1), 256, 0, s>>>(reinterpret_cast<const uint32_t*>(erow), rowptr, E, N);
  GNO_LAUNCHED("rowptr_kernel");
  if (N > 0) {
    row_class_kernel<0><<<grid_for(N), 256, 0, s>>>(rowptr, N, chunk_len, nullptr, nullptr, info);
    GNO_LAUNCHED("row_class_kernel");
  }
  return GNO_OK;
}

int gno_plan_from_rowptr(const int64_t* rowptr, int64_t N, int64_t E, int64_t chunk_len,
                         int32_t* erow, int64_t* info, gno_stream_t stream) {
  cudaStream_t s = (cudaStream_t)stream;
  GNO_CHECK_ARG(N >= 0 && N < (int64_t(1) << 31) - 1 && E >= 0 && E < (int64_t(1) << 31),
                "gno_plan_from_rowptr: N=%lld, E=%lld must be < 2^31", (long long)N, (long long)E);
  GNO_CHECK_ARG(chunk_len >= 32 && chunk_len % 32 == 0, "gno_plan_from_rowptr: chunk_len must be a multiple of 32");
  GNO_CHECK_ARG(rowptr && info && (E == 0 || erow), "gno_plan_from_rowptr: NULL buffer");
  GNO_CUDA(cudaMemsetAsync(info, 0, 4 * sizeof(int64_t), s));
  if (E > 0 && N > 0) {
    expand_rows_kernel<<<grid_for(E), 256, 0, s>>>(rowptr, N, E, erow);
    GNO_LAUNCHED("expand_rows_kernel");
  }
  if (N > 0) {
    row_class_kernel<0><<<grid_for(N), 256, 0, s>>>(rowptr, N, chunk_len, nullptr, nullptr, info);
    GNO_LAUNCHED("row_class_kernel");
  }
  return GNO_OK;
}

int gno_plan_lists_workspace(int64_t N, size_t* bytes) {
  GNO_CHECK_ARG(bytes != nullptr && N >= 0, "gno_plan_lists_workspace: bad argument");
  WorkspaceSizer sz;
  const int64_t n1 = N > 0 ? N : 1;
  for (int i = 0; i < 4; ++i) sz.take<int32_t>((size_t)n1);
  sz.take<int32_t>(scan_workspace_elems(n1));
  *bytes = sz.total();
  return GNO_OK;
}

int gno_plan_lists(const int64_t* rowptr, int64_t N, int64_t chunk_len, int32_t* srow,
                   int32_t* zrow, void* wsp, size_t ws_bytes, gno_stream_t stream) {
  cudaStream_t s = (cudaStream_t)stream;
  GNO_CHECK_ARG(N >= 0 && N < (int64_t(1) << 31) - 1, "gno_plan_lists: N=%lld must be < 2^31", (long long)N);
  GNO_CHECK_ARG(chunk_len >= 32 && chunk_len % 32 == 0, "gno_plan_lists: chunk_len must be a multiple of 32");
  if (N == 0) return GNO_OK;
  GNO_CHECK_ARG(rowptr != nullptr, "gno_plan_lists: rowptr is NULL");
  if (wsp == nullptr) return fail(GNO_ERR_WORKSPACE, "gno_plan_lists: workspace is NULL");
  Workspace ws(wsp, ws_bytes);
  int32_t* span_flag = ws.take<int32_t>((size_t)N);
  int32_t* empty_flag = ws.take<int32_t>((size_t)N);
  int32_t* span_pos = ws.take<int32_t>((size_t)N);
  int32_t* empty_pos = ws.take<int32_t>((size_t)N);
  int32_t* scan_ws = ws.take<int32_t>(scan_workspace_elems(N));
  if (!ws.ok()) return fail(GNO_ERR_WORKSPACE, "gno_plan_lists: workspace too small (%zu < %zu)", ws_bytes, ws.off);
  row_class_kernel<1><<<grid_for(N), 256, 0, s>>>(rowptr, N, chunk_len, span_flag, empty_flag, nullptr);
  GNO_LAUNCHED("row_class_kernel");
  int rc = exclusive_scan_i32(span_flag, span_pos, N, scan_ws, s);
  if (rc) return rc;
  rc = exclusive_scan_i32(empty_flag, empty_pos, N, scan_ws, s);
  if (rc) return rc;
  if (srow) {
    row_compact_kernel<<<grid_for(N), 256, 0, s>>>(span_flag, span_pos, N, srow);
    GNO_LAUNCHED("row_compact_kernel");
  }
  if (zrow) {
    row_compact_kernel<<<grid_for(N), 256, 0, s>>>(empty_flag, empty_pos, N, zrow);
    GNO_LAUNCHED("row_compact_kernel");
  }
  return GNO_OK;
}

int gno_permute_i64_to_i32(const int64_t* src, const int32_t* perm, int32_t* out, int64_t E,
                           gno_stream_t stream) {
  if (E == 0) return GNO_OK;
  GNO_CHECK_ARG(src && perm && out && E > 0, "gno_permute_i64_to_i32: bad argument");
  permute_i64_to_i32_kernel<<<grid_for(E), 256, 0, (cudaStream_t)stream>>>(src, perm, out, E);
  GNO_LAUNCHED("permute_i64_to_i32_kernel");
  return GNO_OK;
}

int gno_narrow_i64_to_i32(const int64_t* src, int32_t* out, int64_t E, gno_stream_t stream) {
  if (E == 0) return GNO_OK;
  GNO_CHECK_ARG(src && out && E > 0, "gno_narrow_i64_to_i32: bad argument");
  narrow_i64_to_i32_kernel<<<grid_for(E), 256, 0, (cudaStream_t)stream>>>(src, out, E);
  GNO_LAUNCHED("narrow_i64_to_i32_kernel");
  return GNO_OK;
}

int gno_permute_rows(const void* src, const int32_t* perm, void* out, int64_t E, int64_t row_bytes,
                     gno_stream_t stream) {
  if (E == 0 || row_bytes == 0) return GNO_OK;
  GNO_CHECK_ARG(src && perm && out && E > 0 && row_bytes > 0, "gno_permute_rows: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  const uintptr_t a = (uintptr_t)src | (uintptr_t)out | (uintptr_t)row_bytes;
  if (a % 16 == 0) {
    const int64_t vpr = row_bytes / 16;
    permute_rows_kernel<uint4><<<grid_for(E * vpr), 256, 0, s>>>((const uint4*)src, perm, (uint4*)out, E, vpr);
  } else if (a % 4 == 0) {
    const int64_t vpr = row_bytes / 4;
    permute_rows_kernel<uint32_t><<<grid_for(E * vpr), 256, 0, s>>>((const uint32_t*)src, perm, (uint32_t*)out, E, vpr);
  } else if (a % 2 == 0) {
    const int64_t vpr = row_bytes / 2;
    permute_rows_kernel<uint16_t><<<grid_for(E * vpr), 256, 0, s>>>((const uint16_t*)src, perm, (uint16_t*)out, E, vpr);
  } else {
    permute_rows_kernel<uint8_t><<<grid_for(E * row_bytes), 256, 0, s>>>((const uint8_t*)src, perm, (uint8_t*)out, E, row_bytes);
  }
  GNO_LAUNCHED("permute_rows_kernel");
  return GNO_OK;
}

}  // extern "C"
