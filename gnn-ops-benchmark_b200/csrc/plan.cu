// Plan builder: dst-sorted CSR (rowptr + perm) of a 1-D index vector, plus
// the split table for rows longer than split_len (power-law graphs).
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

// flag[r] = 1 if row r is longer than split_len; also the max row length.
__global__ void heavy_flag_kernel(const int64_t* __restrict__ rowptr, int32_t* __restrict__ flag,
                                  int64_t N, int64_t split_len, int64_t* __restrict__ info) {
  long long mx = 0;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < N;
       r += (int64_t)gridDim.x * blockDim.x) {
    const long long deg = rowptr[r + 1] - rowptr[r];
    mx = deg > mx ? deg : mx;
    flag[r] = (split_len > 0 && deg > split_len) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    long long u = __shfl_xor_sync(0xffffffffu, mx, o);
    mx = u > mx ? u : mx;
  }
  if (lane_id() == 0 && mx > 0) atomicMax((long long*)&info[1], mx);
}

// pos = exclusive scan of flag. Compacts heavy rows and their chunk counts.
__global__ void heavy_compact_kernel(const int64_t* __restrict__ rowptr,
                                     const int32_t* __restrict__ flag,
                                     const int32_t* __restrict__ pos, int64_t N, int64_t split_len,
                                     int32_t* __restrict__ hrow, int64_t* __restrict__ hcnt,
                                     int64_t* __restrict__ info) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < N;
       r += (int64_t)gridDim.x * blockDim.x) {
    const int f = flag[r];
    if (f) {
      const int64_t deg = rowptr[r + 1] - rowptr[r];
      hrow[pos[r]] = (int32_t)r;
      hcnt[pos[r]] = (deg + split_len - 1) / split_len;
    }
    if (r == N - 1) info[2] = (int64_t)pos[r] + f;
  }
}

__global__ void heavy_total_kernel(const int64_t* __restrict__ hcptr, int64_t* __restrict__ info) {
  info[3] = hcptr[info[2]];
}

// Shared by plan_build and plan_from_rowptr. info[1..3], hrow, hcptr.
static int heavy_analysis(const int64_t* rowptr, int64_t N, int64_t E_cap, int64_t split_len,
                          int64_t* info, int32_t* hrow, int64_t* hcptr, Workspace& ws,
                          cudaStream_t s) {
  const int64_t n1 = N > 0 ? N : 1;
  int32_t* flag = ws.take<int32_t>((size_t)n1);
  int32_t* pos = ws.take<int32_t>((size_t)n1);
  int32_t* scan_ws = ws.take<int32_t>(scan_workspace_elems(n1));
  const int64_t cap = gno_plan_heavy_capacity(E_cap, split_len);
  int64_t* scan_ws64 = ws.take<int64_t>(scan_workspace_elems(cap + 1));
  if (!ws.ok()) return fail(GNO_ERR_WORKSPACE, "plan: workspace too small (%zu < %zu)", ws.size, ws.off);
  GNO_CUDA(cudaMemsetAsync(hcptr, 0, (size_t)(cap + 1) * sizeof(int64_t), s));
  if (N == 0) return GNO_OK;
  heavy_flag_kernel<<<grid_for(N), 256, 0, s>>>(rowptr, flag, N, split_len, info);
  GNO_LAUNCHED("heavy_flag_kernel");
  int rc = exclusive_scan_i32(flag, pos, N, scan_ws, s);
  if (rc) return rc;
  heavy_compact_kernel<<<grid_for(N), 256, 0, s>>>(rowptr, flag, pos, N, split_len, hrow, hcptr, info);
  GNO_LAUNCHED("heavy_compact_kernel");
  rc = exclusive_scan_i64(hcptr, hcptr, cap + 1, scan_ws64, s);
  if (rc) return rc;
  heavy_total_kernel<<<1, 1, 0, s>>>(hcptr, info);
  GNO_LAUNCHED("heavy_total_kernel");
  return GNO_OK;
}

template <typename W>
static void heavy_ws_layout(W& ws, int64_t N, int64_t E_cap, int64_t split_len_min) {
  const int64_t n1 = N > 0 ? N : 1;
  ws.template take<int32_t>((size_t)n1);
  ws.template take<int32_t>((size_t)n1);
  ws.template take<int32_t>(scan_workspace_elems(n1));
  // capacity is largest for the smallest split_len the caller may use (>= 32)
  ws.template take<int64_t>(scan_workspace_elems(E_cap / split_len_min + 2));
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

int64_t gno_plan_heavy_capacity(int64_t E, int64_t split_len) {
  if (split_len <= 0) return 1;
  return E / split_len + 1;
}

int gno_plan_workspace(int64_t E, int64_t N, size_t* bytes) {
  GNO_CHECK_ARG(bytes != nullptr, "gno_plan_workspace: bytes is NULL");
  GNO_CHECK_ARG(E >= 0 && E < (int64_t(1) << 31) && N >= 0 && N < (int64_t(1) << 31) - 1,
                "gno_plan: E=%lld, N=%lld must be < 2^31", (long long)E, (long long)N);
  WorkspaceSizer sz;
  const int64_t e1 = E > 0 ? E : 1;
  sz.take<uint32_t>((size_t)e1);  // narrowed keys
  sz.take<uint32_t>((size_t)e1);  // sorted keys
  sz.take<char>(sort_pairs_workspace(e1, 4, 4));
  heavy_ws_layout(sz, N, e1, 32);
  *bytes = sz.total();
  return GNO_OK;
}

int gno_plan_build(const int64_t* index, int64_t E, int64_t N, int64_t split_len, int64_t* rowptr,
                   int32_t* perm, int64_t* info, int32_t* hrow, int64_t* hcptr, void* wsp,
                   size_t ws_bytes, gno_stream_t stream) {
  cudaStream_t s = (cudaStream_t)stream;
  GNO_CHECK_ARG(E >= 0 && E < (int64_t(1) << 31) && N >= 0 && N < (int64_t(1) << 31) - 1,
                "gno_plan_build: E=%lld, N=%lld must be < 2^31", (long long)E, (long long)N);
  GNO_CHECK_ARG(split_len == 0 || split_len >= 32, "gno_plan_build: split_len must be 0 or >= 32");
  GNO_CHECK_ARG(rowptr && info && hrow && hcptr && (E == 0 || (index && perm)),
                "gno_plan_build: NULL buffer");
  if (wsp == nullptr) return fail(GNO_ERR_WORKSPACE, "gno_plan_build: workspace is NULL");
  Workspace ws(wsp, ws_bytes);
  const int64_t e1 = E > 0 ? E : 1;
  uint32_t* keys = ws.take<uint32_t>((size_t)e1);
  uint32_t* keys_sorted = ws.take<uint32_t>((size_t)e1);
  const size_t sort_bytes = sort_pairs_workspace(e1, 4, 4);
  char* sort_ws = ws.take<char>(sort_bytes);
  if (!ws.ok()) return fail(GNO_ERR_WORKSPACE, "gno_plan_build: workspace too small (%zu < %zu)", ws_bytes, ws.off);
  GNO_CUDA(cudaMemsetAsync(info, 0, 4 * sizeof(int64_t), s));
  if (E > 0) {
    narrow_keys_kernel<<<grid_for(E), 256, 0, s>>>(index, keys, E, N, info);
    GNO_LAUNCHED("narrow_keys_kernel");
    int rc = sort_pairs(keys, keys_sorted, nullptr, perm, E, 4, 4, 0, bits_for(N), sort_ws,
                        sort_bytes, s);
    if (rc) return rc;
  }
  rowptr_kernel<<<grid_for(E + 1), 256, 0, s>>>(keys_sorted, rowptr, E, N);
  GNO_LAUNCHED("rowptr_kernel");
  return heavy_analysis(rowptr, N, e1, split_len, info, hrow, hcptr, ws, s);
}

int gno_plan_from_rowptr_workspace(int64_t N, int64_t E, size_t* bytes) {
  GNO_CHECK_ARG(bytes != nullptr, "gno_plan_from_rowptr_workspace: bytes is NULL");
  GNO_CHECK_ARG(E >= 0 && N >= 0, "gno_plan_from_rowptr_workspace: negative size");
  WorkspaceSizer sz;
  heavy_ws_layout(sz, N, E > 0 ? E : 1, 32);
  *bytes = sz.total();
  return GNO_OK;
}

int gno_plan_from_rowptr(const int64_t* rowptr, int64_t N, int64_t E, int64_t split_len, int64_t* info,
                         int32_t* hrow, int64_t* hcptr, void* wsp, size_t ws_bytes,
                         gno_stream_t stream) {
  cudaStream_t s = (cudaStream_t)stream;
  GNO_CHECK_ARG(N >= 0 && N < (int64_t(1) << 31) - 1 && E >= 0 && E < (int64_t(1) << 31),
                "gno_plan_from_rowptr: N=%lld, E=%lld must be < 2^31", (long long)N, (long long)E);
  GNO_CHECK_ARG(split_len == 0 || split_len >= 32, "gno_plan_from_rowptr: split_len must be 0 or >= 32");
  GNO_CHECK_ARG(rowptr && info && hrow && hcptr, "gno_plan_from_rowptr: NULL buffer");
  if (wsp == nullptr) return fail(GNO_ERR_WORKSPACE, "gno_plan_from_rowptr: workspace is NULL");
  Workspace ws(wsp, ws_bytes);
  GNO_CUDA(cudaMemsetAsync(info, 0, 4 * sizeof(int64_t), s));
  return heavy_analysis(rowptr, N, E > 0 ? E : 1, split_len, info, hrow, hcptr, ws, s);
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
