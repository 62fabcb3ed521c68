// Row gather (index_select along dim 0) and the last-dim segment reduce
// (index_select / index_add_ along dim 1 of a [B, L] matrix).
#include <climits>
#include <cstdlib>

#include "common.cuh"

namespace gno {

template <typename V>
__global__ void __launch_bounds__(256)
    gather_rows_kernel(const V* __restrict__ x, const int64_t* __restrict__ index,
                       V* __restrict__ out, int64_t n_index, int64_t vpr, int64_t x_rows) {
  const int64_t total = n_index * vpr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = i / vpr, j = i - k * vpr;
    int64_t r = index[k];
    r = r < 0 ? 0 : (r >= x_rows ? x_rows - 1 : r);  // clamp: never read outside x
    out[i] = __ldg(x + r * vpr + j);
  }
}

// out[b, i] = reduce_{k in row i} x[b, gidx[k]].  One CTA per source row b:
// the row is staged in shared memory once (coalesced), then every thread
// walks the sorted edges of its destination columns.  RED as gno_reduce.
template <typename T, int RED, bool STAGE>
__global__ void __launch_bounds__(256)
    lastdim_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ gidx,
                   const int32_t* __restrict__ eid, const T* __restrict__ x, int64_t L,
                   int64_t ldx, T* __restrict__ out, int64_t ldo, int64_t* __restrict__ arg,
                   int64_t arg_fill, int64_t N, int mean, int accumulate) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* srow = reinterpret_cast<T*>(smem_raw);
  const int64_t b = blockIdx.x;
  const T* xrow = x + b * ldx;
  if (STAGE) {
    for (int64_t j = threadIdx.x; j < L; j += blockDim.x) srow[j] = xrow[j];
    __syncthreads();
  }
  const T* src = STAGE ? srow : xrow;
  for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
    const int64_t kb = rowptr[i], ke = rowptr[i + 1];
    float a;
    if (RED == GNO_SUM) a = 0.f;
    else if (RED == GNO_MUL) a = 1.f;
    else if (RED == GNO_MAX) a = DType<T>::lowest();
    else a = DType<T>::highest();
    int64_t e = -1;
    for (int64_t k = kb; k < ke; ++k) {
      const int64_t g = gidx ? (int64_t)__ldg(gidx + k) : k;
      const float f = DType<T>::to_f(src[g]);
      if (RED == GNO_SUM) a += f;
      else if (RED == GNO_MUL) a *= f;
      else if (RED == GNO_MAX) { if (f > a) { a = f; e = k; } }
      else { if (f < a) { a = f; e = k; } }
    }
    T* op = out + b * ldo + i;
    if (RED == GNO_SUM) {
      if (mean) a = a / (float)imax64(ke - kb, 1);
      if (accumulate) a += DType<T>::to_f(*op);
    } else if (RED == GNO_MUL) {
      if (accumulate) a *= DType<T>::to_f(*op);
    } else {
      if (e < 0) a = 0.f;
      if (arg) arg[b * N + i] = (e < 0) ? arg_fill : (eid ? (int64_t)__ldg(eid + e) : e);
    }
    *op = DType<T>::from_f(a);
  }
}

// Fused gather + NVLink delivery of the needed-rows exchange: this rank owns x; for every row a
// peer asked for, read it once from local HBM and store it straight into that peer's receive
// buffer through its mapped peer pointer (posted 16-byte stores over NVLink 5 / NVSwitch).  One
// kernel replaces "gather into a send buffer, then all-to-all": no staging copy, and the transfer
// overlaps the gather row by row.  seg[q]..seg[q+1] are the slots requested by rank q; its rows
// land at peer_buf[q] + (row_off[q] + slot - seg[q]) * dst_stride.
constexpr int kMaxPeers = 16;
struct PushParams {
  const char* x;
  const int64_t* serve_rows;
  char* peer_buf[kMaxPeers];
  int64_t seg[kMaxPeers + 1];
  int64_t row_off[kMaxPeers];
  int64_t row_bytes, src_stride, dst_stride, n_serve;
  int64_t start;  // first slot this rank serves: rotating it per rank avoids all ranks pushing to
                  // the same receiver at the same time (NVLink ingress incast)
  int P;
};

// Posted peer stores need no round trip, so a few resident warps per SM keep NVLink busy as long
// as each thread has several row loads in flight (kPushUnroll).  The grid is therefore small by
// choice when the exchange shares the GPU with the reduction (max_blocks): a push kernel that is
// allowed to fill every SM slot at high priority starves the reduction for the whole transfer
// and the two serialise (measured: RMAT-26 at P=2, 25.0 ms sequential vs 26.1 ms "overlapped").
constexpr int kPushUnroll = 4;

template <typename V>
__device__ __forceinline__ void push_locate(const PushParams& p, int64_t i, int64_t vpr, const V*& src,
                                            V*& dst) {
  const int64_t s0 = i / vpr, j = i - s0 * vpr;
  int64_t s = s0 + p.start;
  if (s >= p.n_serve) s -= p.n_serve;
  int q = 0;
#pragma unroll 1
  while (q + 1 < p.P && s >= p.seg[q + 1]) ++q;
  const int64_t r = p.serve_rows ? p.serve_rows[s] : (s - p.seg[q]);  // NULL: every peer gets all rows
  src = reinterpret_cast<const V*>(p.x + r * p.src_stride) + j;
  dst = reinterpret_cast<V*>(p.peer_buf[q] + (p.row_off[q] + (s - p.seg[q])) * p.dst_stride) + j;
}

template <typename V>
__global__ void __launch_bounds__(256) push_rows_kernel(const PushParams p) {
  const int64_t vpr = p.row_bytes / (int64_t)sizeof(V);
  const int64_t total = p.n_serve * vpr;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (kPushUnroll - 1) * stride < total; i += kPushUnroll * stride) {
    const V* src[kPushUnroll];
    V* dst[kPushUnroll];
    V v[kPushUnroll];
#pragma unroll
    for (int u = 0; u < kPushUnroll; ++u) push_locate<V>(p, i + u * stride, vpr, src[u], dst[u]);
#pragma unroll
    for (int u = 0; u < kPushUnroll; ++u) v[u] = __ldg(src[u]);
#pragma unroll
    for (int u = 0; u < kPushUnroll; ++u) *dst[u] = v[u];
  }
  for (; i < total; i += stride) {
    const V* src;
    V* dst;
    push_locate<V>(p, i, vpr, src, dst);
    *dst = __ldg(src);
  }
}


// ---- TMA-staged push ---------------------------------------------------------------------------
// The same exchange driven by the TMA engine instead of by warps of loads and stores.  One warp
// per CTA: its lanes issue one 1-D bulk copy per requested row (global -> shared, completion on
// an mbarrier), and as soon as a chunk of rows has landed ONE bulk copy stores the whole chunk —
// the rows are consecutive in the requester's buffer — from shared memory into the peer's HBM
// over NVLink (cp.async.bulk.global.shared::cta).  A 4-stage ring keeps two chunks of gathers and
// two chunk stores in flight per CTA.  What matters is what the kernel does NOT hold: 32 threads
// and a few registers per SM.  The LDG/STG form needs 300-600 CTAs of 256 threads to saturate
// NVLink (profiles/scaling/r2e_push_micro_grid_cap.txt) and, resident beside the reduction, took
// half of its warp slots and registers: the reduction of the overlapped stage ran 13.7 ms instead
// of 10.3 (profiles/scaling/r2g_*).
constexpr int kTmaStages = 4;
constexpr int kTmaAhead = 2;            // chunks of gathers in flight ahead of the chunk being stored
constexpr int kTmaChunkMax = 16384;  // ring slot size upper bound (GNO_PUSH_CHUNK picks 4096..16384)
static std::atomic<int> g_push_chunk{0};   // gno_push_set_chunk; 0 = GNO_PUSH_CHUNK / default

struct PushTmaParams {
  PushParams b;
  int64_t chunk0[kMaxPeers + 1];  // chunk0[q] = first chunk of requester q (chunks never straddle requesters)
  int64_t start_chunk;            // rotation, so the ranks store to different receivers at any moment
  int rows_per_chunk;
  int chunk_bytes;                // ring slot size
  int hint;                       // L2 evict_first on the rows read and stored (GNO_PUSH_HINT, default on)
};

__global__ void __launch_bounds__(32) push_rows_tma_kernel(const PushTmaParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);            // kTmaStages barriers
  unsigned char* buf = smem_raw + 128;                                // kTmaStages chunk buffers
  const int lane = threadIdx.x;
  const int R = p.rows_per_chunk;
  const uint32_t rb = (uint32_t)p.b.row_bytes;
  const int64_t n_chunks = p.chunk0[p.b.P];
  if (lane == 0) {
    for (int i = 0; i < kTmaStages; ++i) mbar_init(bar + i, 1);
  }
  __syncwarp();
  const int64_t first = blockIdx.x, step = gridDim.x;
  const int64_t n_my = first < n_chunks ? (n_chunks - first + step - 1) / step : 0;
  // rows are read once and stored once: keep them from displacing the hub rows the reduction
  // running beside the push re-reads out of L2
  const uint64_t pol = l2_policy_evict_first();

  auto locate = [&](int64_t it, int& q, int64_t& slot0, int& rows) {
    int64_t c = first + it * step + p.start_chunk;
    if (c >= n_chunks) c -= n_chunks;
    q = 0;
    while (q + 1 < p.b.P && c >= p.chunk0[q + 1]) ++q;
    slot0 = p.b.seg[q] + (c - p.chunk0[q]) * R;
    rows = (int)imin64(R, p.b.seg[q + 1] - slot0);
  };

  for (int64_t it = 0; it < n_my + kTmaAhead; ++it) {
    if (it < n_my) {  // gather chunk `it` into its ring slot
      const int st = (int)(it % kTmaStages);
      // the slot was last read by the store of chunk it - kTmaStages, committed two iterations ago
      if (lane == 0) bulk_wait_read<kTmaStages - kTmaAhead - 1>();
      __syncwarp();
      int q, rows;
      int64_t slot0;
      locate(it, q, slot0, rows);
      if (lane == 0) mbar_expect_tx(bar + st, (uint32_t)rows * rb);
      __syncwarp();
      unsigned char* dst = buf + (size_t)st * p.chunk_bytes;
      for (int r = lane; r < rows; r += 32) {
        const int64_t row = p.b.serve_rows[slot0 + r];
        if (p.hint) bulk_g2s_hint(dst + (size_t)r * rb, p.b.x + row * p.b.src_stride, rb, bar + st, pol);
        else bulk_g2s(dst + (size_t)r * rb, p.b.x + row * p.b.src_stride, rb, bar + st);
      }
    }
    const int64_t j = it - kTmaAhead;  // chunk whose rows have had two iterations to arrive
    if (j >= 0) {
      const int st = (int)(j % kTmaStages);
      mbar_wait(bar + st, (uint32_t)((j / kTmaStages) & 1));
      if (lane == 0) {
        int q, rows;
        int64_t slot0;
        locate(j, q, slot0, rows);
        char* out = p.b.peer_buf[q] + (p.b.row_off[q] + (slot0 - p.b.seg[q])) * p.b.dst_stride;
        if (p.hint) bulk_s2g_hint(out, buf + (size_t)st * p.chunk_bytes, (uint32_t)rows * rb, pol);
        else bulk_s2g(out, buf + (size_t)st * p.chunk_bytes, (uint32_t)rows * rb);
        bulk_commit();
      }
    }
  }
  if (lane == 0) bulk_wait_all();  // every store has been performed before the kernel ends
}

// Batched 2-D transpose out[o, c, r] = in[o, r, c] through a padded 32x32 shared-memory tile
// (coalesced on both sides).  Moves the sorted dim of an inner-dim torch.sort last and back.
template <typename V>
__global__ void __launch_bounds__(256)
    transpose_tiles_kernel(const V* __restrict__ in, V* __restrict__ out, int64_t rows, int64_t cols,
                           int64_t tiles_r, int64_t tiles_c, int64_t n_tiles) {
  __shared__ V tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int64_t o = t / (tiles_r * tiles_c);
    const int64_t rem = t - o * tiles_r * tiles_c;
    const int64_t tr = rem / tiles_c, tc = rem - tr * tiles_c;
    const V* src = in + o * rows * cols;
    V* dst = out + o * rows * cols;
    const int64_t c = tc * 32 + tx;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      const int64_t r = tr * 32 + ty + j;
      if (r < rows && c < cols) tile[ty + j][tx] = src[r * cols + c];
    }
    __syncthreads();
    const int64_t r2 = tr * 32 + tx;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      const int64_t c2 = tc * 32 + ty + j;
      if (r2 < rows && c2 < cols) dst[c2 * rows + r2] = tile[tx][ty + j];
    }
    __syncthreads();
  }
}

template <typename T, int RED>
static int launch_lastdim(const gno_csr* g, const void* x, int64_t B, int64_t L, int64_t ldx,
                          void* out, int64_t ldo, int64_t* arg, int64_t arg_fill, int mean,
                          int accumulate, cudaStream_t s) {
  const size_t bytes = (size_t)L * sizeof(T);
  const bool stage = bytes <= 200 * 1024;
  if (stage) {
    auto k = lastdim_kernel<T, RED, true>;
    if (bytes > 48 * 1024)
      GNO_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    k<<<(unsigned)B, 256, bytes, s>>>(g->rowptr, g->gidx, g->eid, (const T*)x, L, ldx, (T*)out, ldo,
                                       arg, arg_fill, g->N, mean, accumulate);
  } else {
    lastdim_kernel<T, RED, false><<<(unsigned)B, 256, 0, s>>>(
        g->rowptr, g->gidx, g->eid, (const T*)x, L, ldx, (T*)out, ldo, arg, arg_fill, g->N, mean,
        accumulate);
  }
  GNO_LAUNCHED("lastdim_kernel");
  return GNO_OK;
}

template <typename T>
static int dispatch_lastdim(int reduce, const gno_csr* g, const void* x, int64_t B, int64_t L,
                            int64_t ldx, void* out, int64_t ldo, int64_t* arg, int64_t arg_fill,
                            int accumulate, cudaStream_t s) {
  switch (reduce) {
    case GNO_SUM: return launch_lastdim<T, GNO_SUM>(g, x, B, L, ldx, out, ldo, arg, arg_fill, 0, accumulate, s);
    case GNO_MEAN: return launch_lastdim<T, GNO_SUM>(g, x, B, L, ldx, out, ldo, arg, arg_fill, 1, accumulate, s);
    case GNO_MUL: return launch_lastdim<T, GNO_MUL>(g, x, B, L, ldx, out, ldo, arg, arg_fill, 0, accumulate, s);
    case GNO_MIN: return launch_lastdim<T, GNO_MIN>(g, x, B, L, ldx, out, ldo, arg, arg_fill, 0, accumulate, s);
    case GNO_MAX: return launch_lastdim<T, GNO_MAX>(g, x, B, L, ldx, out, ldo, arg, arg_fill, 0, accumulate, s);
  }
  return fail(GNO_ERR_INVALID, "gno_segment_reduce_lastdim: unknown reduce %d", reduce);
}

}  // namespace gno

using namespace gno;

extern "C" {

int gno_gather_rows(const void* x, int64_t x_rows, int64_t row_bytes, const int64_t* index,
                    int64_t n_index, void* out, gno_stream_t stream) {
  if (n_index == 0 || row_bytes == 0) return GNO_OK;
  GNO_CHECK_ARG(x && index && out && x_rows > 0 && n_index > 0 && row_bytes > 0,
                "gno_gather_rows: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  const uintptr_t a = (uintptr_t)x | (uintptr_t)out | (uintptr_t)row_bytes;
  auto grid = [](int64_t n) {
    int64_t b = ceil_div(n, 256);
    return (unsigned)(b > (int64_t)kNumSMs * 32 ? (int64_t)kNumSMs * 32 : b);
  };
  if (a % 16 == 0) {
    const int64_t vpr = row_bytes / 16;
    gather_rows_kernel<uint4><<<grid(n_index * vpr), 256, 0, s>>>((const uint4*)x, index, (uint4*)out, n_index, vpr, x_rows);
  } else if (a % 8 == 0) {
    const int64_t vpr = row_bytes / 8;
    gather_rows_kernel<uint2><<<grid(n_index * vpr), 256, 0, s>>>((const uint2*)x, index, (uint2*)out, n_index, vpr, x_rows);
  } else if (a % 4 == 0) {
    const int64_t vpr = row_bytes / 4;
    gather_rows_kernel<uint32_t><<<grid(n_index * vpr), 256, 0, s>>>((const uint32_t*)x, index, (uint32_t*)out, n_index, vpr, x_rows);
  } else if (a % 2 == 0) {
    const int64_t vpr = row_bytes / 2;
    gather_rows_kernel<uint16_t><<<grid(n_index * vpr), 256, 0, s>>>((const uint16_t*)x, index, (uint16_t*)out, n_index, vpr, x_rows);
  } else {
    gather_rows_kernel<uint8_t><<<grid(n_index * row_bytes), 256, 0, s>>>((const uint8_t*)x, index, (uint8_t*)out, n_index, row_bytes, x_rows);
  }
  GNO_LAUNCHED("gather_rows_kernel");
  return GNO_OK;
}

int gno_pad_rows(const void* src, int64_t rows, int64_t row_bytes, int64_t src_stride_bytes,
                 void* dst, int64_t dst_stride_bytes, gno_stream_t stream) {
  if (rows == 0 || row_bytes == 0) return GNO_OK;
  GNO_CHECK_ARG(src && dst && rows > 0 && row_bytes > 0 && src_stride_bytes >= row_bytes &&
                    dst_stride_bytes >= row_bytes,
                "gno_pad_rows: bad argument");
  GNO_CUDA(cudaMemcpy2DAsync(dst, (size_t)dst_stride_bytes, src, (size_t)src_stride_bytes,
                             (size_t)row_bytes, (size_t)rows, cudaMemcpyDeviceToDevice,
                             (cudaStream_t)stream));
  return GNO_OK;
}

int gno_push_set_chunk(int bytes) {
  GNO_CHECK_ARG(bytes == 0 || (bytes >= 4096 && bytes <= gno::kTmaChunkMax),
                "gno_push_set_chunk: %d outside {0, 4096..16384}", bytes);
  gno::g_push_chunk.store(bytes, std::memory_order_relaxed);
  return GNO_OK;
}

int gno_push_rows(const void* x, int64_t row_bytes, int64_t src_stride_bytes, const int64_t* serve_rows,
                  int64_t n_serve, int n_peers, void* const* peer_bufs, const int64_t* seg,
                  const int64_t* row_off, int64_t dst_stride_bytes, int64_t start_slot,
                  int max_blocks, gno_stream_t stream) {
  if (n_serve == 0 || row_bytes == 0) return GNO_OK;
  GNO_CHECK_ARG(x && peer_bufs && seg && row_off && n_serve > 0 && row_bytes > 0,
                "gno_push_rows: bad argument");
  GNO_CHECK_ARG(n_peers >= 1 && n_peers <= kMaxPeers, "gno_push_rows: 1..%d peers supported", kMaxPeers);
  GNO_CHECK_ARG(src_stride_bytes >= row_bytes && dst_stride_bytes >= row_bytes, "gno_push_rows: bad stride");
  PushParams p;
  p.x = static_cast<const char*>(x);
  p.serve_rows = serve_rows;
  p.row_bytes = row_bytes;
  p.src_stride = src_stride_bytes;
  p.dst_stride = dst_stride_bytes;
  p.n_serve = n_serve;
  p.start = (start_slot >= 0 && start_slot < n_serve) ? start_slot : 0;
  p.P = n_peers;
  uintptr_t a = (uintptr_t)x | (uintptr_t)row_bytes | (uintptr_t)src_stride_bytes | (uintptr_t)dst_stride_bytes;
  for (int q = 0; q < n_peers; ++q) {
    p.peer_buf[q] = static_cast<char*>(peer_bufs[q]);
    p.seg[q] = seg[q];
    p.row_off[q] = row_off[q];
    a |= (uintptr_t)peer_bufs[q];
    GNO_CHECK_ARG(peer_bufs[q] != nullptr || seg[q + 1] == seg[q], "gno_push_rows: NULL peer buffer");
  }
  p.seg[n_peers] = seg[n_peers];
  GNO_CHECK_ARG(seg[0] == 0 && seg[n_peers] == n_serve, "gno_push_rows: seg must cover [0, n_serve)");
  cudaStream_t s = (cudaStream_t)stream;
  // TMA-staged form: rows gathered by index, 16-byte aligned, destination rows contiguous
  static const int tma_env = getenv("GNO_PUSH_TMA") ? atoi(getenv("GNO_PUSH_TMA")) : 1;
  // (chosen when the caller caps the grid, i.e. the exchange runs beside the reduction; alone on
  // the chip the LDG/STG form is faster: 4.4 vs 6.2 ms for RMAT-26's exchange at P=2, where half of
  // the rows are local copies that move at HBM speed)
  if (tma_env && max_blocks > 0 && serve_rows != nullptr && a % 16 == 0 && dst_stride_bytes == row_bytes &&
      row_bytes <= 4096 / 4) {
    // ring slot size: 8 KB x 4 slots = 32 KB per CTA, two CTAs (two issuing warps) per SM by default
    // (gno_push_set_chunk, else GNO_PUSH_CHUNK, else 8 KB.  Measured on RMAT-26, profiles/scaling/r3f_*,
    // r3h_*: where the reduction beside the push is the critical path — 4 GPUs — the light 8 KB ring
    // wins, 12.75 vs 13.09 ms; where the exchange is — 8 GPUs — the 16 KB ring, twice the bytes in
    // flight under HBM contention, wins 7.70 vs 8.22 ms)
    static const int chunk_env = getenv("GNO_PUSH_CHUNK") ? atoi(getenv("GNO_PUSH_CHUNK")) : 8192;
    const int chunk_req = g_push_chunk.load(std::memory_order_relaxed) > 0 ? g_push_chunk.load(std::memory_order_relaxed)
                                                                             : chunk_env;
    const int chunk_bytes = chunk_req < 4096 ? 4096 : (chunk_req > kTmaChunkMax ? kTmaChunkMax : chunk_req / 16 * 16);
    PushTmaParams t;
    t.b = p;
    t.chunk_bytes = chunk_bytes;
    static const int hint_env = getenv("GNO_PUSH_HINT") ? atoi(getenv("GNO_PUSH_HINT")) : 1;
    t.hint = hint_env;
    t.rows_per_chunk = (int)(chunk_bytes / row_bytes);
    t.chunk0[0] = 0;
    for (int q = 0; q < n_peers; ++q)
      t.chunk0[q + 1] = t.chunk0[q] + ceil_div(seg[q + 1] - seg[q], (int64_t)t.rows_per_chunk);
    const int64_t n_chunks = t.chunk0[n_peers];
    // start at the requester the LDG/STG form starts at: the first chunk of the segment holding start_slot
    int q0 = 0;
    while (q0 + 1 < n_peers && p.start >= seg[q0 + 1]) ++q0;
    t.start_chunk = n_chunks > 0 ? t.chunk0[q0] % n_chunks : 0;
    const size_t smem = 128 + (size_t)kTmaStages * chunk_bytes;
    static bool attr_set = false;
    if (!attr_set) {
      GNO_CUDA(cudaFuncSetAttribute(push_rows_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    128 + kTmaStages * kTmaChunkMax));
      attr_set = true;
    }
    // one warp per SM drives 617 GB/s, two 677 GB/s (profiles/scaling/r2h_push_micro_tma.txt);
    // each CTA holds 64 KB of shared memory, so the second one costs the reduction occupancy
    int64_t blocks = (int64_t)kNumSMs * 2;
    if (max_blocks < blocks) blocks = max_blocks;
    if (blocks > n_chunks) blocks = n_chunks;
    if (blocks > 0) {
      push_rows_tma_kernel<<<(unsigned)blocks, 32, smem, s>>>(t);
      GNO_LAUNCHED("push_rows_tma_kernel");
    }
    return GNO_OK;
  }
  const int64_t cap = max_blocks > 0 ? (int64_t)max_blocks : (int64_t)kNumSMs * 32;
  auto grid = [cap](int64_t n) {
    int64_t b = ceil_div(n, 256);
    return (unsigned)(b > cap ? cap : b);
  };
  // Kernels that prefer different L1 / shared-memory splits cannot share an SM; give the push
  // kernel the split of the reduction it runs beside (GNO_PUSH_CARVEOUT = percent of shared memory).
  static const int carve = getenv("GNO_PUSH_CARVEOUT") ? atoi(getenv("GNO_PUSH_CARVEOUT")) : -1;
  static bool carve_set = false;
  if (carve >= 0 && !carve_set) {
    GNO_CUDA(cudaFuncSetAttribute(push_rows_kernel<uint4>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
    carve_set = true;
  }
  if (a % 16 == 0) {
    push_rows_kernel<uint4><<<grid(n_serve * (row_bytes / 16)), 256, 0, s>>>(p);
  } else if (a % 4 == 0) {
    push_rows_kernel<uint32_t><<<grid(n_serve * (row_bytes / 4)), 256, 0, s>>>(p);
  } else if (a % 2 == 0) {
    push_rows_kernel<uint16_t><<<grid(n_serve * (row_bytes / 2)), 256, 0, s>>>(p);
  } else {
    push_rows_kernel<uint8_t><<<grid(n_serve * row_bytes), 256, 0, s>>>(p);
  }
  GNO_LAUNCHED("push_rows_kernel");
  return GNO_OK;
}

int gno_transpose_batched(const void* in, void* out, int64_t outer, int64_t rows, int64_t cols,
                          int elem_bytes, gno_stream_t stream) {
  GNO_CHECK_ARG(outer >= 0 && rows >= 0 && cols >= 0, "gno_transpose_batched: negative size");
  GNO_CHECK_ARG(elem_bytes == 4 || elem_bytes == 8, "gno_transpose_batched: elem_bytes must be 4 or 8");
  if (outer == 0 || rows == 0 || cols == 0) return GNO_OK;
  GNO_CHECK_ARG(in && out && in != out, "gno_transpose_batched: NULL or aliased buffer");
  GNO_CHECK_ARG((((uintptr_t)in | (uintptr_t)out) % elem_bytes) == 0, "gno_transpose_batched: misaligned buffer");
  const int64_t tiles_r = ceil_div(rows, 32), tiles_c = ceil_div(cols, 32);
  const int64_t n_tiles = outer * tiles_r * tiles_c;
  const unsigned grid = (unsigned)imin64(n_tiles, (int64_t)kNumSMs * 64);
  cudaStream_t s = (cudaStream_t)stream;
  if (elem_bytes == 4)
    transpose_tiles_kernel<uint32_t><<<grid, 256, 0, s>>>((const uint32_t*)in, (uint32_t*)out, rows, cols,
                                                           tiles_r, tiles_c, n_tiles);
  else
    transpose_tiles_kernel<uint64_t><<<grid, 256, 0, s>>>((const uint64_t*)in, (uint64_t*)out, rows, cols,
                                                           tiles_r, tiles_c, n_tiles);
  GNO_LAUNCHED("transpose_tiles_kernel");
  return GNO_OK;
}

int gno_segment_reduce_lastdim(const gno_csr* g, const void* x, int64_t B, int64_t L, int64_t ldx,
                               void* out, int64_t ldo, int64_t* arg, int64_t arg_fill, int dtype,
                               int reduce, int accumulate, gno_stream_t stream) {
  GNO_CHECK_ARG(g != nullptr, "gno_segment_reduce_lastdim: graph is NULL");
  GNO_CHECK_ARG(B >= 0 && L >= 0 && ldx >= L && ldo >= g->N, "gno_segment_reduce_lastdim: bad sizes");
  GNO_CHECK_ARG(!accumulate || reduce == GNO_SUM || reduce == GNO_MUL,
                "gno_segment_reduce_lastdim: accumulate only for SUM/MUL");
  GNO_CHECK_ARG(arg == nullptr || reduce == GNO_MIN || reduce == GNO_MAX,
                "gno_segment_reduce_lastdim: arg output only for MIN/MAX");
  if (B == 0 || g->N == 0) return GNO_OK;
  GNO_CHECK_ARG(g->rowptr && out && (x || g->E == 0), "gno_segment_reduce_lastdim: NULL buffer");
  cudaStream_t s = (cudaStream_t)stream;
  switch (dtype) {
    case GNO_F32: return dispatch_lastdim<float>(reduce, g, x, B, L, ldx, out, ldo, arg, arg_fill, accumulate, s);
    case GNO_F16: return dispatch_lastdim<__half>(reduce, g, x, B, L, ldx, out, ldo, arg, arg_fill, accumulate, s);
    case GNO_BF16: return dispatch_lastdim<__nv_bfloat16>(reduce, g, x, B, L, ldx, out, ldo, arg, arg_fill, accumulate, s);
  }
  return fail(GNO_ERR_INVALID, "gno_segment_reduce_lastdim: unknown dtype %d", dtype);
}

}  // extern "C"
