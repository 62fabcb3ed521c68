// On-chip LSD radix sort (stable), 8 bits per pass.
//
// Per pass: (1) per-tile digit histogram, (2) device-wide exclusive scan of
// the digit-major count matrix, (3) scatter: every 256-thread block ranks a
// 4096-key sub-tile entirely on chip — warp-level match-by-ballot ranking (one
// ballot per digit bit) into per-warp shared-memory histograms, a
// cross-warp/cross-digit scan, a shared-memory exchange into sorted order —
// then writes runs of equal digits to consecutive global addresses.  Order inside a sub-tile is
// (warp, item, lane) == ascending input index, so the sort is stable.
//
// HBM-bound integer work: per pass it reads keys twice and payload once and
// writes both once.  No tensor cores (nothing to contract).
#include "common.cuh"

namespace gno {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortItems = 16;
constexpr int kSortSub = kSortThreads * kSortItems;  // 4096 keys per sub-tile
constexpr int kRadix = 256;

template <typename KeyT>
__global__ void __launch_bounds__(kSortThreads)
    radix_hist_kernel(const KeyT* __restrict__ keys, int32_t* __restrict__ counts, int64_t n,
                      int shift, unsigned mask, int64_t tile_keys, int64_t num_tiles) {
  __shared__ int h[kSortWarps][kRadix];
  const int t = threadIdx.x, w = t >> 5;
#pragma unroll
  for (int i = 0; i < kSortWarps; ++i) h[i][t] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * tile_keys;
  const int64_t end = min(base + tile_keys, n);
  for (int64_t i = base + t; i < end; i += kSortThreads) {
    unsigned d = (unsigned)(keys[i] >> shift) & mask;
    atomicAdd(&h[w][d], 1);
  }
  __syncthreads();
  int s = 0;
#pragma unroll
  for (int i = 0; i < kSortWarps; ++i) s += h[i][t];
  counts[(int64_t)t * num_tiles + blockIdx.x] = s;
}

// Shared memory of the scatter kernel (dynamic): separate key and payload
// exchange buffers so both move with one barrier pair and nothing but the
// ranks stays in registers across it (<= 64 registers ⇒ 4 CTAs / 32 warps per SM).
template <typename KeyT, typename ValT, bool HAS_VAL>
struct ScatterSmem {
  int warp_hist[kSortWarps][kRadix];
  int digit_start[kRadix];
  int gbase[kRadix];
  int scan_tmp[12];
  KeyT ex_k[kSortSub];
  ValT ex_v[HAS_VAL ? kSortSub : 1];
};

// Lanes of the warp that hold the same NBITS-bit digit as this lane: one ballot per digit bit.
// Per bit: VOTE, SEL, LOP3 (ptxas moves the digit bits into predicates with one R2P per key): three
// issue slots; the C++ form (shift, and, compare, ballot, select, not, and) cost eight
// (profiles/r1n_sort_scatter_source_lines.txt).
template <int B>
__device__ __forceinline__ void match_step(unsigned& peers, unsigned d) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b32 v, t, m;\n\t"
      "and.b32 t, %1, %2;\n\t"
      "setp.ne.u32 p, t, 0;\n\t"
      "vote.sync.ballot.b32 v, p, 0xffffffff;\n\t"
      "selp.b32 m, 0, 0xffffffff, p;\n\t"
      "lop3.b32 %0, %0, v, m, 0x60;\n\t}"   // peers & (v ^ m): v for a set bit, ~v for a clear one
      : "+r"(peers)
      : "r"(d), "n"(1u << B));
}
template <int NBITS>
__device__ __forceinline__ unsigned match_digit(unsigned d) {
  unsigned peers = 0xffffffffu;
  match_step<0>(peers, d);
  if constexpr (NBITS > 1) match_step<1>(peers, d);
  if constexpr (NBITS > 2) match_step<2>(peers, d);
  if constexpr (NBITS > 3) match_step<3>(peers, d);
  if constexpr (NBITS > 4) match_step<4>(peers, d);
  if constexpr (NBITS > 5) match_step<5>(peers, d);
  if constexpr (NBITS > 6) match_step<6>(peers, d);
  if constexpr (NBITS > 7) match_step<7>(peers, d);
  return peers;
}

template <typename KeyT, typename ValT, bool HAS_VAL, int NBITS>
__global__ void __launch_bounds__(kSortThreads, sizeof(KeyT) == 4 ? 4 : 3)
    radix_scatter_kernel(const KeyT* __restrict__ keys_in, KeyT* __restrict__ keys_out,
                         const ValT* __restrict__ vals_in, ValT* __restrict__ vals_out,
                         const int32_t* __restrict__ offsets, int64_t n, int shift, unsigned mask,
                         int subtiles, int64_t num_tiles) {
  // NBITS = ballots per key: the digit (key >> shift) & mask has at most NBITS bits
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ScatterSmem<KeyT, ValT, HAS_VAL>& sm = *reinterpret_cast<ScatterSmem<KeyT, ValT, HAS_VAL>*>(smem_raw);
  constexpr int kGroup = 4;  // rounds whose ballots are issued together (ILP across the RMW chain)

  const int t = threadIdx.x, w = t >> 5, l = t & 31;
  const unsigned lt_mask = (1u << l) - 1u;
  const int64_t tile = blockIdx.x;
  int running = offsets[(int64_t)t * num_tiles + tile];  // thread t owns digit t

  for (int sub = 0; sub < subtiles; ++sub) {
    const int64_t base = (tile * subtiles + sub) * (int64_t)kSortSub;
    if (base >= n) break;
    const bool full = base + kSortSub <= n;
#pragma unroll
    for (int i = 0; i < kSortWarps; ++i) sm.warp_hist[i][t] = 0;
    __syncthreads();

    KeyT key[kSortItems];
    int pos[kSortItems];
    const int64_t wbase = base + (int64_t)w * (kSortItems * 32) + l;
    if (full) {
#pragma unroll
      for (int i = 0; i < kSortItems; ++i) key[i] = keys_in[wbase + i * 32];
    } else {
#pragma unroll
      for (int i = 0; i < kSortItems; ++i) {
        const int64_t idx = wbase + i * 32;
        key[i] = (idx < n) ? keys_in[idx] : ~KeyT(0);
      }
    }
    // Warp-level ranking: equal digits in one round get consecutive ranks in
    // lane order; rounds are ordered, so ranks follow input order.
    int* hist = sm.warp_hist[w];
#pragma unroll
    for (int i0 = 0; i0 < kSortItems; i0 += kGroup) {
      unsigned dg[kGroup], peers[kGroup];
#pragma unroll
      for (int g = 0; g < kGroup; ++g) {
        dg[g] = (unsigned)(key[i0 + g] >> shift) & mask;
        peers[g] = match_digit<NBITS>(dg[g]);
      }
#pragma unroll
      for (int g = 0; g < kGroup; ++g) {
        const int leader = __ffs(peers[g]) - 1;
        int old = 0;
        if (l == leader) {
          old = hist[dg[g]];
          hist[dg[g]] = old + __popc(peers[g]);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        pos[i0 + g] = old + __popc(peers[g] & lt_mask);
        __syncwarp();
      }
    }
    __syncthreads();
    // Thread t owns digit t: exclusive scan over warps, then over digits.
    int sum = 0;
#pragma unroll
    for (int i = 0; i < kSortWarps; ++i) {
      const int c = sm.warp_hist[i][t];
      sm.warp_hist[i][t] = sum;
      sum += c;
    }
    int total;
    const int start = block_exclusive_scan_256(sum, sm.scan_tmp, &total);
    sm.digit_start[t] = start;
    sm.gbase[t] = running - start;
    running += sum;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
      const unsigned d = (unsigned)(key[i] >> shift) & mask;
      pos[i] += sm.digit_start[d] + hist[d];
      sm.ex_k[pos[i]] = key[i];
    }
    if (HAS_VAL) {  // payload goes straight from global to its sorted slot
      if (vals_in != nullptr && full) {
        ValT v[kSortItems];
#pragma unroll
        for (int i = 0; i < kSortItems; ++i) v[i] = vals_in[wbase + i * 32];
#pragma unroll
        for (int i = 0; i < kSortItems; ++i) sm.ex_v[pos[i]] = v[i];
      } else {
#pragma unroll
        for (int i = 0; i < kSortItems; ++i) {
          const int64_t idx = wbase + i * 32;
          ValT v;
          if (vals_in != nullptr)
            v = (idx < n) ? vals_in[idx] : ValT(0);
          else
            v = (ValT)idx;  // identity payload: first pass of an argsort
          sm.ex_v[pos[i]] = v;
        }
      }
    }
    __syncthreads();
    if (full) {
#pragma unroll
      for (int j = 0; j < kSortItems; ++j) {
        const int p = j * kSortThreads + t;
        const KeyT k = sm.ex_k[p];
        const unsigned d = (unsigned)(k >> shift) & mask;
        const int gpos = sm.gbase[d] + p;
        keys_out[gpos] = k;
        if (HAS_VAL) vals_out[gpos] = sm.ex_v[p];
      }
    } else {
      const int valid = (int)(n - base);
#pragma unroll
      for (int j = 0; j < kSortItems; ++j) {
        const int p = j * kSortThreads + t;
        if (p < valid) {
          const KeyT k = sm.ex_k[p];
          const unsigned d = (unsigned)(k >> shift) & mask;
          const int gpos = sm.gbase[d] + p;
          keys_out[gpos] = k;
          if (HAS_VAL) vals_out[gpos] = sm.ex_v[p];
        }
      }
    }
    __syncthreads();
  }
}

template <typename T>
__global__ void copy_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = in[i];
}
template <typename T>
__global__ void iota_kernel(T* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (T)i;
}

struct SortGeom {
  int subtiles;
  int64_t tile_keys;
  int64_t num_tiles;
};
static SortGeom sort_geom(int64_t n) {
  SortGeom g;
  int64_t s = n / ((int64_t)kSortSub * kNumSMs * 4);
  g.subtiles = (int)(s < 1 ? 1 : (s > 8 ? 8 : s));
  g.tile_keys = (int64_t)g.subtiles * kSortSub;
  g.num_tiles = ceil_div(n, g.tile_keys);
  return g;
}

template <typename W>
static void sort_layout(W& ws, int64_t n, int key_bytes, int val_bytes, const SortGeom& g) {
  ws.template take<char>((size_t)n * key_bytes);
  if (val_bytes) ws.template take<char>((size_t)n * val_bytes);
  ws.template take<int32_t>((size_t)kRadix * g.num_tiles);
  ws.template take<int32_t>(scan_workspace_elems((int64_t)kRadix * g.num_tiles));
}

template <typename KeyT, typename ValT, bool HAS_VAL>
static int sort_impl(const KeyT* keys_in, KeyT* keys_out, const ValT* vals_in, ValT* vals_out,
                     int64_t n, int begin_bit, int end_bit, void* wsp, size_t ws_bytes,
                     cudaStream_t s) {
  const SortGeom g = sort_geom(n);
  Workspace ws(wsp, ws_bytes);
  KeyT* keys_alt = reinterpret_cast<KeyT*>(ws.take<char>((size_t)n * sizeof(KeyT)));
  ValT* vals_alt = nullptr;
  if (HAS_VAL) vals_alt = reinterpret_cast<ValT*>(ws.take<char>((size_t)n * sizeof(ValT)));
  int32_t* counts = ws.take<int32_t>((size_t)kRadix * g.num_tiles);
  int32_t* scan_ws = ws.take<int32_t>(scan_workspace_elems((int64_t)kRadix * g.num_tiles));
  if (wsp == nullptr || !ws.ok())
    return fail(GNO_ERR_WORKSPACE, "gno_sort_pairs: workspace too small (%zu < %zu)", ws_bytes, ws.off);

  const int passes = (end_bit - begin_bit + 7) / 8;
  const size_t smem_bytes = sizeof(ScatterSmem<KeyT, ValT, HAS_VAL>);
  if (smem_bytes > 48 * 1024) {
    GNO_CUDA(cudaFuncSetAttribute(radix_scatter_kernel<KeyT, ValT, HAS_VAL, 5>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    GNO_CUDA(cudaFuncSetAttribute(radix_scatter_kernel<KeyT, ValT, HAS_VAL, 6>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    GNO_CUDA(cudaFuncSetAttribute(radix_scatter_kernel<KeyT, ValT, HAS_VAL, 7>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    GNO_CUDA(cudaFuncSetAttribute(radix_scatter_kernel<KeyT, ValT, HAS_VAL, 8>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  }
  if (passes == 0) {
    const int grid = (int)gno::imin64(ceil_div(n, 256), (int64_t)kNumSMs * 16);
    copy_kernel<KeyT><<<grid, 256, 0, s>>>(keys_in, keys_out, n);
    GNO_LAUNCHED("copy_kernel");
    if (HAS_VAL) {
      if (vals_in) {
        copy_kernel<ValT><<<grid, 256, 0, s>>>(vals_in, vals_out, n);
      } else {
        iota_kernel<ValT><<<grid, 256, 0, s>>>(vals_out, n);
      }
      GNO_LAUNCHED("copy_kernel");
    }
    return GNO_OK;
  }
  const KeyT* src_k = keys_in;
  const ValT* src_v = vals_in;
  // split the bit range evenly over the passes (18 bits -> 6+6+6 rather than 8+8+2): the ranking
  // costs one ballot per digit bit
  const int total_bits = end_bit - begin_bit;
  int shift = begin_bit;
  for (int p = 0; p < passes; ++p) {
    const bool to_out = ((passes - 1 - p) % 2) == 0;
    KeyT* dst_k = to_out ? keys_out : keys_alt;
    ValT* dst_v = to_out ? vals_out : vals_alt;
    const int nb = total_bits / passes + (p < total_bits % passes ? 1 : 0);
    const unsigned mask = (1u << nb) - 1u;
    radix_hist_kernel<KeyT><<<(unsigned)g.num_tiles, kSortThreads, 0, s>>>(
        src_k, counts, n, shift, mask, g.tile_keys, g.num_tiles);
    GNO_LAUNCHED("radix_hist_kernel");
    int rc = exclusive_scan_i32(counts, counts, (int64_t)kRadix * g.num_tiles, scan_ws, s);
    if (rc) return rc;
#define GNO_SCATTER(NB)                                                                          \
  radix_scatter_kernel<KeyT, ValT, HAS_VAL, NB><<<(unsigned)g.num_tiles, kSortThreads, smem_bytes, s>>>( \
      src_k, dst_k, src_v, dst_v, counts, n, shift, mask, g.subtiles, g.num_tiles)
    if (nb <= 5) GNO_SCATTER(5);
    else if (nb == 6) GNO_SCATTER(6);
    else if (nb == 7) GNO_SCATTER(7);
    else GNO_SCATTER(8);
#undef GNO_SCATTER
    GNO_LAUNCHED("radix_scatter_kernel");
    src_k = dst_k;
    src_v = dst_v;
    shift += nb;
  }
  return GNO_OK;
}

// Internal entry used by plan.cu / coalesce.cu (same TU-external linkage).
int sort_pairs(const void* keys_in, void* keys_out, const void* vals_in, void* vals_out, int64_t n,
               int key_bytes, int val_bytes, int begin_bit, int end_bit, void* ws, size_t ws_bytes,
               cudaStream_t s) {
  if (n == 0) return GNO_OK;
#define GNO_SORT_CASE(KB, VB, KT, VT, HV)                                                      \
  if (key_bytes == KB && val_bytes == VB)                                                      \
    return sort_impl<KT, VT, HV>((const KT*)keys_in, (KT*)keys_out, (const VT*)vals_in,        \
                                 (VT*)vals_out, n, begin_bit, end_bit, ws, ws_bytes, s);
  GNO_SORT_CASE(4, 0, uint32_t, uint32_t, false)
  GNO_SORT_CASE(4, 4, uint32_t, uint32_t, true)
  GNO_SORT_CASE(4, 8, uint32_t, uint64_t, true)
  GNO_SORT_CASE(8, 0, uint64_t, uint32_t, false)
  GNO_SORT_CASE(8, 4, uint64_t, uint32_t, true)
  GNO_SORT_CASE(8, 8, uint64_t, uint64_t, true)
#undef GNO_SORT_CASE
  return fail(GNO_ERR_INVALID, "gno_sort_pairs: key_bytes must be 4|8 and val_bytes 0|4|8");
}

size_t sort_pairs_workspace(int64_t n, int key_bytes, int val_bytes) {
  WorkspaceSizer sz;
  const SortGeom g = sort_geom(n < 1 ? 1 : n);
  sort_layout(sz, n < 1 ? 1 : n, key_bytes, val_bytes, g);
  return sz.total();
}

}  // namespace gno

extern "C" {

int gno_sort_pairs_workspace(int64_t n, int key_bytes, int val_bytes, size_t* bytes) {
  GNO_CHECK_ARG(bytes != nullptr, "gno_sort_pairs_workspace: bytes is NULL");
  GNO_CHECK_ARG(n >= 0 && n < (int64_t(1) << 31), "gno_sort_pairs: n=%lld outside [0, 2^31)", (long long)n);
  GNO_CHECK_ARG((key_bytes == 4 || key_bytes == 8) && (val_bytes == 0 || val_bytes == 4 || val_bytes == 8),
                "gno_sort_pairs: key_bytes must be 4|8 and val_bytes 0|4|8");
  *bytes = gno::sort_pairs_workspace(n, key_bytes, val_bytes);
  return GNO_OK;
}

int gno_sort_pairs(const void* keys_in, void* keys_out, const void* vals_in, void* vals_out,
                   int64_t n, int key_bytes, int val_bytes, int begin_bit, int end_bit, void* ws,
                   size_t ws_bytes, gno_stream_t stream) {
  GNO_CHECK_ARG(n >= 0 && n < (int64_t(1) << 31), "gno_sort_pairs: n=%lld outside [0, 2^31)", (long long)n);
  GNO_CHECK_ARG(begin_bit >= 0 && end_bit >= begin_bit && end_bit <= key_bytes * 8,
                "gno_sort_pairs: bad bit range [%d, %d)", begin_bit, end_bit);
  GNO_CHECK_ARG(n == 0 || (keys_in && keys_out), "gno_sort_pairs: NULL key buffer");
  GNO_CHECK_ARG(n == 0 || val_bytes == 0 || vals_out, "gno_sort_pairs: NULL vals_out");
  return gno::sort_pairs(keys_in, keys_out, vals_in, vals_out, n, key_bytes, val_bytes, begin_bit,
                         end_bit, ws, ws_bytes, (cudaStream_t)stream);
}

}  // extern "C"
