// Error state, launch counter and the device-wide exclusive scan shared by
// the plan builder, the radix sort and coalesce.
#include "common.cuh"

namespace gno {

std::atomic<int64_t> g_launches{0};

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

// ------------------------------------------------------------------ scan --
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanChunk = kScanThreads * kScanItems;  // 2048
constexpr int64_t kScanSmallMax = 1 << 16;

// One block walks [begin, end) in chunks, carrying the running sum.
template <typename T>
__device__ void block_scan_range(const T* in, T* out, int64_t begin, int64_t end,
                                 T carry, T* warp_sums) {
  for (int64_t base = begin; base < end; base += kScanChunk) {
    T item[kScanItems];
    T local = 0;
    const int64_t t0 = base + (int64_t)threadIdx.x * kScanItems;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
      item[i] = (t0 + i < end) ? in[t0 + i] : T(0);
      local += item[i];
    }
    T total;
    T ex = block_exclusive_scan_256(local, warp_sums, &total) + carry;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
      if (t0 + i < end) out[t0 + i] = ex;
      ex += item[i];
    }
    carry += total;
  }
}

template <typename T>
__global__ void __launch_bounds__(kScanThreads) scan_small_kernel(const T* in, T* out, int64_t n) {
  __shared__ T warp_sums[kScanThreads / 32 + 1];
  block_scan_range(in, out, 0, n, T(0), warp_sums);
}

template <typename T>
__global__ void __launch_bounds__(kScanThreads)
    scan_reduce_kernel(const T* __restrict__ in, T* __restrict__ partial, int64_t n, int64_t per_block) {
  __shared__ T warp_sums[kScanThreads / 32];
  const int64_t begin = (int64_t)blockIdx.x * per_block;
  const int64_t end = min(begin + per_block, n);
  T local = 0;
  for (int64_t i = begin + threadIdx.x; i < end; i += kScanThreads) local += in[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if (lane_id() == 0) warp_sums[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    T s = 0;
    for (int w = 0; w < kScanThreads / 32; ++w) s += warp_sums[w];
    partial[blockIdx.x] = s;
  }
}

template <typename T>
__global__ void __launch_bounds__(kScanThreads)
    scan_down_kernel(const T* in, T* out, const T* __restrict__ partial, int64_t n, int64_t per_block) {
  __shared__ T warp_sums[kScanThreads / 32 + 1];
  const int64_t begin = (int64_t)blockIdx.x * per_block;
  const int64_t end = min(begin + per_block, n);
  block_scan_range(in, out, begin, end, partial[blockIdx.x], warp_sums);
}

static int64_t scan_blocks(int64_t n, int64_t* per_block) {
  // Enough blocks to fill the chip a few times, each a multiple of the chunk.
  int64_t chunks = ceil_div(n, kScanChunk);
  int64_t blocks = chunks < (int64_t)kNumSMs * 8 ? chunks : (int64_t)kNumSMs * 8;
  int64_t cpb = ceil_div(chunks, blocks);
  *per_block = cpb * kScanChunk;
  return ceil_div(n, *per_block);
}

size_t scan_workspace_elems(int64_t n) {
  if (n <= kScanSmallMax) return 1;
  int64_t per_block;
  return (size_t)scan_blocks(n, &per_block) + 1;
}

template <typename T>
static int exclusive_scan_impl(const T* in, T* out, int64_t n, T* ws, cudaStream_t s) {
  if (n <= 0) return GNO_OK;
  if (n <= kScanSmallMax) {
    scan_small_kernel<T><<<1, kScanThreads, 0, s>>>(in, out, n);
    GNO_LAUNCHED("scan_small_kernel");
    return GNO_OK;
  }
  int64_t per_block;
  int64_t blocks = scan_blocks(n, &per_block);
  scan_reduce_kernel<T><<<(unsigned)blocks, kScanThreads, 0, s>>>(in, ws, n, per_block);
  GNO_LAUNCHED("scan_reduce_kernel");
  scan_small_kernel<T><<<1, kScanThreads, 0, s>>>(ws, ws, blocks);
  GNO_LAUNCHED("scan_small_kernel");
  scan_down_kernel<T><<<(unsigned)blocks, kScanThreads, 0, s>>>(in, out, ws, n, per_block);
  GNO_LAUNCHED("scan_down_kernel");
  return GNO_OK;
}

int exclusive_scan_i64(const int64_t* in, int64_t* out, int64_t n, int64_t* ws, cudaStream_t s) {
  return exclusive_scan_impl<int64_t>(in, out, n, ws, s);
}
int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, int32_t* ws, cudaStream_t s) {
  return exclusive_scan_impl<int32_t>(in, out, n, ws, s);
}

}  // namespace gno

extern "C" {
int gno_abi_version(void) { return GNO_ABI_VERSION; }
const char* gno_last_error(void) { return gno::last_error_buf(); }
int64_t gno_launch_count(void) { return gno::g_launches.load(std::memory_order_relaxed); }
}
