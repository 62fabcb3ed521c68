// Shared helpers for the gno_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/gno_b200.h"

namespace gno {

// ---------------------------------------------------------------- errors --
char* last_error_buf();  // thread-local, 512 bytes
int fail(int code, const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

#define GNO_CHECK_ARG(cond, ...)                       \
  do {                                                 \
    if (!(cond)) return gno::fail(GNO_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define GNO_CUDA(call)                                                        \
  do {                                                                        \
    cudaError_t _e = (call);                                                  \
    if (_e != cudaSuccess)                                                    \
      return gno::fail(GNO_ERR_CUDA, "%s failed: %s (%s:%d)", #call,          \
                       cudaGetErrorString(_e), __FILE__, __LINE__);           \
  } while (0)

// Count the launch and check for a launch error (no sync).
#define GNO_LAUNCHED(name)                                                    \
  do {                                                                        \
    gno::g_launches.fetch_add(1, std::memory_order_relaxed);                  \
    cudaError_t _e = cudaGetLastError();                                      \
    if (_e != cudaSuccess)                                                    \
      return gno::fail(GNO_ERR_CUDA, "launch of %s failed: %s", name,         \
                       cudaGetErrorString(_e));                               \
  } while (0)

constexpr int kNumSMs = 148;  // B200

__host__ __device__ static inline int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }
__host__ __device__ static inline int64_t imax64(int64_t a, int64_t b) { return a > b ? a : b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Bump allocator over a caller-supplied workspace.
struct Workspace {
  char* base;
  size_t size;
  size_t off = 0;
  Workspace(void* p, size_t n) : base(static_cast<char*>(p)), size(n) {}
  template <typename T>
  T* take(size_t count) {
    off = align_up(off, 256);
    T* r = reinterpret_cast<T*>(base + off);
    off += count * sizeof(T);
    return r;
  }
  bool ok() const { return off <= size; }
};
// Same arithmetic, no memory: used by the *_workspace queries.
struct WorkspaceSizer {
  size_t off = 0;
  template <typename T>
  void take(size_t count) {
    off = align_up(off, 256);
    off += count * sizeof(T);
  }
  size_t total() const { return align_up(off, 256); }
};

// --------------------------------------------------------- device helpers --
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }

// L2 eviction policies (createpolicy): evict_first for read-once streams so
// gathered feature rows keep the cache.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// Streaming 32-bit index load: read once, keep out of L1, evict early from L2.
__device__ __forceinline__ int ld_stream_i32(const int* p, uint64_t pol) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;"
               : "=r"(v)
               : "l"(p), "l"(pol));
  return v;
}

// ---- 1-D TMA bulk copies (cp.async.bulk, SASS UBLKCP) and their mbarriers --------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// the same copy with an L2 eviction policy for the global-memory side (createpolicy handle)
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                                              uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}

// shared -> global (the peer's memory over NVLink, in the push exchange): bulk-group completion
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_s2g_hint(void* dst, const void* src, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst),
               "r"(smem_u32(src)), "r"(bytes), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {  // at most N committed groups still reading shared memory
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <typename T>
struct DType;
template <>
struct DType<float> {
  static constexpr int id = GNO_F32;
  __device__ static float to_f(float v) { return v; }
  __device__ static float from_f(float v) { return v; }
  __host__ __device__ static float lowest() { return -3.402823466e+38f; }
  __host__ __device__ static float highest() { return 3.402823466e+38f; }
};
template <>
struct DType<__half> {
  static constexpr int id = GNO_F16;
  __device__ static float to_f(__half v) { return __half2float(v); }
  __device__ static __half from_f(float v) { return __float2half_rn(v); }
  __host__ __device__ static float lowest() { return -65504.f; }
  __host__ __device__ static float highest() { return 65504.f; }
};
template <>
struct DType<__nv_bfloat16> {
  static constexpr int id = GNO_BF16;
  __device__ static float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  __device__ static __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
  // 0xFF7F / 0x7F7F: largest finite bf16
  __host__ __device__ static float lowest() { return -3.3895313892515355e+38f; }
  __host__ __device__ static float highest() { return 3.3895313892515355e+38f; }
};

// Warp / block scans (256-thread blocks).
template <typename T>
__device__ __forceinline__ T warp_inclusive_scan(T v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T u = __shfl_up_sync(0xffffffffu, v, o);
    if (lane_id() >= o) v += u;
  }
  return v;
}

// Exclusive scan of one value per thread over a 256-thread block.
// warp_sums: shared T[9].  Ends with a barrier, so warp_sums is reusable.
template <typename T>
__device__ __forceinline__ T block_exclusive_scan_256(T v, T* warp_sums, T* total) {
  const int w = threadIdx.x >> 5;
  T inc = warp_inclusive_scan(v);
  if (lane_id() == 31) warp_sums[w] = inc;
  __syncthreads();
  if (w == 0) {
    T s = (lane_id() < 8) ? warp_sums[lane_id()] : T(0);
    T si = warp_inclusive_scan(s);
    if (lane_id() < 8) warp_sums[lane_id()] = si - s;
    if (lane_id() == 7) warp_sums[8] = si;
  }
  __syncthreads();
  T r = inc - v + warp_sums[w];
  *total = warp_sums[8];
  __syncthreads();
  return r;
}

// ------------------------------------------------- generic device scans ---
// Exclusive prefix sum of n int64 values (in place allowed). ws from
// scan_workspace_elems(n) int64 elements.
size_t scan_workspace_elems(int64_t n);
int exclusive_scan_i64(const int64_t* in, int64_t* out, int64_t n, int64_t* ws,
                       cudaStream_t s);
int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, int32_t* ws,
                       cudaStream_t s);

}  // namespace gno
