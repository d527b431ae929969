// cuda_emul.h — TEST INFRASTRUCTURE ONLY (never built into, loaded by or shipped with the product).
//
// A ~150-line host shim of the CUDA execution model, just wide enough to compile the HBM-bound byte / index kernels of
// csrc/train_glue.cu, csrc/infer.cu, csrc/augment.cu (staged) and csrc/seg_loss.cu, csrc/morph.cu (GPU-verified: they validate
// the shim itself) with g++ and run them in the GPU-less build container:
//   * a launch runs the blocks one after the other; the threads of a block run
//       - sequentially when the kernel has no barrier / shuffle (any interleaving of independent threads is legal, and
//         for the lock-free union-find the sequential one is a legal schedule too), or
//       - as blockDim.x real threads with a std::barrier when it is listed as cooperative (__syncthreads, __shfl_xor_sync);
//   * threadIdx is thread_local, blockIdx / gridDim / blockDim are set per block, __shared__ is function-local static;
//   * atomics are GCC __atomic builtins, rounding intrinsics are plain IEEE operations (build with -ffp-contract=off).
// tests/emul/build_emul.py rewrites `kernel<<<grid, block, smem, stream>>>(args)` into EMU_LAUNCH(...) and swaps the
// rsb_common.cuh include for this header.  It exists to catch indexing / logic errors before the first run on a B200; it
// says nothing about performance and is not a CPU fallback.
#pragma once
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdarg>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define RSB_DEVICE inline
#define __builtin_assume(x) ((void)0)
#define __isGlobal(p) true

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
inline thread_local dim3 threadIdx;
inline dim3 blockIdx, gridDim, blockDim;
typedef void* cudaStream_t;
typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }

struct float4 { float x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct uint4 { unsigned x, y, z, w; };
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
struct uint2 { unsigned x, y; };
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
struct int4 { int x, y, z, w; };
inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
struct float2 { float x, y; };
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)}; }

inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
using std::min;
using std::max;
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
inline float __expf(float x) { return std::exp(x); }

// bf16 storage type: round-to-nearest-even conversion like cvt.rn.bf16.f32
struct __nv_bfloat16 { uint16_t bits; };
inline __nv_bfloat16 __float2bfloat16_rn(float f) {
  unsigned u = __float_as_uint(f);
  if ((u & 0x7fffffffu) > 0x7f800000u) return __nv_bfloat16{static_cast<uint16_t>((u >> 16) | 0x40)};   // NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  return __nv_bfloat16{static_cast<uint16_t>(u >> 16)};
}
inline float __bfloat162float(__nv_bfloat16 b) { return __uint_as_float(static_cast<unsigned>(b.bits) << 16); }
struct __nv_bfloat162 { __nv_bfloat16 x, y; };
inline __nv_bfloat162 __floats2bfloat162_rn(float lo, float hi) { return __nv_bfloat162{__float2bfloat16_rn(lo), __float2bfloat16_rn(hi)}; }

// ---- cooperative-block machinery -----------------------------------------------------------------------------------
namespace emu {
inline std::barrier<>* g_barrier = nullptr;      // non-null while a cooperative block runs
inline std::vector<std::barrier<>*> g_warp_barrier;   // one per warp: shuffles synchronise a warp, not the block
inline std::vector<double> g_xchg;               // shuffle exchange slots (one per thread)
inline std::vector<double> g_dyn_smem;           // dynamic shared memory of the running block (8-byte aligned)
inline int g_or_flag[2] = {0, 0};
inline bool g_coop = false;
}  // namespace emu

inline void __syncthreads() {
  if (emu::g_coop) emu::g_barrier->arrive_and_wait();   // sequential mode: kernels with barriers are never run that way
}

inline int __syncthreads_or(int pred) {
  // two alternating accumulators so that back-to-back calls cannot mix
  static thread_local int phase = 0;
  const int p = phase;
  phase ^= 1;
  if (pred) __atomic_store_n(&emu::g_or_flag[p], 1, __ATOMIC_RELAXED);
  emu::g_barrier->arrive_and_wait();
  const int r = __atomic_load_n(&emu::g_or_flag[p], __ATOMIC_RELAXED);
  emu::g_barrier->arrive_and_wait();
  if (threadIdx.x == 0) emu::g_or_flag[p] = 0;
  emu::g_barrier->arrive_and_wait();
  return r;
}

template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
  static_assert(sizeof(T) <= sizeof(double), "shuffle payload");
  double slot = 0;
  std::memcpy(&slot, &v, sizeof(T));
  std::barrier<>* wb = emu::g_warp_barrier[threadIdx.x >> 5];
  emu::g_xchg[threadIdx.x] = slot;
  wb->arrive_and_wait();
  const unsigned src = (threadIdx.x & ~31u) | ((threadIdx.x & 31u) ^ static_cast<unsigned>(lane_mask));
  slot = emu::g_xchg[src];
  wb->arrive_and_wait();
  T out;
  std::memcpy(&out, &slot, sizeof(T));
  return out;
}

template <typename T>
inline T __shfl_sync(unsigned, T v, int src_lane) {
  static_assert(sizeof(T) <= sizeof(double), "shuffle payload");
  double slot = 0;
  std::memcpy(&slot, &v, sizeof(T));
  std::barrier<>* wb = emu::g_warp_barrier[threadIdx.x >> 5];
  emu::g_xchg[threadIdx.x] = slot;
  wb->arrive_and_wait();
  slot = emu::g_xchg[(threadIdx.x & ~31u) | (static_cast<unsigned>(src_lane) & 31u)];
  wb->arrive_and_wait();
  T out;
  std::memcpy(&out, &slot, sizeof(T));
  return out;
}

inline unsigned __ballot_sync(unsigned, int pred) {
  std::barrier<>* wb = emu::g_warp_barrier[threadIdx.x >> 5];
  emu::g_xchg[threadIdx.x] = pred ? 1.0 : 0.0;
  wb->arrive_and_wait();
  unsigned mask = 0u;
  const unsigned base = threadIdx.x & ~31u;
  for (unsigned l = 0; l < 32 && base + l < blockDim.x; ++l)
    if (emu::g_xchg[base + l] != 0.0) mask |= 1u << l;
  wb->arrive_and_wait();
  return mask;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0u; }

inline unsigned emu_block(unsigned b) { return b; }
inline unsigned emu_block(int b) { return static_cast<unsigned>(b); }
inline unsigned emu_block(dim3 b) { return b.x; }

template <typename F>
inline void emu_launch(bool coop, dim3 grid, unsigned block, size_t smem_bytes, F&& body) {
  gridDim = grid;
  blockDim.x = block;
  emu::g_dyn_smem.assign(smem_bytes / 8 + 2, 0.0);
  for (unsigned b = 0; b < grid.x * grid.y * grid.z; ++b) {
    blockIdx.x = b % grid.x;
    blockIdx.y = (b / grid.x) % grid.y;
    blockIdx.z = b / (grid.x * grid.y);
    if (!coop) {
      for (unsigned t = 0; t < block; ++t) {
        threadIdx.x = t;
        body();
      }
      continue;
    }
    std::barrier<> bar(block);
    emu::g_barrier = &bar;
    std::vector<std::unique_ptr<std::barrier<>>> warps;
    emu::g_warp_barrier.clear();
    for (unsigned w = 0; w * 32 < block; ++w) {
      warps.emplace_back(new std::barrier<>(std::min(32u, block - w * 32)));
      emu::g_warp_barrier.push_back(warps.back().get());
    }
    emu::g_xchg.assign(block, 0.0);
    emu::g_coop = true;
    std::vector<std::thread> th;
    th.reserve(block);
    for (unsigned t = 0; t < block; ++t)
      th.emplace_back([t, &body]() {
        threadIdx.x = t;
        body();
      });
    for (auto& x : th) x.join();
    emu::g_coop = false;
    emu::g_barrier = nullptr;
  }
  threadIdx.x = 0;
}

#define EMU_LAUNCH(coop, kernel, grid, block, smem, ...) emu_launch(coop, dim3(grid), emu_block(block), static_cast<size_t>(smem), [&]() { kernel(__VA_ARGS__); })

// ---- intrinsics --------------------------------------------------------------------------------------------------
template <typename T> inline T __ldg(const T* p) { return *p; }
inline int __ldcg(const int* p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fdiv_rn(float a, float b) { return a / b; }

inline int atomicMin(int* p, int v) {
  int old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline float atomicAdd(float* p, float v) {
  float old = *p, want;
  do { want = old + v; } while (!__atomic_compare_exchange(p, &old, &want, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
  return old;
}
inline int atomicMax(int* p, int v) {
  int old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <typename K> inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
enum cudaMemcpyKind { cudaMemcpyDeviceToDevice = 3 };
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(d, s, n); return cudaSuccess; }
inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) {
  unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}

// ---- the host plumbing of rsb_common.cuh / api.cu ------------------------------------------------------------------------
namespace rsb {
inline thread_local char g_last_error[512] = "";
inline void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}
inline int check_launch(const char*) { return 0; }
}  // namespace rsb
extern "C" int rsb_num_sms(void);

#define RSB_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      ::rsb::set_last_error(__VA_ARGS__); \
      return -1;                          \
    }                                     \
  } while (0)
