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

// ---- cooperative-block machinery -----------------------------------------------------------------------------------
namespace emu {
inline std::barrier<>* g_barrier = nullptr;      // non-null while a cooperative block runs
inline std::vector<std::barrier<>*> g_warp_barrier;   // one per warp: shuffles synchronise a warp, not the block
inline std::vector<double> g_xchg;               // shuffle exchange slots (one per thread)
inline bool g_coop = false;
}  // namespace emu

inline void __syncthreads() {
  if (emu::g_coop) emu::g_barrier->arrive_and_wait();   // sequential mode: kernels with barriers are never run that way
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

template <typename F>
inline void emu_launch(bool coop, dim3 grid, unsigned block, F&& body) {
  gridDim = grid;
  blockDim.x = block;
  for (unsigned b = 0; b < grid.x * grid.y; ++b) {
    blockIdx.x = b % grid.x;
    blockIdx.y = b / grid.x;
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

#define EMU_LAUNCH(coop, kernel, grid, block, ...) emu_launch(coop, dim3(grid), static_cast<unsigned>(block), [&]() { kernel(__VA_ARGS__); })

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
inline float atomicAdd(float* p, float v) {
  float old = *p, want;
  do { want = old + v; } while (!__atomic_compare_exchange(p, &old, &want, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
  return old;
}
inline float warp_sum_emu(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
namespace rsb { inline float warp_sum(float v) { return warp_sum_emu(v); } }
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

#define RSB_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      ::rsb::set_last_error(__VA_ARGS__); \
      return -1;                          \
    }                                     \
  } while (0)
