// rsb_common.cuh — shared device helpers for the rsuper_b200 kernels (sm_100a only).
//
// Thin inline-PTX wrappers for mbarrier, bulk async copy (TMA 1D), tcgen05 (alloc / mma /
// commit / ld / fences) plus small packing helpers.  Nothing here is a port of the reference:
// MrGiovanni/R-Super ships no native code; every kernel in csrc/ replaces an
// ATen / cuDNN library call made by rsuper_train/model/dim3/*.py or
// rsuper_train/training/losses_foundation.py.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#ifndef RSB_DEVICE
#define RSB_DEVICE __device__ __forceinline__
#endif

namespace rsb {

// ------------------------------------------------------------------------------------------
// error plumbing (host)
// ------------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);
int check_launch(const char* what);

#define RSB_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      ::rsb::set_last_error(__VA_ARGS__);      \
      return -1;                               \
    }                                          \
  } while (0)

// ------------------------------------------------------------------------------------------
// generic helpers
// ------------------------------------------------------------------------------------------
RSB_DEVICE uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

RSB_DEVICE bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------
RSB_DEVICE void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
RSB_DEVICE void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
RSB_DEVICE void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
RSB_DEVICE void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
RSB_DEVICE bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a protocol bug becomes a trap (reported as a launch failure) instead of a hung box.
RSB_DEVICE void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) {
      printf("rsb mbar_wait timeout: block %d thread %d barrier smem+0x%x parity %u\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// ------------------------------------------------------------------------------------------
// proxies / fences
// ------------------------------------------------------------------------------------------
// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma / bulk copies)
RSB_DEVICE void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
RSB_DEVICE void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
RSB_DEVICE void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ------------------------------------------------------------------------------------------
// bulk async copy global -> shared (TMA without a tensor map; SASS: UBLKCP)
// ------------------------------------------------------------------------------------------
RSB_DEVICE void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

// ------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA, commit, load
// ------------------------------------------------------------------------------------------
RSB_DEVICE void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(ncols)
               : "memory");
}
RSB_DEVICE void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
RSB_DEVICE void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}

// Shared-memory matrix descriptor, SWIZZLE_NONE ("interleave") canonical layouts
// (sm_100 descriptor: start[0,14) lbo[16,30) sbo[32,46) version=1 at [46,48), layout_type[61,64)=0).
//   K-major : ((8,m),(T,2)) : ((1T,SBO),(1,LBO))      core matrix = 8 rows x 16 B, rows 16 B apart
//   MN-major: ((T,1,m),(8,k)) : ((1,T,SBO),(1T,LBO))  core matrix = 8 k    x 16 B, k's  16 B apart
// In both cases SBO strides core matrices along M/N and LBO strides them along K.
RSB_DEVICE uint64_t make_desc_noswz(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell)
  return d;
}

// Instruction descriptor for kind::f16 with BF16 A/B and FP32 accumulate.
//   c_format[4,6)=1 (F32)  a_format[7,10)=1 (BF16)  b_format[10,13)=1 (BF16)
//   a_major bit15, b_major bit16 (0 = K-major, 1 = MN-major)   n_dim[17,23)=N>>3   m_dim[24,29)=M>>4
__host__ __device__ inline uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 1u << 7;
  d |= 1u << 10;
  d |= (a_mn_major ? 1u : 0u) << 15;
  d |= (b_mn_major ? 1u : 0u) << 16;
  d |= (static_cast<uint32_t>(N >> 3) & 0x3Fu) << 17;
  d |= (static_cast<uint32_t>(M >> 4) & 0x1Fu) << 24;
  return d;
}

// D[tmem] (+)= A[smem] * B[smem]   (single CTA, one thread issues)
RSB_DEVICE void umma_bf16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// All previously issued MMAs of this thread arrive (once) on the mbarrier when complete.
RSB_DEVICE void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}

// TMEM -> registers: 32 lanes x 16 consecutive fp32 columns (thread i gets lane base+i).
RSB_DEVICE void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
RSB_DEVICE void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// named barriers (sub-CTA sync)
// ------------------------------------------------------------------------------------------
RSB_DEVICE void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------------------------------
// packing / vector access
// ------------------------------------------------------------------------------------------
RSB_DEVICE uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
RSB_DEVICE float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
RSB_DEVICE float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

// 8 consecutive channels of a voxel as fp32, from bf16 or fp32 storage.
template <typename T>
struct Vec8;
template <>
struct Vec8<__nv_bfloat16> {
  static constexpr int kBytes = 16;
  RSB_DEVICE static void load(const __nv_bfloat16* p, float (&f)[8]) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x);
    f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
    f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z);
    f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
  }
  RSB_DEVICE static void store(__nv_bfloat16* p, const float (&f)[8]) {
    uint4 u;
    u.x = pack_bf16x2(f[0], f[1]);
    u.y = pack_bf16x2(f[2], f[3]);
    u.z = pack_bf16x2(f[4], f[5]);
    u.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};
template <>
struct Vec8<float> {
  static constexpr int kBytes = 32;
  RSB_DEVICE static void load(const float* p, float (&f)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  RSB_DEVICE static void store(float* p, const float (&f)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
  }
};

// Unconverted 8 / 16-channel payloads: lets a kernel issue global loads early and convert late
// (software prefetch without the convert instruction forcing an early wait on the load).
template <typename T>
struct Raw8;
template <>
struct Raw8<__nv_bfloat16> {
  uint4 u;
  RSB_DEVICE void load(const __nv_bfloat16* p) { u = *reinterpret_cast<const uint4*>(p); }
  RSB_DEVICE void zero() { u = make_uint4(0u, 0u, 0u, 0u); }
  RSB_DEVICE void to_float(float (&f)[8]) const {
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x);
    f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
    f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z);
    f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
  }
};
template <>
struct Raw8<float> {
  float4 a, b;
  RSB_DEVICE void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p);
    b = *reinterpret_cast<const float4*>(p + 4);
  }
  RSB_DEVICE void zero() { a = make_float4(0.f, 0.f, 0.f, 0.f); b = a; }
  RSB_DEVICE void to_float(float (&f)[8]) const {
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
};
template <typename T>
struct Raw16 {
  Raw8<T> lo, hi;
  RSB_DEVICE void load(const T* p, bool both) {
    lo.load(p);
    if (both) hi.load(p + 8); else hi.zero();
  }
  RSB_DEVICE void zero() { lo.zero(); hi.zero(); }
  RSB_DEVICE void to_float(float (&f)[16]) const {
    float t[8];
    lo.to_float(t);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = t[j];
    hi.to_float(t);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[8 + j] = t[j];
  }
};

RSB_DEVICE float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// InstanceNorm statistics are carried as raw (sum, sumsq) pairs; consumers derive mean / rstd.
// eps follows nn.InstanceNorm3d(ch, eps=1e-4) — rsuper_train/model/dim3/conv_layers.py:39-42.
RSB_DEVICE void stats_to_mean_rstd(float s, float ss, float inv_count, float eps, float& mean,
                                   float& rstd) {
  mean = s * inv_count;
  float var = fmaf(-mean, mean, ss * inv_count);
  var = fmaxf(var, 0.f);
  rstd = rsqrtf(var + eps);
}

}  // namespace rsb
