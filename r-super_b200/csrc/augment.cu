// augment.cu — the six online intensity augmentations of the reference loader on the device (second half):
// dataset_abdomenatlas_UFO.py:1048-1061 applies, each with probability 0.3 and in this order,
//   brightness_multiply  x * f                                         training/augmentation.py:86-103
//   brightness_additive  x + a                                         :69-83
//   gamma                pow((x - min) / rng, g) * rng + min, then re-standardised to the input's mean / std   :106-138
//   contrast             clamp((x - mean) * f + mean, min, max)        :140-168
//   gaussian_blur        conv3d with the normalised (2*ceil(3 sigma)+1)^3 Gaussian, zero padding                :48-66
//   gaussian_noise       x + noise * std                               :17-19
// The random draws are made by the host mirror (rsuper_b200/augment.py) exactly as the reference makes them and arrive
// here as scalars; the kernels are the deterministic part.  All HBM bound on one fp32 volume: whole-volume statistics are a
// two-stage deterministic reduction in double (block partials, then one block), elementwise passes are one read + one write,
// the blur is three separable axis passes (the reference's 3-D kernel is exactly the outer product of its 1-D marginals).
// Products and sums that torch performs as separate roundings use __fmul_rn / __fadd_rn so that no FMA contraction changes
// the last bit of the bit-exact ops (multiply, additive, noise, contrast).
#include "rsb_common.cuh"

#include "../../include/rsuper_b200.h"

namespace rsb {

constexpr int AUG_THREADS = 256;
constexpr int AUG_MAX_BLOCKS = 148 * 8;
constexpr int AUG_MAX_TAPS = 33;

static inline int aug_grid(long long n) {
  const long long b = (n + AUG_THREADS - 1) / AUG_THREADS;
  return static_cast<int>(b < 1 ? 1 : (b > AUG_MAX_BLOCKS ? AUG_MAX_BLOCKS : b));
}

struct AugPartial {
  double mn, mx, sum, sumsq;
};

RSB_DEVICE void aug_block_reduce(AugPartial& v, AugPartial* sh /* AUG_THREADS / 32 */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v.mn = fmin(v.mn, __shfl_xor_sync(0xffffffffu, v.mn, o));
    v.mx = fmax(v.mx, __shfl_xor_sync(0xffffffffu, v.mx, o));
    v.sum += __shfl_xor_sync(0xffffffffu, v.sum, o);
    v.sumsq += __shfl_xor_sync(0xffffffffu, v.sumsq, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    AugPartial t = sh[0];
    for (int w = 1; w < AUG_THREADS / 32; ++w) {
      t.mn = fmin(t.mn, sh[w].mn);
      t.mx = fmax(t.mx, sh[w].mx);
      t.sum += sh[w].sum;
      t.sumsq += sh[w].sumsq;
    }
    v = t;
  }
}

__global__ void __launch_bounds__(AUG_THREADS) aug_stats_partial_kernel(const float* __restrict__ x, long long n,
                                                                         AugPartial* __restrict__ partials) {
  __shared__ AugPartial sh[AUG_THREADS / 32];
  AugPartial v{1e300, -1e300, 0.0, 0.0};
  for (long long i = blockIdx.x * static_cast<long long>(AUG_THREADS) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * AUG_THREADS) {
    const double d = static_cast<double>(x[i]);
    v.mn = fmin(v.mn, d);
    v.mx = fmax(v.mx, d);
    v.sum += d;
    v.sumsq += d * d;
  }
  aug_block_reduce(v, sh);
  if (threadIdx.x == 0) partials[blockIdx.x] = v;
}

__global__ void __launch_bounds__(AUG_THREADS) aug_stats_final_kernel(const AugPartial* __restrict__ partials, int n_partials,
                                                                       long long n, float* __restrict__ stats4) {
  __shared__ AugPartial sh[AUG_THREADS / 32];
  AugPartial v{1e300, -1e300, 0.0, 0.0};
  for (int i = threadIdx.x; i < n_partials; i += AUG_THREADS) {
    const AugPartial p = partials[i];
    v.mn = fmin(v.mn, p.mn);
    v.mx = fmax(v.mx, p.mx);
    v.sum += p.sum;
    v.sumsq += p.sumsq;
  }
  aug_block_reduce(v, sh);
  if (threadIdx.x == 0) {
    const double mean = v.sum / static_cast<double>(n);
    // torch.std: unbiased (n - 1)
    const double var = n > 1 ? fmax(v.sumsq - v.sum * mean, 0.0) / static_cast<double>(n - 1) : 0.0;
    stats4[0] = static_cast<float>(v.mn);
    stats4[1] = static_cast<float>(v.mx);
    stats4[2] = static_cast<float>(mean);
    stats4[3] = static_cast<float>(sqrt(var));
  }
}

__global__ void __launch_bounds__(AUG_THREADS) aug_affine_kernel(const float* __restrict__ x, float* __restrict__ y, long long n,
                                                                  float mul, int has_mul, float add, int has_add,
                                                                  const float* __restrict__ noise, float noise_std) {
  for (long long i = blockIdx.x * static_cast<long long>(AUG_THREADS) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * AUG_THREADS) {
    float v = x[i];
    if (has_mul) v = __fmul_rn(v, mul);
    if (has_add) v = __fadd_rn(v, add);
    if (noise != nullptr) v = __fadd_rn(v, __fmul_rn(noise[i], noise_std));
    y[i] = v;
  }
}

__global__ void __launch_bounds__(AUG_THREADS) aug_gamma_kernel(const float* __restrict__ x, float* __restrict__ y, long long n,
                                                                 const float* __restrict__ stats4, float gamma) {
  const float mn = stats4[0], rng = __fsub_rn(stats4[1], stats4[0]);
  for (long long i = blockIdx.x * static_cast<long long>(AUG_THREADS) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * AUG_THREADS) {
    const float t = __fdiv_rn(__fsub_rn(x[i], mn), rng);
    y[i] = __fadd_rn(__fmul_rn(powf(t, gamma), rng), mn);
  }
}

__global__ void __launch_bounds__(AUG_THREADS) aug_renorm_kernel(float* __restrict__ y, long long n, const float* __restrict__ stats_y,
                                                                  const float* __restrict__ stats_x) {
  const float my = stats_y[2], sy = stats_y[3], mx = stats_x[2], sx = stats_x[3];
  for (long long i = blockIdx.x * static_cast<long long>(AUG_THREADS) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * AUG_THREADS)
    y[i] = __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(y[i], my), sy), sx), mx);  // (y - mean_y) / std_y * std + mean
}

__global__ void __launch_bounds__(AUG_THREADS) aug_contrast_kernel(const float* __restrict__ x, float* __restrict__ y, long long n,
                                                                    const float* __restrict__ stats4, float factor) {
  const float mn = stats4[0], mx = stats4[1], mean = stats4[2];
  for (long long i = blockIdx.x * static_cast<long long>(AUG_THREADS) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * AUG_THREADS) {
    const float v = __fadd_rn(__fmul_rn(__fsub_rn(x[i], mean), factor), mean);
    y[i] = fminf(fmaxf(v, mn), mx);
  }
}

struct AugTaps {
  float w[AUG_MAX_TAPS];
};

__global__ void __launch_bounds__(AUG_THREADS) aug_blur_axis_kernel(const float* __restrict__ x, float* __restrict__ y, int n_vol,
                                                                     int D, int H, int W, int axis, AugTaps taps, int ntaps) {
  const long long V = static_cast<long long>(D) * H * W;
  const long long total = static_cast<long long>(n_vol) * V;
  const int half = ntaps >> 1;
  const int len = axis == 0 ? D : (axis == 1 ? H : W);
  const long long stride = axis == 0 ? static_cast<long long>(H) * W : (axis == 1 ? W : 1);
  for (long long i = blockIdx.x * static_cast<long long>(AUG_THREADS) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * AUG_THREADS) {
    const long long v = i % V;
    const int pos = axis == 0 ? static_cast<int>(v / (static_cast<long long>(H) * W))
                              : (axis == 1 ? static_cast<int>((v / W) % H) : static_cast<int>(v % W));
    float acc = 0.f;
    for (int t = 0; t < ntaps; ++t) {
      const int q = pos + t - half;
      if (q >= 0 && q < len) acc += taps.w[t] * x[i + static_cast<long long>(t - half) * stride];  // zero padding
    }
    y[i] = acc;
  }
}

}  // namespace rsb

using namespace rsb;

extern "C" size_t rsb_aug_workspace_bytes(void) { return sizeof(AugPartial) * AUG_MAX_BLOCKS; }

extern "C" int rsb_aug_stats(const float* x, long long n, void* workspace, float* stats4, void* stream) {
  RSB_REQUIRE(x != nullptr && workspace != nullptr && stats4 != nullptr, "aug_stats: null pointer");
  RSB_REQUIRE(n > 0, "aug_stats: empty volume");
  RSB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "aug_stats: workspace must be 8-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = aug_grid(n);
  AugPartial* part = static_cast<AugPartial*>(workspace);
  aug_stats_partial_kernel<<<grid, AUG_THREADS, 0, st>>>(x, n, part);
  if (int rc = check_launch("aug_stats_partial_kernel")) return rc;
  aug_stats_final_kernel<<<1, AUG_THREADS, 0, st>>>(part, grid, n, stats4);
  return check_launch("aug_stats_final_kernel");
}

extern "C" int rsb_aug_affine(const float* x, float* y, long long n, float mul, int has_mul, float add, int has_add,
                              const float* noise, float noise_std, void* stream) {
  RSB_REQUIRE(x != nullptr && y != nullptr && n > 0, "aug_affine: null pointer / empty volume");
  aug_affine_kernel<<<aug_grid(n), AUG_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(x, y, n, mul, has_mul, add, has_add, noise,
                                                                                       noise_std);
  return check_launch("aug_affine_kernel");
}

extern "C" int rsb_aug_gamma(const float* x, float* y, long long n, const float* stats4, float gamma, void* stream) {
  RSB_REQUIRE(x != nullptr && y != nullptr && stats4 != nullptr && n > 0, "aug_gamma: null pointer / empty volume");
  aug_gamma_kernel<<<aug_grid(n), AUG_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(x, y, n, stats4, gamma);
  return check_launch("aug_gamma_kernel");
}

extern "C" int rsb_aug_renorm(float* y, long long n, const float* stats_y, const float* stats_x, void* stream) {
  RSB_REQUIRE(y != nullptr && stats_y != nullptr && stats_x != nullptr && n > 0, "aug_renorm: null pointer / empty volume");
  aug_renorm_kernel<<<aug_grid(n), AUG_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(y, n, stats_y, stats_x);
  return check_launch("aug_renorm_kernel");
}

extern "C" int rsb_aug_contrast(const float* x, float* y, long long n, const float* stats4, float factor, void* stream) {
  RSB_REQUIRE(x != nullptr && y != nullptr && stats4 != nullptr && n > 0, "aug_contrast: null pointer / empty volume");
  aug_contrast_kernel<<<aug_grid(n), AUG_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(x, y, n, stats4, factor);
  return check_launch("aug_contrast_kernel");
}

extern "C" int rsb_aug_blur_axis(const float* x, float* y, int n_vol, int D, int H, int W, int axis, const float* taps_host,
                                 int ntaps, void* stream) {
  RSB_REQUIRE(x != nullptr && y != nullptr && x != y && taps_host != nullptr, "aug_blur_axis: needs distinct source / destination and taps");
  RSB_REQUIRE(n_vol > 0 && D > 0 && H > 0 && W > 0 && axis >= 0 && axis <= 2, "aug_blur_axis: bad geometry");
  RSB_REQUIRE(ntaps >= 1 && ntaps <= AUG_MAX_TAPS && (ntaps & 1), "aug_blur_axis: ntaps must be odd and <= %d (got %d)", AUG_MAX_TAPS, ntaps);
  AugTaps t;
  for (int i = 0; i < AUG_MAX_TAPS; ++i) t.w[i] = i < ntaps ? taps_host[i] : 0.f;   // HOST pointer: passed by value to the kernel
  aug_blur_axis_kernel<<<aug_grid(static_cast<long long>(n_vol) * D * H * W), AUG_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      x, y, n_vol, D, H, W, axis, t, ntaps);
  return check_launch("aug_blur_axis_kernel");
}
