// elementwise.cu — HBM-bound satellites of the UNet train step on NDHWC tensors with a channel
// pitch: layout conversion, per-(n,c) statistics, MaxPool3d(2) fwd/bwd, trilinear(align_corners)
// upsample fwd/bwd, and the InstanceNorm backward apply.  All kernels move 8 channels per thread
// with 128-bit (bf16) / 2x128-bit (fp32) accesses; consecutive threads walk consecutive channel
// groups of a voxel, then the next voxel, so every warp access is a contiguous span.
//
// Reference call sites replaced (rsuper_train/model/dim3):
//   nn.MaxPool3d(down_scale)                                   unet_utils.py:36
//   F.interpolate(mode='trilinear', align_corners=True)        unet_utils.py:69
//   nn.InstanceNorm3d(eps=1e-4) backward (autograd)            conv_layers.py:39-49
#include "rsb_common.cuh"

#include "../../include/rsuper_b200.h"

namespace rsb {

// Thread mapping shared by all kernels here: blockIdx.y = sample n; inside a sample the linear
// element id e = voxel * CG + cg (CG = C/8 channel groups); blockDim.x is a multiple of CG so a
// thread keeps the same channel group over its grid-stride loop.
struct ClMap {
  int cg;         // channel group of this thread
  long long v0;   // first voxel
  long long vstride;  // voxels per grid sweep (blockDim.x % CG == 0, so a thread keeps its channel group)
};
RSB_DEVICE ClMap cl_map(int CG) {
  ClMap m;
  const unsigned vpb = blockDim.x / CG;  // voxels per block
  m.cg = static_cast<int>(threadIdx.x % CG);
  m.v0 = static_cast<long long>(blockIdx.x) * vpb + threadIdx.x / CG;
  m.vstride = static_cast<long long>(gridDim.x) * vpb;
  return m;
}

static inline int cl_block(int CG) {
  if (CG >= 256) return CG <= 1024 ? CG : 0;
  return (256 / CG) * CG;
}
static inline int cl_grid(long long elems, int block, int sms, int waves = 8) {
  long long want = (elems + block - 1) / block;
  long long cap = static_cast<long long>(sms) * waves;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return static_cast<int>(want);
}

// Block-level accumulation of per-channel (a, b) pairs into global stats[(n*C + c)*2 + {0,1}].
// sm_acc must hold 2*C floats, zeroed before use (done here) — one shared atomic per value per
// thread, one global atomic per value per block.
RSB_DEVICE void block_stats_flush(float* sm_acc, float (&s1)[8], float (&s2)[8], int cg,
                                  int C, float* stats_n) {
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm_acc[i] = 0.f;
  // Lanes l, l + CG, l + 2 CG, ... of a warp hold partials of the SAME channels: reduce them with shuffles first.  Going
  // straight to shared-memory atomics makes every atomic an 8-way same-address conflict at C = 32 — ncu showed that
  // serialisation (short-scoreboard + barrier stalls), not HBM, bounding maxpool / channel_stats / upsample forward.
  const int CG = C / 8;
  bool owner = true;
  if (CG < 32 && (CG & (CG - 1)) == 0) {
    for (int o = 16; o >= CG; o >>= 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
        s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o);
      }
    }
    owner = (threadIdx.x & 31) < CG;
  }
  __syncthreads();
  if (owner) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&sm_acc[(cg * 8 + j) * 2 + 0], s1[j]);
      atomicAdd(&sm_acc[(cg * 8 + j) * 2 + 1], s2[j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&stats_n[i], sm_acc[i]);
}

// ------------------------------------------------------------------------------------------
// layout conversion
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void ncdhw_to_ndhwc_kernel(const float* __restrict__ src, T* __restrict__ dst, long long pitch,
                                      int C, long long V) {
  const int CG = C / 8;
  const int n = blockIdx.y;
  ClMap m = cl_map(CG);
  for (long long v = m.v0; v < V; v += m.vstride) {
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = src[(static_cast<long long>(n) * C + m.cg * 8 + j) * V + v];
    Vec8<T>::store(dst + (static_cast<long long>(n) * V + v) * pitch + m.cg * 8, f);
  }
}

template <typename T>
__global__ void ndhwc_to_ncdhw_kernel(const T* __restrict__ src, long long pitch, float* __restrict__ dst,
                                      int C, long long V) {
  const int CG = C / 8;
  const int n = blockIdx.y;
  ClMap m = cl_map(CG);
  for (long long v = m.v0; v < V; v += m.vstride) {
    float f[8];
    Vec8<T>::load(src + (static_cast<long long>(n) * V + v) * pitch + m.cg * 8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[(static_cast<long long>(n) * C + m.cg * 8 + j) * V + v] = f[j];
  }
}

template <typename T>
__global__ void channel_stats_kernel(const T* __restrict__ x, long long pitch, float* __restrict__ stats,
                                     int C, long long V) {
  extern __shared__ float sm_acc[];
  const int CG = C / 8;
  const int n = blockIdx.y;
  ClMap m = cl_map(CG);
  float s1[8] = {0}, s2[8] = {0};
  for (long long v = m.v0; v < V; v += m.vstride) {
    float f[8];
    Vec8<T>::load(x + (static_cast<long long>(n) * V + v) * pitch + m.cg * 8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1[j] += f[j]; s2[j] = fmaf(f[j], f[j], s2[j]); }
  }
  block_stats_flush(sm_acc, s1, s2, m.cg, C, stats + static_cast<long long>(n) * pitch * 2);
}

// ------------------------------------------------------------------------------------------
// MaxPool3d(2)
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void maxpool2_fwd_kernel(const T* __restrict__ x, long long xp, T* __restrict__ y, long long yp,
                                    float* __restrict__ stats, int D, int H, int W, int C) {
  extern __shared__ float sm_acc[];
  const int CG = C / 8;
  const int n = blockIdx.y;
  const int Do = D / 2, Ho = H / 2, Wo = W / 2;
  const long long Vo = static_cast<long long>(Do) * Ho * Wo;
  ClMap m = cl_map(CG);
  float s1[8] = {0}, s2[8] = {0};
  for (long long v = m.v0; v < Vo; v += m.vstride) {
    const unsigned vu = static_cast<unsigned>(v);   // per-sample voxel index: host checks Vo < 2^31
    const unsigned row = vu / static_cast<unsigned>(Wo);
    const int xo = static_cast<int>(vu - row * Wo);
    const int zo = static_cast<int>(row / static_cast<unsigned>(Ho));
    const int yo = static_cast<int>(row - static_cast<unsigned>(zo) * Ho);
    float best[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) best[j] = -INFINITY;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int z = 2 * zo + (k >> 2), yy = 2 * yo + ((k >> 1) & 1), xx = 2 * xo + (k & 1);
      const long long vin = ((static_cast<long long>(n) * D + z) * H + yy) * W + xx;
      float f[8];
      Vec8<T>::load(x + vin * xp + m.cg * 8, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) best[j] = fmaxf(best[j], f[j]);
    }
    Vec8<T>::store(y + (static_cast<long long>(n) * Vo + v) * yp + m.cg * 8, best);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1[j] += best[j]; s2[j] = fmaf(best[j], best[j], s2[j]); }
  }
  if (stats != nullptr) block_stats_flush(sm_acc, s1, s2, m.cg, C, stats + static_cast<long long>(n) * yp * 2);
}

// ------------------------------------------------------------------------------------------
// ConvTranspose3d(kernel = stride = 2) epilogues: the transposed conv is a 1x1x1 conv to 8 * C channels at the input
// resolution (column block k = 4a + 2b + c holds the output voxel (2z + a, 2y + b, 2x + c)); these two kernels move between
// that layout and the up-sampled NDHWC tensor (8 channels = one 128-bit access).
//   depth_to_space2:  up[n, 2z+a, 2y+b, 2x+c, co] = q[n, z, y, x, k * C + co] + bias[co]   (+ InstanceNorm statistics of up)
//   space_to_depth2:  q[n, z, y, x, k * C + co]   = up[n, 2z+a, 2y+b, 2x+c, co]           (its adjoint, for the backward pass)
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void depth_to_space2_kernel(const T* __restrict__ q, long long qp, const float* __restrict__ bias, T* __restrict__ up,
                                       long long upp, float* __restrict__ stats, int D, int H, int W, int C) {
  extern __shared__ float sm_acc[];
  const int CG = C / 8;
  const int n = blockIdx.y;
  const long long Vi = static_cast<long long>(D) * H * W;
  ClMap m = cl_map(CG);
  float b8[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) b8[j] = bias != nullptr ? bias[m.cg * 8 + j] : 0.f;
  float s1[8] = {0}, s2[8] = {0};
  for (long long v = m.v0; v < Vi; v += m.vstride) {
    const unsigned vu = static_cast<unsigned>(v);
    const unsigned row = vu / static_cast<unsigned>(W);
    const int xi = static_cast<int>(vu - row * W);
    const int zi = static_cast<int>(row / static_cast<unsigned>(H));
    const int yi = static_cast<int>(row - static_cast<unsigned>(zi) * H);
    const T* src = q + (static_cast<long long>(n) * Vi + v) * qp + m.cg * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float f[8];
      Vec8<T>::load(src + k * C, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += b8[j];
      const int z = 2 * zi + (k >> 2), yy = 2 * yi + ((k >> 1) & 1), xx = 2 * xi + (k & 1);
      const long long vo = ((static_cast<long long>(n) * (2 * D) + z) * (2 * H) + yy) * (2 * W) + xx;
      Vec8<T>::store(up + vo * upp + m.cg * 8, f);
      // statistics of the STORED values' fp32 originals, like every other producer
#pragma unroll
      for (int j = 0; j < 8; ++j) { s1[j] += f[j]; s2[j] = fmaf(f[j], f[j], s2[j]); }
    }
  }
  if (stats != nullptr) block_stats_flush(sm_acc, s1, s2, m.cg, C, stats + static_cast<long long>(n) * upp * 2);
}

template <typename T>
__global__ void space_to_depth2_kernel(const T* __restrict__ up, long long upp, T* __restrict__ q, long long qp, int D, int H, int W,
                                       int C) {
  const int CG = C / 8;
  const int n = blockIdx.y;
  const long long Vi = static_cast<long long>(D) * H * W;
  ClMap m = cl_map(CG);
  for (long long v = m.v0; v < Vi; v += m.vstride) {
    const unsigned vu = static_cast<unsigned>(v);
    const unsigned row = vu / static_cast<unsigned>(W);
    const int xi = static_cast<int>(vu - row * W);
    const int zi = static_cast<int>(row / static_cast<unsigned>(H));
    const int yi = static_cast<int>(row - static_cast<unsigned>(zi) * H);
    T* dst = q + (static_cast<long long>(n) * Vi + v) * qp + m.cg * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int z = 2 * zi + (k >> 2), yy = 2 * yi + ((k >> 1) & 1), xx = 2 * xi + (k & 1);
      const long long vo = ((static_cast<long long>(n) * (2 * D) + z) * (2 * H) + yy) * (2 * W) + xx;
      float f[8];
      Vec8<T>::load(up + vo * upp + m.cg * 8, f);
      Vec8<T>::store(dst + k * C, f);
    }
  }
}

// dx[window voxel k] = (k is the first maximum in (d,h,w) scan order ? dy : 0) + dskip
// (ATen max_pool3d_with_indices keeps the first maximum: it updates only on `val > max`).
template <typename T>
__global__ void maxpool2_bwd_kernel(const T* __restrict__ x, long long xp, const T* __restrict__ dy,
                                    long long dyp, const T* __restrict__ dskip, long long dsp,
                                    T* __restrict__ dx, long long dxp, int D, int H, int W, int C) {
  const int CG = C / 8;
  const int n = blockIdx.y;
  const int Do = D / 2, Ho = H / 2, Wo = W / 2;
  const long long Vo = static_cast<long long>(Do) * Ho * Wo;
  ClMap m = cl_map(CG);
  for (long long v = m.v0; v < Vo; v += m.vstride) {
    const unsigned vu = static_cast<unsigned>(v);   // per-sample voxel index: host checks Vo < 2^31
    const unsigned row = vu / static_cast<unsigned>(Wo);
    const int xo = static_cast<int>(vu - row * Wo);
    const int zo = static_cast<int>(row / static_cast<unsigned>(Ho));
    const int yo = static_cast<int>(row - static_cast<unsigned>(zo) * Ho);
    float best[8];
    int arg[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; arg[j] = 0; }
    long long vin[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int z = 2 * zo + (k >> 2), yy = 2 * yo + ((k >> 1) & 1), xx = 2 * xo + (k & 1);
      vin[k] = ((static_cast<long long>(n) * D + z) * H + yy) * W + xx;
      float f[8];
      Vec8<T>::load(x + vin[k] * xp + m.cg * 8, f);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (f[j] > best[j]) { best[j] = f[j]; arg[j] = k; }
    }
    float g[8];
    Vec8<T>::load(dy + (static_cast<long long>(n) * Vo + v) * dyp + m.cg * 8, g);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float o[8];
      if (dskip != nullptr) {
        Vec8<T>::load(dskip + vin[k] * dsp + m.cg * 8, o);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += (arg[j] == k) ? g[j] : 0.f;
      Vec8<T>::store(dx + vin[k] * dxp + m.cg * 8, o);
    }
  }
}

// ------------------------------------------------------------------------------------------
// trilinear upsample, align_corners=True (ATen upsample_trilinear3d semantics:
// src = dst * (in-1)/(out-1) in fp32, i0 = (int)src, i1 = i0 + (i0 < in-1), lambda = src - i0)
// ------------------------------------------------------------------------------------------
RSB_DEVICE int zchunk_count(int Di, int zchunk) { return (Di + zchunk - 1) / zchunk; }

struct Lerp {
  int i0, i1;
  float w0, w1;
};
RSB_DEVICE Lerp lerp_src(int o, float scale, int in_size) {
  Lerp l;
  const float src = scale * static_cast<float>(o);
  l.i0 = static_cast<int>(src);
  if (l.i0 > in_size - 1) l.i0 = in_size - 1;
  l.i1 = l.i0 + (l.i0 < in_size - 1 ? 1 : 0);
  l.w1 = src - static_cast<float>(l.i0);
  l.w0 = 1.f - l.w1;
  return l;
}
static inline float ac_scale(int in_size, int out_size) {
  return out_size > 1 ? static_cast<float>(in_size - 1) / static_cast<float>(out_size - 1) : 0.f;
}

// z-walking organisation: a thread owns one output column (yo, xo, 8 channels) and walks zo.  The (y, x) bilinear
// value of an input plane, P(zi), stays in registers for as long as consecutive output planes interpolate between the
// same two input planes, so an output costs ~2 vector loads instead of 8 (x2 upsampling advances zi every other zo)
// and no per-element index arithmetic.  The nesting t*(h*(w..)) is ATen's upsample_trilinear3d CUDA formula.
// (First version: one thread per output element, 8 loads + 3 divisions each — 625 us vs an 82 us HBM bound; the
// row-wise version that followed: 490 us.)
template <typename T>
__global__ void upsample_fwd_kernel(const T* __restrict__ x, long long xp, T* __restrict__ y, long long yp,
                                    float* __restrict__ stats, int Di, int Hi, int Wi, int Do, int Ho,
                                    int Wo, int C, float sd, float sh, float sw, int zchunks) {
  extern __shared__ float sm_acc[];  // [2*C] statistics scratch
  const int CG = C / 8;
  // z is walked in `zchunks` independent pieces: the walk is a chain of dependent loads (ncu, one piece: 36.7 % of HBM with
  // every warp waiting on long_scoreboard at 32 % occupancy); more, shorter walks put more loads in flight
  const int n = blockIdx.z / zchunks, zc = blockIdx.z % zchunks, yo = blockIdx.y;
  const int zlen = (Do + zchunks - 1) / zchunks;
  const int zo_begin = zc * zlen, zo_end = min(Do, zo_begin + zlen);
  const int cg = threadIdx.x % CG;  // blockDim.x is a multiple of CG
  const int xo = blockIdx.x * (blockDim.x / CG) + threadIdx.x / CG;
  float s1[8] = {0}, s2[8] = {0};
  if (xo < Wo) {
    const Lerp ly = lerp_src(yo, sh, Hi), lx = lerp_src(xo, sw, Wi);
    const long long o00 = (static_cast<long long>(ly.i0) * Wi + lx.i0) * xp, o01 = (static_cast<long long>(ly.i0) * Wi + lx.i1) * xp;
    const long long o10 = (static_cast<long long>(ly.i1) * Wi + lx.i0) * xp, o11 = (static_cast<long long>(ly.i1) * Wi + lx.i1) * xp;
    const long long plane = static_cast<long long>(Hi) * Wi * xp;
    const T* xb = x + static_cast<long long>(n) * Di * plane + cg * 8;
    auto bilinear = [&](int zi, float (&P)[8]) {
      const T* pz = xb + zi * plane;
      float a[8], b[8], c[8], d[8];
      Vec8<T>::load(pz + o00, a);
      Vec8<T>::load(pz + o01, b);
      Vec8<T>::load(pz + o10, c);
      Vec8<T>::load(pz + o11, d);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        P[j] = ly.w0 * (lx.w0 * a[j] + lx.w1 * b[j]) + ly.w1 * (lx.w0 * c[j] + lx.w1 * d[j]);
    };
    float P0[8], P1[8];
    int z0 = -1, z1 = -1;  // input planes held in P0 / P1
    T* yc = y + ((static_cast<long long>(n) * Do * Ho + yo) * Wo + xo) * yp + cg * 8;
    const long long ystep = static_cast<long long>(Ho) * Wo * yp;
    for (int zo = zo_begin; zo < zo_end; ++zo) {
      const Lerp lz = lerp_src(zo, sd, Di);
      if (lz.i0 != z0) {
        if (lz.i0 == z1) {
#pragma unroll
          for (int j = 0; j < 8; ++j) P0[j] = P1[j];
        } else {
          bilinear(lz.i0, P0);
        }
        z0 = lz.i0;
      }
      if (lz.i1 != z1) {
        if (lz.i1 == z0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) P1[j] = P0[j];
        } else {
          bilinear(lz.i1, P1);
        }
        z1 = lz.i1;
      }
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o[j] = lz.w0 * P0[j] + lz.w1 * P1[j];
        s1[j] += o[j];
        s2[j] = fmaf(o[j], o[j], s2[j]);
      }
      Vec8<T>::store(yc + zo * ystep, o);
    }
  }
  if (stats != nullptr) block_stats_flush(sm_acc, s1, s2, cg, C, stats + static_cast<long long>(n) * yp * 2);
}

// Adjoint as a deterministic gather: input index i receives from the outputs o whose source coordinate lies in
// (i-1, i+1).  They are consecutive, so the taps are a dense window o0 .. o0+5 with zero weights outside the support
// (at most ceil(2 / scale) <= 6 outputs for ratios up to 2.5); dense + compile-time indexing keeps everything in
// registers (a compacted {index, weight} list needs dynamic indexing => local memory: that version ran at 710 us).
struct Win {
  int o0;
  float w[6];
};
RSB_DEVICE float adjoint_weight(int o, int i, float scale, int in_size) {
  const Lerp l = lerp_src(o, scale, in_size);
  return (l.i0 == i ? l.w0 : 0.f) + (l.i1 == i ? l.w1 : 0.f);
}
RSB_DEVICE Win adjoint_window(int i, float scale, int in_size, int out_size) {
  Win t;
  int lo = 0;
  if (scale > 0.f) {
    lo = static_cast<int>(floorf((static_cast<float>(i) - 1.f) / scale)) - 1;
    if (lo < 0) lo = 0;
    for (int k = 0; k < 4 && lo < out_size - 1 && adjoint_weight(lo, i, scale, in_size) == 0.f; ++k) ++lo;
  }
  t.o0 = lo;
#pragma unroll
  for (int k = 0; k < 6; ++k) t.w[k] = (lo + k < out_size) ? adjoint_weight(lo + k, i, scale, in_size) : 0.f;
  return t;
}

// Adjoint, z-walking: a thread owns one input column (yi, xi, 8 channels) of a z-chunk [za, zb) of input planes and walks
// the output planes that touch it.  Per output plane it gathers the (y, x) adjoint Q(zo) (independent vector loads) and
// adds lz.w0 * Q / lz.w1 * Q to the two open input-plane accumulators; a finished input plane is stored once.
// Deterministic (no atomics).
// A block covers an (xt x yt) patch of input columns: the output rows / columns its threads gather from overlap, so
// the patch shape sets the L2 -> L1 traffic (19 x 19 output voxels per 8 x 8 inputs: 1.4x the output tensor; a single
// input row per block reads it 2.7x and was L2-bandwidth bound).
template <typename T>
__global__ void __launch_bounds__(512)
upsample_bwd_kernel(const T* __restrict__ dy, long long dyp, T* __restrict__ dx, long long dxp,
                    int Di, int Hi, int Wi, int Do, int Ho, int Wo, int C, float sd,
                    float sh, float sw, int zchunk, int xt) {
  const int CG = C / 8;
  const int nzc = zchunk_count(Di, zchunk);
  const int n = blockIdx.z / nzc;
  const int zc = blockIdx.z % nzc;
  const int cg = threadIdx.x % CG;
  const int col = threadIdx.x / CG;
  const int xi = blockIdx.x * xt + col % xt;
  const int yi = blockIdx.y * (blockDim.x / (CG * xt)) + col / xt;
  if (xi >= Wi || yi >= Hi) return;
  const int za = zc * zchunk, zb = min(Di, za + zchunk);
  const Win ty = adjoint_window(yi, sh, Hi, Ho), tx = adjoint_window(xi, sw, Wi, Wo);
  const long long oplane = static_cast<long long>(Ho) * Wo * dyp;
  const T* db = dy + static_cast<long long>(n) * Do * oplane + (static_cast<long long>(ty.o0) * Wo + tx.o0) * dyp + cg * 8;
  const long long orow = static_cast<long long>(Wo) * dyp;
  T* xc = dx + ((static_cast<long long>(n) * Di * Hi + yi) * Wi + xi) * dxp + cg * 8;
  const long long xstep = static_cast<long long>(Hi) * Wi * dxp;
  // first output plane whose upper source plane reaches za
  int zo = 0;
  if (za > 0 && sd > 0.f) {
    zo = static_cast<int>(floorf((static_cast<float>(za) - 1.f) / sd)) - 1;
    if (zo < 0) zo = 0;
  }
  float accA[8] = {0}, accB[8] = {0};  // input planes cur, cur + 1
  int cur = za;
  for (; zo < Do; ++zo) {
    const Lerp lz = lerp_src(zo, sd, Di);
    if (lz.i1 < za) continue;
    if (lz.i0 >= zb) break;
    while (cur < lz.i0) {  // plane cur is complete
      Vec8<T>::store(xc + cur * xstep, accA);
#pragma unroll
      for (int j = 0; j < 8; ++j) { accA[j] = accB[j]; accB[j] = 0.f; }
      ++cur;
    }
    float q[8] = {0};
    const T* pz = db + zo * oplane;
    // rows one at a time (weight recomputed, not indexed: a fully unrolled 6 x 6 gather hoists 36 loads and spills)
#pragma unroll 1
    for (int b = 0; b < 6; ++b) {
      const float wyb = (ty.o0 + b < Ho) ? adjoint_weight(ty.o0 + b, yi, sh, Hi) : 0.f;
      if (wyb != 0.f) {
        const T* prow = pz + b * orow;
        float r[8] = {0};
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          if (tx.w[c] != 0.f) {
            float f[8];
            Vec8<T>::load(prow + c * dyp, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = fmaf(tx.w[c], f[j], r[j]);
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) q[j] = fmaf(wyb, r[j], q[j]);
      }
    }
    // lz.i0 == cur or (lz.i0 < za: only its upper plane belongs to this chunk)
    const float wa = (lz.i0 == cur ? lz.w0 : 0.f) + (lz.i1 == cur ? lz.w1 : 0.f);
    const float wb = (lz.i1 == cur + 1) ? lz.w1 : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      accA[j] = fmaf(wa, q[j], accA[j]);
      accB[j] = fmaf(wb, q[j], accB[j]);
    }
  }
  for (; cur < zb; ++cur) {
    Vec8<T>::store(xc + cur * xstep, accA);
#pragma unroll
    for (int j = 0; j < 8; ++j) { accA[j] = accB[j]; accB[j] = 0.f; }
  }
}

// Separable adjoint: one generic pass per axis.  The tensor is viewed as [outer][L (walk axis)][inner][C]; a thread
// owns one (outer, inner, 8 channels) line of a chunk [za, zb) of the SHORT axis and walks the long axis with ONE
// vector load per element, keeping the two open short-axis accumulators in registers (lerp_src is monotone, so at most
// two input positions are open at any time).  z, then y, then x: 1 load + 2 FMAs per element and per pass instead of a
// 5 x 5 x 5 gather; the intermediates shrink by the upsampling ratio after every pass.
template <typename T>
__global__ void upsample_bwd_axis_kernel(const T* __restrict__ src, long long sp, T* __restrict__ dst, long long dp,
                                         int Li, int Lo, long long inner, long long lines, int C, float scale, int chunk) {
  const int CG = C / 8;
  const int nch = zchunk_count(Li, chunk);
  const int zc = blockIdx.y;
  const int za = zc * chunk, zb = min(Li, za + chunk);
  (void)nch;
  ClMap m = cl_map(CG);
  int w_first = 0;
  if (za > 0 && scale > 0.f) {
    w_first = static_cast<int>(floorf((static_cast<float>(za) - 1.f) / scale)) - 1;
    if (w_first < 0) w_first = 0;
  }
  for (long long ln = m.v0; ln < lines; ln += m.vstride) {
    const long long outer = ln / inner, in = ln - outer * inner;
    const T* ps = src + ((outer * Lo) * inner + in) * sp + m.cg * 8;
    T* pd = dst + ((outer * Li) * inner + in) * dp + m.cg * 8;
    const long long sstep = inner * sp, dstep = inner * dp;
    float accA[8] = {0}, accB[8] = {0};
    int cur = za;
    for (int w = w_first; w < Lo; ++w) {
      const Lerp l = lerp_src(w, scale, Li);
      if (l.i1 < za) continue;
      if (l.i0 >= zb) break;
      while (cur < l.i0) {
        Vec8<T>::store(pd + cur * dstep, accA);
#pragma unroll
        for (int j = 0; j < 8; ++j) { accA[j] = accB[j]; accB[j] = 0.f; }
        ++cur;
      }
      float q[8];
      Vec8<T>::load(ps + w * sstep, q);
      const float wa = (l.i0 == cur ? l.w0 : 0.f) + (l.i1 == cur ? l.w1 : 0.f);
      const float wb = (l.i1 == cur + 1) ? l.w1 : 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        accA[j] = fmaf(wa, q[j], accA[j]);
        accB[j] = fmaf(wb, q[j], accB[j]);
      }
    }
    for (; cur < zb; ++cur) {
      Vec8<T>::store(pd + cur * dstep, accA);
#pragma unroll
      for (int j = 0; j < 8; ++j) { accA[j] = accB[j]; accB[j] = 0.f; }
    }
  }
}

// ------------------------------------------------------------------------------------------
// InstanceNorm backward apply: dx = rstd * (g - S1/V - xhat * S2/V) + add
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void instnorm_bwd_apply_kernel(const T* __restrict__ g, long long gp, const T* __restrict__ x,
                                          long long xp, const float* __restrict__ x_stats,
                                          const float* __restrict__ bwd_sums, const T* __restrict__ add,
                                          long long ap, T* __restrict__ dx, long long dxp, float eps, int C,
                                          long long V) {
  const int CG = C / 8;
  const int n = blockIdx.y;
  ClMap m = cl_map(CG);
  const float inv = 1.f / static_cast<float>(V);
  float mean[8], rstd[8], m1[8], m2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const long long sidx = (static_cast<long long>(n) * xp + m.cg * 8 + j) * 2;
    stats_to_mean_rstd(x_stats[sidx], x_stats[sidx + 1], inv, eps, mean[j], rstd[j]);
    m1[j] = bwd_sums[sidx] * inv;
    m2[j] = bwd_sums[sidx + 1] * inv;
  }
  for (long long vl = m.v0; vl < V; vl += m.vstride) {
    const long long v = static_cast<long long>(n) * V + vl;
    float gv[8], xv[8], o[8];
    Vec8<T>::load(g + v * gp + m.cg * 8, gv);
    Vec8<T>::load(x + v * xp + m.cg * 8, xv);
    if (add != nullptr) {
      Vec8<T>::load(add + v * ap + m.cg * 8, o);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (xv[j] - mean[j]) * rstd[j];
      o[j] += rstd[j] * (gv[j] - m1[j] - xh * m2[j]);
    }
    Vec8<T>::store(dx + v * dxp + m.cg * 8, o);
  }
}

// ------------------------------------------------------------------------------------------
// norm_act: the conv operand  a = bf16(act((x - mean) * rstd))  (pre-activation InstanceNorm3d(eps=1e-4,
// affine=False) + (Leaky)ReLU of ConvNormAct, conv_layers.py:39-49), written once and then read by the TMA
// units of the fprop and wgrad kernels.  With lo != nullptr the fp32 value is split into hi + lo bf16
// parts (lo = bf16(a - hi)) for the 3-pass split-precision mode.  stats == nullptr => plain cast / split.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void norm_act_kernel(const T* __restrict__ x, long long xp, const float* __restrict__ stats,
                                __nv_bfloat16* __restrict__ hi, long long hp, __nv_bfloat16* __restrict__ lo,
                                long long lp, __nv_bfloat16* __restrict__ lo2, long long l2p, T* __restrict__ full, long long fp,
                                float eps, float slope, int C, long long V) {
  const int CG = C / 8;
  const int n = blockIdx.y;
  ClMap m = cl_map(CG);
  const float inv = 1.f / static_cast<float>(V);
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = 1.f;
    sh[j] = 0.f;
    if (stats != nullptr) {
      const long long sidx = (static_cast<long long>(n) * xp + m.cg * 8 + j) * 2;
      float mean, rstd;
      stats_to_mean_rstd(stats[sidx], stats[sidx + 1], inv, eps, mean, rstd);
      sc[j] = rstd;
      sh[j] = -mean * rstd;
    }
  }
  for (long long vl = m.v0; vl < V; vl += m.vstride) {
    const long long v = static_cast<long long>(n) * V + vl;
    float f[8], h[8];
    Vec8<T>::load(x + v * xp + m.cg * 8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = fmaf(f[j], sc[j], sh[j]);
      if (stats != nullptr) t = t > 0.f ? t : t * slope;
      f[j] = t;
      h[j] = __bfloat162float(__float2bfloat16_rn(t));
    }
    if (full != nullptr) Vec8<T>::store(full + v * fp + m.cg * 8, f);   // the activation itself in the storage dtype (post-activation blocks)
    if (hi != nullptr) Vec8<__nv_bfloat16>::store(hi + v * hp + m.cg * 8, f);
    if (lo != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] -= h[j];
      Vec8<__nv_bfloat16>::store(lo + v * lp + m.cg * 8, f);
      if (lo2 != nullptr) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] -= __bfloat162float(__float2bfloat16_rn(f[j]));
        Vec8<__nv_bfloat16>::store(lo2 + v * l2p + m.cg * 8, f);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Backward of a POST-activation a = act(instnorm(y)) (SingleConv = ConvNormAct(preact=False), conv_layers.py:50-68)
// when the gradient d(a) does not come out of a dgrad epilogue (pooling / upsampling / head):
// g = d * act'(yhat), S1 += sum g, S2 += sum g * yhat per (n, c); rsb_instnorm_backward_apply then turns g into d(y).
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void act_backward_stats_kernel(const T* __restrict__ d, long long dp, const T* __restrict__ y, long long yp,
                                          const float* __restrict__ y_stats, float* __restrict__ sums, T* __restrict__ g,
                                          long long gp, float eps, float slope, int C, long long V) {
  extern __shared__ float sm_acc[];
  const int CG = C / 8;
  const int n = blockIdx.y;
  ClMap m = cl_map(CG);
  const float inv = 1.f / static_cast<float>(V);
  float mean[8], rstd[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const long long sidx = (static_cast<long long>(n) * yp + m.cg * 8 + j) * 2;
    stats_to_mean_rstd(y_stats[sidx], y_stats[sidx + 1], inv, eps, mean[j], rstd[j]);
  }
  float s1[8] = {0}, s2[8] = {0};
  for (long long vl = m.v0; vl < V; vl += m.vstride) {
    const long long v = static_cast<long long>(n) * V + vl;
    float dv[8], yv[8];
    Vec8<T>::load(d + v * dp + m.cg * 8, dv);
    Vec8<T>::load(y + v * yp + m.cg * 8, yv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float h = (yv[j] - mean[j]) * rstd[j];
      dv[j] = h > 0.f ? dv[j] : dv[j] * slope;
      s1[j] += dv[j];
      s2[j] = fmaf(dv[j], h, s2[j]);
    }
    Vec8<T>::store(g + v * gp + m.cg * 8, dv);
  }
  block_stats_flush(sm_acc, s1, s2, m.cg, C, sums + static_cast<long long>(n) * yp * 2);
}

}  // namespace rsb

using namespace rsb;

#define RSB_CL_COMMON(C_, N_)                                                            \
  RSB_REQUIRE((C_) > 0 && (C_) % 8 == 0, "channel count must be a positive multiple of 8 (got %d)", (C_)); \
  const int CG = (C_) / 8;                                                               \
  const int block = cl_block(CG);                                                        \
  RSB_REQUIRE(block > 0, "too many channels (%d)", (C_));                                \
  const int sms = rsb_num_sms();                                                         \
  RSB_REQUIRE(sms > 0, "no CUDA device");                                                \
  RSB_REQUIRE((N_) > 0 && (N_) <= 65535, "bad batch size %d", (N_));                     \
  cudaStream_t st = static_cast<cudaStream_t>(stream);

#define RSB_BY_DTYPE(dtype, CALL_BF16, CALL_F32)           \
  if ((dtype) == RSB_BF16) { CALL_BF16; }                  \
  else if ((dtype) == RSB_F32) { CALL_F32; }               \
  else { set_last_error("bad dtype %d", (dtype)); return -1; }

extern "C" int rsb_ncdhw_to_ndhwc(const float* src, void* dst, int dst_pitch, int dtype, int N, int C,
                                  int D, int H, int W, void* stream) {
  RSB_REQUIRE(src && dst, "ncdhw_to_ndhwc: null pointer");
  RSB_CL_COMMON(C, N)
  const long long V = static_cast<long long>(D) * H * W;
  dim3 grid(cl_grid(V * CG, block, sms), N);
  RSB_BY_DTYPE(dtype,
               (ncdhw_to_ndhwc_kernel<__nv_bfloat16><<<grid, block, 0, st>>>(src, (__nv_bfloat16*)dst, dst_pitch, C, V)),
               (ncdhw_to_ndhwc_kernel<float><<<grid, block, 0, st>>>(src, (float*)dst, dst_pitch, C, V)))
  return check_launch("ncdhw_to_ndhwc");
}

extern "C" int rsb_ndhwc_to_ncdhw(const void* src, int src_pitch, int dtype, float* dst, int N, int C,
                                  int D, int H, int W, void* stream) {
  RSB_REQUIRE(src && dst, "ndhwc_to_ncdhw: null pointer");
  RSB_CL_COMMON(C, N)
  const long long V = static_cast<long long>(D) * H * W;
  dim3 grid(cl_grid(V * CG, block, sms), N);
  RSB_BY_DTYPE(dtype,
               (ndhwc_to_ncdhw_kernel<__nv_bfloat16><<<grid, block, 0, st>>>((const __nv_bfloat16*)src, src_pitch, dst, C, V)),
               (ndhwc_to_ncdhw_kernel<float><<<grid, block, 0, st>>>((const float*)src, src_pitch, dst, C, V)))
  return check_launch("ndhwc_to_ncdhw");
}

extern "C" int rsb_channel_stats(const void* x, int x_pitch, int dtype, float* stats, int N, int D, int H,
                                 int W, int C, void* stream) {
  RSB_REQUIRE(x && stats, "channel_stats: null pointer");
  RSB_CL_COMMON(C, N)
  const long long V = static_cast<long long>(D) * H * W;
  dim3 grid(cl_grid(V * CG, block, sms, 4), N);
  const size_t sm = sizeof(float) * 2 * C;
  RSB_BY_DTYPE(dtype,
               (channel_stats_kernel<__nv_bfloat16><<<grid, block, sm, st>>>((const __nv_bfloat16*)x, x_pitch, stats, C, V)),
               (channel_stats_kernel<float><<<grid, block, sm, st>>>((const float*)x, x_pitch, stats, C, V)))
  return check_launch("channel_stats");
}

extern "C" int rsb_maxpool2_forward(const void* x, int x_pitch, void* y, int y_pitch, int dtype,
                                    float* out_stats, int N, int D, int H, int W, int C, void* stream) {
  RSB_REQUIRE(x && y, "maxpool2_forward: null pointer");
  RSB_REQUIRE(D % 2 == 0 && H % 2 == 0 && W % 2 == 0 && D > 0, "maxpool2: spatial dims must be even (%d,%d,%d)", D, H, W);
  RSB_CL_COMMON(C, N)
  const long long Vo = static_cast<long long>(D / 2) * (H / 2) * (W / 2);
  RSB_REQUIRE(Vo < (1LL << 31), "maxpool2: volume too large");
  dim3 grid(cl_grid(Vo * CG, block, sms, 8), N);
  const size_t sm = sizeof(float) * 2 * C;
  RSB_BY_DTYPE(dtype,
               (maxpool2_fwd_kernel<__nv_bfloat16><<<grid, block, sm, st>>>((const __nv_bfloat16*)x, x_pitch, (__nv_bfloat16*)y, y_pitch, out_stats, D, H, W, C)),
               (maxpool2_fwd_kernel<float><<<grid, block, sm, st>>>((const float*)x, x_pitch, (float*)y, y_pitch, out_stats, D, H, W, C)))
  return check_launch("maxpool2_forward");
}

extern "C" int rsb_maxpool2_backward(const void* x, int x_pitch, const void* dy, int dy_pitch,
                                     const void* dskip, int dskip_pitch, void* dx, int dx_pitch, int dtype,
                                     int N, int D, int H, int W, int C, void* stream) {
  RSB_REQUIRE(x && dy && dx, "maxpool2_backward: null pointer");
  RSB_REQUIRE(D % 2 == 0 && H % 2 == 0 && W % 2 == 0 && D > 0, "maxpool2: spatial dims must be even (%d,%d,%d)", D, H, W);
  RSB_CL_COMMON(C, N)
  const long long Vo = static_cast<long long>(D / 2) * (H / 2) * (W / 2);
  RSB_REQUIRE(Vo < (1LL << 31), "maxpool2: volume too large");
  dim3 grid(cl_grid(Vo * CG, block, sms), N);
  RSB_BY_DTYPE(dtype,
               (maxpool2_bwd_kernel<__nv_bfloat16><<<grid, block, 0, st>>>((const __nv_bfloat16*)x, x_pitch, (const __nv_bfloat16*)dy, dy_pitch, (const __nv_bfloat16*)dskip, dskip_pitch, (__nv_bfloat16*)dx, dx_pitch, D, H, W, C)),
               (maxpool2_bwd_kernel<float><<<grid, block, 0, st>>>((const float*)x, x_pitch, (const float*)dy, dy_pitch, (const float*)dskip, dskip_pitch, (float*)dx, dx_pitch, D, H, W, C)))
  return check_launch("maxpool2_backward");
}

extern "C" int rsb_depth_to_space2(const void* q, int q_pitch, const float* bias, void* up, int up_pitch, int dtype, float* out_stats,
                                   int N, int D, int H, int W, int C, void* stream) {
  RSB_REQUIRE(q && up, "depth_to_space2: null pointer");
  RSB_REQUIRE(D > 0 && H > 0 && W > 0 && q_pitch >= 8 * C, "depth_to_space2: bad geometry (q needs 8 * C channels)");
  RSB_CL_COMMON(C, N)
  const long long Vi = static_cast<long long>(D) * H * W;
  RSB_REQUIRE(Vi < (1LL << 31), "depth_to_space2: volume too large");
  dim3 grid(cl_grid(Vi * CG, block, sms, 8), N);
  const size_t sm = sizeof(float) * 2 * C;
  RSB_BY_DTYPE(dtype,
               (depth_to_space2_kernel<__nv_bfloat16><<<grid, block, sm, st>>>((const __nv_bfloat16*)q, q_pitch, bias, (__nv_bfloat16*)up, up_pitch, out_stats, D, H, W, C)),
               (depth_to_space2_kernel<float><<<grid, block, sm, st>>>((const float*)q, q_pitch, bias, (float*)up, up_pitch, out_stats, D, H, W, C)))
  return check_launch("depth_to_space2");
}

extern "C" int rsb_space_to_depth2(const void* up, int up_pitch, void* q, int q_pitch, int dtype, int N, int D, int H, int W, int C,
                                   void* stream) {
  RSB_REQUIRE(q && up, "space_to_depth2: null pointer");
  RSB_REQUIRE(D > 0 && H > 0 && W > 0 && q_pitch >= 8 * C, "space_to_depth2: bad geometry (q needs 8 * C channels)");
  RSB_CL_COMMON(C, N)
  const long long Vi = static_cast<long long>(D) * H * W;
  RSB_REQUIRE(Vi < (1LL << 31), "space_to_depth2: volume too large");
  dim3 grid(cl_grid(Vi * CG, block, sms), N);
  RSB_BY_DTYPE(dtype,
               (space_to_depth2_kernel<__nv_bfloat16><<<grid, block, 0, st>>>((const __nv_bfloat16*)up, up_pitch, (__nv_bfloat16*)q, q_pitch, D, H, W, C)),
               (space_to_depth2_kernel<float><<<grid, block, 0, st>>>((const float*)up, up_pitch, (float*)q, q_pitch, D, H, W, C)))
  return check_launch("space_to_depth2");
}

extern "C" int rsb_upsample_trilinear_forward(const void* x, int x_pitch, void* y, int y_pitch, int dtype,
                                              float* out_stats, int N, int Di, int Hi, int Wi, int Do, int Ho,
                                              int Wo, int C, void* stream) {
  RSB_REQUIRE(x && y, "upsample_forward: null pointer");
  RSB_REQUIRE(Di > 0 && Hi > 0 && Wi > 0 && Do > 0 && Ho > 0 && Wo > 0, "upsample: bad geometry");
  RSB_CL_COMMON(C, N)
  RSB_REQUIRE(Ho <= 65535, "upsample: output height too large");
  (void)sms;
  const int xt = block / CG;
  int zchunks = 1;
  {
    // enough blocks for ~8 per SM, pieces of at least 8 output planes (a piece re-computes two bilinear input planes)
    const long long blocks = static_cast<long long>((Wo + xt - 1) / xt) * Ho * N;
    const int sms_now = rsb_num_sms() > 0 ? rsb_num_sms() : 148;
    while (zchunks < 8 && blocks * zchunks < 8LL * sms_now && Do / (zchunks * 2) >= 8) zchunks *= 2;
    if (static_cast<long long>(N) * zchunks > 65535) zchunks = 1;
  }
  dim3 grid(static_cast<unsigned>((Wo + xt - 1) / xt), Ho, N * zchunks);
  const size_t sm = sizeof(float) * 2 * C;
  const float sd = ac_scale(Di, Do), sh = ac_scale(Hi, Ho), sw = ac_scale(Wi, Wo);
  RSB_BY_DTYPE(dtype,
               (upsample_fwd_kernel<__nv_bfloat16><<<grid, block, sm, st>>>((const __nv_bfloat16*)x, x_pitch, (__nv_bfloat16*)y, y_pitch, out_stats, Di, Hi, Wi, Do, Ho, Wo, C, sd, sh, sw, zchunks)),
               (upsample_fwd_kernel<float><<<grid, block, sm, st>>>((const float*)x, x_pitch, (float*)y, y_pitch, out_stats, Di, Hi, Wi, Do, Ho, Wo, C, sd, sh, sw, zchunks)))
  return check_launch("upsample_trilinear_forward");
}

extern "C" int rsb_upsample_trilinear_backward(const void* dy, int dy_pitch, void* dx, int dx_pitch, int dtype,
                                               int N, int Di, int Hi, int Wi, int Do, int Ho, int Wo, int C,
                                               void* workspace, void* stream) {
  RSB_REQUIRE(dy && dx, "upsample_backward: null pointer");
  RSB_REQUIRE(Di > 0 && Hi > 0 && Wi > 0 && Do > 0 && Ho > 0 && Wo > 0, "upsample: bad geometry");
  RSB_CL_COMMON(C, N)
  RSB_REQUIRE(Do >= Di && Ho >= Hi && Wo >= Wi, "upsample_backward: output must not be smaller than the input");
  RSB_REQUIRE(2 * Ho <= 5 * Hi + 1 && 2 * Wo <= 5 * Wi + 1, "upsample_backward: y / x ratios above 2.5 are not supported");
  if (workspace != nullptr) {
    // separable adjoint: z pass into tmp1 [N][Di][Ho][Wo][C], y pass into tmp2 [N][Di][Hi][Wo][C], x pass into dx
    char* ws = static_cast<char*>(workspace);
    const size_t esz = dtype == RSB_BF16 ? 2 : 4;
    void* tmp1 = ws;
    void* tmp2 = ws + static_cast<size_t>(N) * Di * Ho * Wo * C * esz;
    struct Pass { const void* src; long long sp; void* dst; long long dp; int Li, Lo; long long inner, outer; };
    const Pass passes[3] = {
        {dy, dy_pitch, tmp1, C, Di, Do, static_cast<long long>(Ho) * Wo, N},
        {tmp1, C, tmp2, C, Hi, Ho, Wo, static_cast<long long>(N) * Di},
        {tmp2, C, dx, dx_pitch, Wi, Wo, 1, static_cast<long long>(N) * Di * Hi}};
    for (const Pass& ps : passes) {
      const long long lines = ps.outer * ps.inner;
      const float scale = ac_scale(ps.Li, ps.Lo);
      int nc = 1;
      while (nc < 16 && (lines * CG / block) * nc < sms * 8LL && (ps.Li + nc * 2 - 1) / (nc * 2) >= 4) nc *= 2;
      const int chunk = (ps.Li + nc - 1) / nc;
      const int nchunks = (ps.Li + chunk - 1) / chunk;
      dim3 g(cl_grid(lines * CG, block, sms, 16), nchunks);
      RSB_BY_DTYPE(dtype,
                   (upsample_bwd_axis_kernel<__nv_bfloat16><<<g, block, 0, st>>>((const __nv_bfloat16*)ps.src, ps.sp, (__nv_bfloat16*)ps.dst, ps.dp, ps.Li, ps.Lo, ps.inner, lines, C, scale, chunk)),
                   (upsample_bwd_axis_kernel<float><<<g, block, 0, st>>>((const float*)ps.src, ps.sp, (float*)ps.dst, ps.dp, ps.Li, ps.Lo, ps.inner, lines, C, scale, chunk)))
      const int rc1 = check_launch("upsample_bwd_axis_kernel");
      if (rc1) return rc1;
    }
    return 0;
  }
  // patch of input columns per block: xt x yt with CG * xt * yt <= 512 threads
  int xt = 8;
  while (xt > 1 && CG * xt > 512) xt >>= 1;
  RSB_REQUIRE(CG * xt <= 512, "upsample_backward: too many channels (%d)", C);
  int yt = 512 / (CG * xt);
  if (yt > 8) yt = 8;
  if (yt > Hi) yt = Hi;
  const int threads = CG * xt * yt;
  const int gx = (Wi + xt - 1) / xt, gy = (Hi + yt - 1) / yt;
  // z chunks: enough threads to fill the machine (each thread walks ~Do / nz output planes)
  const long long blocks = static_cast<long long>(N) * gx * gy;
  int nz = 1;
  while (nz < 16 && blocks * nz < sms * 3LL && (Di + nz * 2 - 1) / (nz * 2) >= 2) nz *= 2;
  const int zchunk = (Di + nz - 1) / nz;
  const int nzc = (Di + zchunk - 1) / zchunk;
  RSB_REQUIRE(static_cast<long long>(N) * nzc <= 65535 && gy <= 65535, "upsample_backward: grid too large");
  dim3 grid(gx, gy, N * nzc);
  const float sd = ac_scale(Di, Do), sh = ac_scale(Hi, Ho), sw = ac_scale(Wi, Wo);
  RSB_BY_DTYPE(dtype,
               (upsample_bwd_kernel<__nv_bfloat16><<<grid, threads, 0, st>>>((const __nv_bfloat16*)dy, dy_pitch, (__nv_bfloat16*)dx, dx_pitch, Di, Hi, Wi, Do, Ho, Wo, C, sd, sh, sw, zchunk, xt)),
               (upsample_bwd_kernel<float><<<grid, threads, 0, st>>>((const float*)dy, dy_pitch, (float*)dx, dx_pitch, Di, Hi, Wi, Do, Ho, Wo, C, sd, sh, sw, zchunk, xt)))
  return check_launch("upsample_trilinear_backward");
}

extern "C" int rsb_instnorm_backward_apply(const void* g, int g_pitch, const void* x, int x_pitch,
                                           const float* x_stats, const float* bwd_sums, const void* add,
                                           int add_pitch, void* dx, int dx_pitch, int dtype, float eps, int N,
                                           int D, int H, int W, int C, void* stream) {
  RSB_REQUIRE(g && x && x_stats && bwd_sums && dx, "instnorm_backward_apply: null pointer");
  RSB_CL_COMMON(C, N)
  const long long V = static_cast<long long>(D) * H * W;
  dim3 grid(cl_grid(V * CG, block, sms), N);
  RSB_BY_DTYPE(dtype,
               (instnorm_bwd_apply_kernel<__nv_bfloat16><<<grid, block, 0, st>>>((const __nv_bfloat16*)g, g_pitch, (const __nv_bfloat16*)x, x_pitch, x_stats, bwd_sums, (const __nv_bfloat16*)add, add_pitch, (__nv_bfloat16*)dx, dx_pitch, eps, C, V)),
               (instnorm_bwd_apply_kernel<float><<<grid, block, 0, st>>>((const float*)g, g_pitch, (const float*)x, x_pitch, x_stats, bwd_sums, (const float*)add, add_pitch, (float*)dx, dx_pitch, eps, C, V)))
  return check_launch("instnorm_backward_apply");
}

extern "C" int rsb_norm_act(const void* x, int x_pitch, int dtype, const float* stats, float eps, float slope,
                            void* hi, int hi_pitch, void* lo, int lo_pitch, void* lo2, int lo2_pitch, void* full, int full_pitch,
                            int N, int D, int H, int W, int C, void* stream) {
  RSB_REQUIRE(x && (hi || full), "norm_act: null pointer");
  RSB_REQUIRE(x_pitch % 8 == 0 && (!hi || hi_pitch % 8 == 0) && (!lo || (hi && lo_pitch % 8 == 0)) && (!lo2 || (lo && lo2_pitch % 8 == 0)) &&
                  (!full || full_pitch % 8 == 0),
              "norm_act: pitches must be multiples of 8 (lo needs hi, lo2 needs lo)");
  RSB_CL_COMMON(C, N)
  const long long V = static_cast<long long>(D) * H * W;
  dim3 grid(cl_grid(V * CG, block, sms), N);
  RSB_BY_DTYPE(dtype,
               (norm_act_kernel<__nv_bfloat16><<<grid, block, 0, st>>>((const __nv_bfloat16*)x, x_pitch, stats, (__nv_bfloat16*)hi, hi_pitch, (__nv_bfloat16*)lo, lo_pitch, (__nv_bfloat16*)lo2, lo2_pitch, (__nv_bfloat16*)full, full_pitch, eps, slope, C, V)),
               (norm_act_kernel<float><<<grid, block, 0, st>>>((const float*)x, x_pitch, stats, (__nv_bfloat16*)hi, hi_pitch, (__nv_bfloat16*)lo, lo_pitch, (__nv_bfloat16*)lo2, lo2_pitch, (float*)full, full_pitch, eps, slope, C, V)))
  return check_launch("norm_act");
}

extern "C" int rsb_act_backward_stats(const void* d, int d_pitch, const void* y, int y_pitch, const float* y_stats,
                                      float* bwd_sums, void* g, int g_pitch, int dtype, float eps, float slope, int N, int D,
                                      int H, int W, int C, void* stream) {
  RSB_REQUIRE(d && y && y_stats && bwd_sums && g, "act_backward_stats: null pointer");
  RSB_CL_COMMON(C, N)
  const long long V = static_cast<long long>(D) * H * W;
  dim3 grid(cl_grid(V * CG, block, sms, 4), N);
  const size_t sm = sizeof(float) * 2 * C;
  RSB_BY_DTYPE(dtype,
               (act_backward_stats_kernel<__nv_bfloat16><<<grid, block, sm, st>>>((const __nv_bfloat16*)d, d_pitch, (const __nv_bfloat16*)y, y_pitch, y_stats, bwd_sums, (__nv_bfloat16*)g, g_pitch, eps, slope, C, V)),
               (act_backward_stats_kernel<float><<<grid, block, sm, st>>>((const float*)d, d_pitch, (const float*)y, y_pitch, y_stats, bwd_sums, (float*)g, g_pitch, eps, slope, C, V)))
  return check_launch("act_backward_stats");
}
