// stem_head.cu — the two convolutions of the UNet that do not belong on the tensor pipe:
//   * stem  inconv.conv1 = nn.Conv3d(1, base, 3, padding=1, bias=False)   unet_utils.py:15,18
//     K = 27: bandwidth-bound; direct CUDA-core conv, one voxel per thread, all Cout in registers.
//   * head  UNet.outc   = nn.Conv3d(base, num_classes, kernel_size=1)     unet.py:47,62
//     N = num_classes (2..42): a per-voxel mat-vec, reads NDHWC features once, writes NCDHW fp32
//     logits (the layout calculate_loss consumes, losses_foundation.py:859).
#include "rsb_common.cuh"

#include "../../include/rsuper_b200.h"

namespace rsb {

// butterfly over 32 lanes for 16 values (same scheme as conv3_igemm.cu's epilogue)
RSB_DEVICE float bfly16(float (&v)[16], int lane) {
  float b8[8], b4[4], b2[2];
  {
    const bool up = lane & 16;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float send = up ? v[j] : v[j + 8], keep = up ? v[j + 8] : v[j];
      b8[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool up = lane & 8;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float send = up ? b8[j] : b8[j + 4], keep = up ? b8[j + 4] : b8[j];
      b4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool up = lane & 4;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float send = up ? b4[j] : b4[j + 2], keep = up ? b4[j + 2] : b4[j];
      b2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  const bool up = lane & 2;
  float send = up ? b2[0] : b2[1], keep = up ? b2[1] : b2[0];
  float b1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  b1 += __shfl_xor_sync(0xffffffffu, b1, 1);
  return b1;
}
RSB_DEVICE int bfly_col(int lane) {
  return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}

constexpr int kStemThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kStemThreads) stem_fwd_kernel(const float* __restrict__ x,
                                                                const float* __restrict__ w, T* __restrict__ y,
                                                                long long yp, float* __restrict__ stats,
                                                                int D, int H, int W, int Cout) {
  extern __shared__ float sm[];
  float* sw = sm;                    // [27][CoutPad16]
  const int cpad = (Cout + 15) / 16 * 16;
  float* sacc = sm + 27 * cpad;      // [Cout][2]
  for (int i = threadIdx.x; i < 27 * cpad; i += blockDim.x) {
    const int tap = i / cpad, co = i % cpad;
    sw[i] = co < Cout ? w[co * 27 + tap] : 0.f;
  }
  for (int i = threadIdx.x; i < 2 * Cout; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int n = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const long long V = static_cast<long long>(D) * H * W;
  const long long vbase = static_cast<long long>(blockIdx.x) * blockDim.x;
  const long long v = vbase + threadIdx.x;
  const bool ok = v < V;
  float xin[27];
  {
    const int xq = static_cast<int>(v % W), yq = static_cast<int>((v / W) % H), zq = static_cast<int>(v / (static_cast<long long>(W) * H));
#pragma unroll
    for (int tap = 0; tap < 27; ++tap) {
      const int z = zq + tap / 9 - 1, yy = yq + (tap / 3) % 3 - 1, xx = xq + tap % 3 - 1;
      const bool inb = ok && z >= 0 && z < D && yy >= 0 && yy < H && xx >= 0 && xx < W;
      xin[tap] = inb ? x[((static_cast<long long>(n) * D + z) * H + yy) * W + xx] : 0.f;
    }
  }
  for (int c0 = 0; c0 < Cout; c0 += 16) {
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 27; ++tap) {
      const float* wr = sw + tap * cpad + c0;
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = fmaf(xin[tap], wr[j], acc[j]);
    }
    const int nvalid = Cout - c0;
    if (ok) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = acc[j];
      Vec8<T>::store(y + (static_cast<long long>(n) * V + v) * yp + c0, o);
      if (nvalid > 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = acc[8 + j];
        Vec8<T>::store(y + (static_cast<long long>(n) * V + v) * yp + c0 + 8, o);
      }
    }
    if (stats != nullptr) {
      float s1[16], s2[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        s1[j] = ok ? acc[j] : 0.f;
        s2[j] = s1[j] * s1[j];
      }
      const float t1 = bfly16(s1, lane), t2 = bfly16(s2, lane);
      const int col = c0 + bfly_col(lane);
      if ((lane & 1) == 0 && col < Cout) {
        atomicAdd(&sacc[col * 2], t1);
        atomicAdd(&sacc[col * 2 + 1], t2);
      }
    }
  }
  if (stats != nullptr) {
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * Cout; i += blockDim.x)
      atomicAdd(&stats[static_cast<long long>(n) * yp * 2 + i], sacc[i]);
  }
}

// Register-tiled variant (W % 4 == 0): a thread owns FOUR consecutive voxels of an x row and 16 output channels at a time.
// The one-voxel kernel above reads a weight from shared memory for every FMA (ncu: 72 % LSU wavefronts, 10 % of HBM, 30 % of
// the FP32 pipe); here a 128-bit broadcast load of four weights feeds 16 FMAs, the 3 x 3 x 6 input window lives in registers
// and the statistics butterfly runs once per four voxels.
constexpr int kStemVX = 4;
template <typename T>
__global__ void __launch_bounds__(kStemThreads) stem_fwd_tiled_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                      T* __restrict__ y, long long yp, float* __restrict__ stats,
                                                                      int D, int H, int W, int Cout) {
  extern __shared__ float sm[];   // dynamic shared memory is 16-byte aligned (float4 weight loads below)
  float* sw = sm;                    // [27][CoutPad16]
  const int cpad = (Cout + 15) / 16 * 16;
  float* sacc = sm + 27 * cpad;      // [Cout][2]
  for (int i = threadIdx.x; i < 27 * cpad; i += blockDim.x) {
    const int tap = i / cpad, co = i % cpad;
    sw[i] = co < Cout ? w[co * 27 + tap] : 0.f;
  }
  for (int i = threadIdx.x; i < 2 * Cout; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int n = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const long long V = static_cast<long long>(D) * H * W;
  const int WG = W / kStemVX;
  const long long groups = static_cast<long long>(D) * H * WG;
  const long long gidx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool ok = gidx < groups;
  const int xg = static_cast<int>(gidx % WG), yq = static_cast<int>((gidx / WG) % H), zq = static_cast<int>(gidx / (static_cast<long long>(WG) * H));
  const int x0 = xg * kStemVX;
  float xin[9][kStemVX + 2];         // (dz, dy) rows of the window, x0 - 1 .. x0 + 4
#pragma unroll
  for (int r = 0; r < 9; ++r) {
    const int z = zq + r / 3 - 1, yy = yq + r % 3 - 1;
    const bool rin = ok && z >= 0 && z < D && yy >= 0 && yy < H;
    const float* row = x + ((static_cast<long long>(n) * D + (rin ? z : 0)) * H + (rin ? yy : 0)) * W;
#pragma unroll
    for (int i = 0; i < kStemVX + 2; ++i) {
      const int xx = x0 - 1 + i;
      xin[r][i] = (rin && xx >= 0 && xx < W) ? __ldg(row + xx) : 0.f;
    }
  }
  const long long v0 = (static_cast<long long>(zq) * H + yq) * W + x0;
  for (int c0 = 0; c0 < Cout; c0 += 16) {
    // packed FP32 FMA (FFMA2, sm_100): the three-register scalar FFMA issues every second cycle per scheduler, the packed form
    // carries two FMAs per issue slot — this loop is FMA-issue bound (864 FMA per voxel).  Same rounding as fmaf.
    float2 acc2[kStemVX][8];
#pragma unroll
    for (int q = 0; q < kStemVX; ++q)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc2[q][j] = make_float2(0.f, 0.f);
#pragma unroll
    for (int r = 0; r < 9; ++r) {
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const float4* wr = reinterpret_cast<const float4*>(sw + (r * 3 + kw) * cpad + c0);
        const float4 w0 = wr[0], w1 = wr[1], w2 = wr[2], w3 = wr[3];
        const float2 wv[8] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y), make_float2(w1.z, w1.w),
                              make_float2(w2.x, w2.y), make_float2(w2.z, w2.w), make_float2(w3.x, w3.y), make_float2(w3.z, w3.w)};
#pragma unroll
        for (int q = 0; q < kStemVX; ++q) {
          const float2 xv = make_float2(xin[r][q + kw], xin[r][q + kw]);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc2[q][j] = __ffma2_rn(xv, wv[j], acc2[q][j]);
        }
      }
    }
    float acc[kStemVX][16];
#pragma unroll
    for (int q = 0; q < kStemVX; ++q)
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc[q][2 * j] = acc2[q][j].x; acc[q][2 * j + 1] = acc2[q][j].y; }
    const int nvalid = Cout - c0;
    if (ok) {
#pragma unroll
      for (int q = 0; q < kStemVX; ++q) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = acc[q][j];
        T* dst = y + (static_cast<long long>(n) * V + v0 + q) * yp + c0;
        Vec8<T>::store(dst, o);
        if (nvalid > 8) {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = acc[q][8 + j];
          Vec8<T>::store(dst + 8, o);
        }
      }
    }
    if (stats != nullptr) {
      float s1[16], s2[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        s1[j] = 0.f; s2[j] = 0.f;
        if (ok) {
#pragma unroll
          for (int q = 0; q < kStemVX; ++q) { s1[j] += acc[q][j]; s2[j] = fmaf(acc[q][j], acc[q][j], s2[j]); }
        }
      }
      const float t1 = bfly16(s1, lane), t2 = bfly16(s2, lane);
      const int col = c0 + bfly_col(lane);
      if ((lane & 1) == 0 && col < Cout) {
        atomicAdd(&sacc[col * 2], t1);
        atomicAdd(&sacc[col * 2 + 1], t2);
      }
    }
  }
  if (stats != nullptr) {
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * Cout; i += blockDim.x)
      atomicAdd(&stats[static_cast<long long>(n) * yp * 2 + i], sacc[i]);
  }
}

// dW[co][tap] = sum_{n,v} dy[v][co] * x[v + tap - 1]; thread = (tap, 8-channel group), block = voxel span
template <typename T>
__global__ void stem_wgrad_kernel(const float* __restrict__ x, const T* __restrict__ dy, long long dyp,
                                  float* __restrict__ dw, int N, int D, int H, int W, int Cout,
                                  int vox_per_block) {
  const int CG = Cout / 8;
  const int tap = threadIdx.x / CG, cg = threadIdx.x % CG;
  if (tap >= 27) return;
  const int dz = tap / 9 - 1, dyy = (tap / 3) % 3 - 1, dxx = tap % 3 - 1;
  const long long V = static_cast<long long>(D) * H * W;
  const long long total = V * N;
  const long long v0 = static_cast<long long>(blockIdx.x) * vox_per_block;
  long long v1 = v0 + vox_per_block;
  if (v1 > total) v1 = total;
  float acc[8] = {0};
  for (long long gv = v0; gv < v1; ++gv) {
    const long long v = gv % V;
    const int n = static_cast<int>(gv / V);
    const int xq = static_cast<int>(v % W), yq = static_cast<int>((v / W) % H), zq = static_cast<int>(v / (static_cast<long long>(W) * H));
    const int z = zq + dz, yy = yq + dyy, xx = xq + dxx;
    if (z < 0 || z >= D || yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
    const float xv = x[((static_cast<long long>(n) * D + z) * H + yy) * W + xx];
    float g[8];
    Vec8<T>::load(dy + gv * dyp + cg * 8, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaf(xv, g[j], acc[j]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) atomicAdd(&dw[(cg * 8 + j) * 27 + tap], acc[j]);
}

// Tiled variant (Cout <= 64): a block stages a 1 x 8 x 32 voxel tile of dy (as fp32) and the haloed
// 3 x 10 x 34 input tile in shared memory.  Thread (tile row g, (kd, kh), 8-channel group) walks the 32 voxels of
// its row with a sliding 3-wide input window and keeps 3 (kw) x 8 (channels) partial sums in registers across all
// its tiles: 3 shared loads per 24 FMAs (the first tiled version — one tap per thread, 3 loads per 8 FMAs — was
// bound by shared-memory issue: 650 us at 2 x 128^3 x 32).
constexpr int kSwTX = 32, kSwTY = 8;
constexpr int kSwVox = kSwTX * kSwTY;
template <typename T>
__global__ void __launch_bounds__(576) stem_wgrad_tiled_kernel(const float* __restrict__ x, const T* __restrict__ dy,
                                                               long long dyp, float* __restrict__ dw, int N, int D,
                                                               int H, int W, int Cout, int tiles_x, int tiles_y,
                                                               long long num_tiles) {
  extern __shared__ float sm[];
  float* sx = sm;                         // [3][10][34]
  float* sdy = sm + 3 * 10 * 34;          // [256][Cout]
  float* sacc = sdy + kSwVox * Cout;      // [27][Cout] block-level reduction over the 8 row groups
  const int CG = Cout / 8;
  const int g = threadIdx.x / (9 * CG), r = threadIdx.x % (9 * CG);   // blockDim.x == 8 * 9 * CG
  const int khd = r / CG, cg = r % CG;
  const int dz = khd / 3, dyy = khd % 3;
  for (int i = threadIdx.x; i < 27 * Cout; i += blockDim.x) sacc[i] = 0.f;
  float2 acc2[3][4];                       // packed FP32 FMA (FFMA2): two channels per instruction, same rounding as fmaf
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc2[k][j] = make_float2(0.f, 0.f);
  for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    long long t = tile;
    const int xt = static_cast<int>(t % tiles_x); t /= tiles_x;
    const int yt = static_cast<int>(t % tiles_y); t /= tiles_y;
    const int z = static_cast<int>(t % D);
    const int n = static_cast<int>(t / D);
    const int x0 = xt * kSwTX, y0 = yt * kSwTY;
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * 10 * 34; i += blockDim.x) {
      const int xx = i % 34, yy = (i / 34) % 10, zz = i / 340;
      const int gz = z - 1 + zz, gy = y0 - 1 + yy, gx = x0 - 1 + xx;
      const bool inb = gz >= 0 && gz < D && gy >= 0 && gy < H && gx >= 0 && gx < W;
      sx[i] = inb ? x[((static_cast<long long>(n) * D + gz) * H + gy) * W + gx] : 0.f;
    }
    for (int i = threadIdx.x; i < kSwVox * CG; i += blockDim.x) {
      const int c8 = i % CG, v = i / CG;
      const int gy = y0 + v / kSwTX, gx = x0 + v % kSwTX;
      float f[8];
      if (gy < H && gx < W) {
        Vec8<T>::load(dy + (((static_cast<long long>(n) * D + z) * H + gy) * W + gx) * dyp + c8 * 8, f);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = 0.f;
      }
      float4* d4 = reinterpret_cast<float4*>(sdy + v * Cout + c8 * 8);
      d4[0] = make_float4(f[0], f[1], f[2], f[3]);
      d4[1] = make_float4(f[4], f[5], f[6], f[7]);
    }
    __syncthreads();
    const float* xr = sx + (dz * 10 + g + dyy) * 34;
    const float* dr = sdy + (g * kSwTX) * Cout + cg * 8;
    float xa = xr[0], xb = xr[1];
#pragma unroll 4
    for (int vx = 0; vx < kSwTX; ++vx) {
      const float xc = xr[vx + 2];
      const float4* d4 = reinterpret_cast<const float4*>(dr + vx * Cout);
      const float4 a = d4[0], b = d4[1];
      const float2 d[4] = {make_float2(a.x, a.y), make_float2(a.z, a.w), make_float2(b.x, b.y), make_float2(b.z, b.w)};
      const float2 xa2 = make_float2(xa, xa), xb2 = make_float2(xb, xb), xc2 = make_float2(xc, xc);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc2[0][j] = __ffma2_rn(xa2, d[j], acc2[0][j]);
        acc2[1][j] = __ffma2_rn(xb2, d[j], acc2[1][j]);
        acc2[2][j] = __ffma2_rn(xc2, d[j], acc2[2][j]);
      }
      xa = xb;
      xb = xc;
    }
  }
  float acc[3][8];
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[k][2 * j] = acc2[k][j].x; acc[k][2 * j + 1] = acc2[k][j].y; }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&sacc[(khd * 3 + k) * Cout + cg * 8 + j], acc[k][j]);
  __syncthreads();
  for (int i = threadIdx.x; i < 27 * Cout; i += blockDim.x) {
    const int tap = i / Cout, co = i % Cout;
    atomicAdd(&dw[co * 27 + tap], sacc[i]);
  }
}

// ------------------------------------------------------------------------------------------
// head 1x1x1 + bias
// ------------------------------------------------------------------------------------------
constexpr int kHeadMaxCin = 64;

template <typename T>
__global__ void head_fwd_kernel(const T* __restrict__ x, long long xp, const float* __restrict__ w,
                                const float* __restrict__ bias, float* __restrict__ logits, long long V,
                                int Cin, int C) {
  extern __shared__ float sm[];  // [C][Cin] + [C]
  for (int i = threadIdx.x; i < C * Cin; i += blockDim.x) sm[i] = w[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sm[C * Cin + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int n = blockIdx.y;
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < V;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    float f[kHeadMaxCin];
    const T* px = x + (static_cast<long long>(n) * V + v) * xp;
#pragma unroll
    for (int k = 0; k < kHeadMaxCin; k += 8) {
      if (k < Cin) {
        float t[8];
        Vec8<T>::load(px + k, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[k + j] = t[j];
      }
    }
    for (int c = 0; c < C; ++c) {
      float acc = sm[C * Cin + c];
      const float* wr = sm + c * Cin;
#pragma unroll
      for (int k = 0; k < kHeadMaxCin; ++k)
        if (k < Cin) acc = fmaf(f[k], wr[k], acc);
      logits[(static_cast<long long>(n) * C + c) * V + v] = acc;
    }
  }
}

// dx[v][ci] = sum_c dl[c][v] w[c][ci]
template <typename T>
__global__ void head_bwd_data_kernel(const float* __restrict__ dl, const float* __restrict__ w,
                                     T* __restrict__ dx, long long dxp, long long V, int Cin, int C) {
  extern __shared__ float sm[];
  for (int i = threadIdx.x; i < C * Cin; i += blockDim.x) sm[i] = w[i];
  __syncthreads();
  const int n = blockIdx.y;
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < V;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    float acc[kHeadMaxCin];
#pragma unroll
    for (int k = 0; k < kHeadMaxCin; ++k) acc[k] = 0.f;
    for (int c = 0; c < C; ++c) {
      const float g = dl[(static_cast<long long>(n) * C + c) * V + v];
      const float* wr = sm + c * Cin;
#pragma unroll
      for (int k = 0; k < kHeadMaxCin; ++k)
        if (k < Cin) acc[k] = fmaf(g, wr[k], acc[k]);
    }
    T* pd = dx + (static_cast<long long>(n) * V + v) * dxp;
#pragma unroll
    for (int k = 0; k < kHeadMaxCin; k += 8) {
      if (k < Cin) {
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = acc[k + j];
        Vec8<T>::store(pd + k, t);
      }
    }
  }
}

// dW[c][ci] = sum_v dl[c][v] x[v][ci];  db[c] = sum_v dl[c][v].
// Block stages 128 voxels of x (fp32) and dl in smem; thread (c, ci) reduces over them.
constexpr int kHeadWgTile = 128;
template <typename T>
__global__ void head_bwd_weight_kernel(const T* __restrict__ x, long long xp, const float* __restrict__ dl,
                                       float* __restrict__ dw, float* __restrict__ db, long long V, int Cin,
                                       int C, int tiles_per_block) {
  extern __shared__ float sm[];
  float* sx = sm;                          // [tile][Cin+1]
  float* sd = sm + kHeadWgTile * (Cin + 1);  // [C][tile]
  const int n = blockIdx.y;
  const int CG = Cin / 8;
  const int nout = C * Cin + C;
  // every thread may own several outputs
  float acc[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) acc[q] = 0.f;
  for (int t = 0; t < tiles_per_block; ++t) {
    const long long vt = (static_cast<long long>(blockIdx.x) * tiles_per_block + t) * kHeadWgTile;
    if (vt >= V) break;
    __syncthreads();
    for (int i = threadIdx.x; i < kHeadWgTile * CG; i += blockDim.x) {
      const int vi = i / CG, cg = i % CG;
      float f[8];
      if (vt + vi < V) {
        Vec8<T>::load(x + (static_cast<long long>(n) * V + vt + vi) * xp + cg * 8, f);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) sx[vi * (Cin + 1) + cg * 8 + j] = f[j];
    }
    for (int i = threadIdx.x; i < C * kHeadWgTile; i += blockDim.x) {
      const int c = i / kHeadWgTile, vi = i % kHeadWgTile;
      sd[i] = (vt + vi < V) ? dl[(static_cast<long long>(n) * C + c) * V + vt + vi] : 0.f;
    }
    __syncthreads();
    int q = 0;
    for (int o = threadIdx.x; o < nout && q < 8; o += blockDim.x, ++q) {
      float a = 0.f;
      if (o < C * Cin) {
        const int c = o / Cin, ci = o % Cin;
        for (int vi = 0; vi < kHeadWgTile; ++vi) a = fmaf(sd[c * kHeadWgTile + vi], sx[vi * (Cin + 1) + ci], a);
      } else {
        const int c = o - C * Cin;
        for (int vi = 0; vi < kHeadWgTile; ++vi) a += sd[c * kHeadWgTile + vi];
      }
      acc[q] += a;
    }
  }
  int q = 0;
  for (int o = threadIdx.x; o < nout && q < 8; o += blockDim.x, ++q) {
    if (o < C * Cin) atomicAdd(&dw[o], acc[q]);
    else atomicAdd(&db[o - C * Cin], acc[q]);
  }
}

// ------------------------------------------------------------------------------------------
// head, (voxel, 8-channel group) thread mapping: every global access is a fully coalesced 128-bit access (the
// one-voxel-per-thread kernels above read 64-byte-strided 16-byte pieces and ran at ~1.7 TB/s); the Cin reduction of
// the forward finishes with CGC-lane shuffles, the backward fuses dX, dW and db into one pass over (x, dlogits).
// ------------------------------------------------------------------------------------------
constexpr int kHeadCgMaxC = 8;

template <typename T, int CGC>
__global__ void __launch_bounds__(256) head_fwd_cg_kernel(const T* __restrict__ x, long long xp, const float* __restrict__ w,
                                                          const float* __restrict__ bias, float* __restrict__ logits,
                                                          long long V, int C) {
  extern __shared__ float sm[];  // [C][Cin] + [C]
  constexpr int Cin = CGC * 8;
  for (int i = threadIdx.x; i < C * Cin; i += blockDim.x) sm[i] = w[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sm[C * Cin + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int n = blockIdx.y;
  const int cg = threadIdx.x % CGC, vl = threadIdx.x / CGC;
  constexpr int vpw = 32 / CGC;                      // voxels per warp
  const long long vstride = static_cast<long long>(gridDim.x) * (256 / CGC);
  // warp-uniform trip count: the shuffles below need all 32 lanes
  for (long long vb = static_cast<long long>(blockIdx.x) * (256 / CGC) + (vl / vpw) * vpw; vb < V; vb += vstride) {
    const long long v = vb + vl % vpw;
    const bool ok = v < V;
    float f[8] = {0};
    if (ok) Vec8<T>::load(x + (static_cast<long long>(n) * V + v) * xp + cg * 8, f);
    for (int c = 0; c < C; ++c) {
      const float* wr = sm + c * Cin + cg * 8;
      float p = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) p = fmaf(f[j], wr[j], p);
#pragma unroll
      for (int o = CGC / 2; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
      if (ok && cg == 0) logits[(static_cast<long long>(n) * C + c) * V + v] = p + sm[C * Cin + c];
    }
  }
}

template <typename T, int CGC, int CMAX>
__global__ void __launch_bounds__(256) head_bwd_cg_kernel(const T* __restrict__ x, long long xp, const float* __restrict__ w,
                                                          const float* __restrict__ dl, T* __restrict__ dx, long long dxp,
                                                          float* __restrict__ dw, float* __restrict__ db, long long V, int C) {
  extern __shared__ float sm[];  // w [C][Cin] | dW partial [C][Cin] | db partial [C]
  constexpr int Cin = CGC * 8;
  constexpr int U = 4;           // voxels in flight per thread (memory-level parallelism)
  float* sw = sm;
  float* sacc = sm + C * Cin;
  for (int i = threadIdx.x; i < C * Cin; i += blockDim.x) { sw[i] = w[i]; sacc[i] = 0.f; }
  for (int i = threadIdx.x; i < C; i += blockDim.x) sacc[C * Cin + i] = 0.f;
  __syncthreads();
  const int n = blockIdx.y;
  const int cg = threadIdx.x % CGC;
  const long long vstride = static_cast<long long>(gridDim.x) * (256 / CGC);
  float aw[CMAX][8], ab[CMAX];
#pragma unroll
  for (int c = 0; c < CMAX; ++c) {
    ab[c] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) aw[c][j] = 0.f;
  }
  for (long long v0 = static_cast<long long>(blockIdx.x) * (256 / CGC) + threadIdx.x / CGC; v0 < V; v0 += U * vstride) {
    Raw8<T> fx[U];
    float g[U][CMAX];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long v = v0 + u * vstride;
      if (v < V) {
        fx[u].load(x + (static_cast<long long>(n) * V + v) * xp + cg * 8);
#pragma unroll
        for (int c = 0; c < CMAX; ++c) g[u][c] = c < C ? dl[(static_cast<long long>(n) * C + c) * V + v] : 0.f;
      } else {
        fx[u].zero();
#pragma unroll
        for (int c = 0; c < CMAX; ++c) g[u][c] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long v = v0 + u * vstride;
      float f[8], o[8] = {0};
      fx[u].to_float(f);
#pragma unroll
      for (int c = 0; c < CMAX; ++c) {
        if (c < C) {
          const float* wr = sw + c * Cin + cg * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            o[j] = fmaf(g[u][c], wr[j], o[j]);
            aw[c][j] = fmaf(g[u][c], f[j], aw[c][j]);
          }
          ab[c] += g[u][c];
        }
      }
      if (v < V) Vec8<T>::store(dx + (static_cast<long long>(n) * V + v) * dxp + cg * 8, o);
    }
  }
#pragma unroll
  for (int c = 0; c < CMAX; ++c) {
    if (c < C) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&sacc[c * Cin + cg * 8 + j], aw[c][j]);
      if (cg == 0) atomicAdd(&sacc[C * Cin + c], ab[c]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * Cin; i += blockDim.x) atomicAdd(&dw[i], sacc[i]);
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&db[i], sacc[C * Cin + i]);
}

template <typename T>
static int launch_head_fwd_cg(int CG, dim3 grid, size_t smb, cudaStream_t st, const T* x, long long xp, const float* w,
                              const float* bias, float* logits, long long V, int C) {
  switch (CG) {
    case 1: head_fwd_cg_kernel<T, 1><<<grid, 256, smb, st>>>(x, xp, w, bias, logits, V, C); break;
    case 2: head_fwd_cg_kernel<T, 2><<<grid, 256, smb, st>>>(x, xp, w, bias, logits, V, C); break;
    case 4: head_fwd_cg_kernel<T, 4><<<grid, 256, smb, st>>>(x, xp, w, bias, logits, V, C); break;
    case 8: head_fwd_cg_kernel<T, 8><<<grid, 256, smb, st>>>(x, xp, w, bias, logits, V, C); break;
    default: return 1;
  }
  return 0;
}
template <typename T, int CMAX>
static int launch_head_bwd_cg2(int CG, dim3 grid, size_t smb, cudaStream_t st, const T* x, long long xp, const float* w,
                               const float* dl, T* dx, long long dxp, float* dw, float* db, long long V, int C) {
  switch (CG) {
    case 1: head_bwd_cg_kernel<T, 1, CMAX><<<grid, 256, smb, st>>>(x, xp, w, dl, dx, dxp, dw, db, V, C); break;
    case 2: head_bwd_cg_kernel<T, 2, CMAX><<<grid, 256, smb, st>>>(x, xp, w, dl, dx, dxp, dw, db, V, C); break;
    case 4: head_bwd_cg_kernel<T, 4, CMAX><<<grid, 256, smb, st>>>(x, xp, w, dl, dx, dxp, dw, db, V, C); break;
    case 8: head_bwd_cg_kernel<T, 8, CMAX><<<grid, 256, smb, st>>>(x, xp, w, dl, dx, dxp, dw, db, V, C); break;
    default: return 1;
  }
  return 0;
}
template <typename T>
static int launch_head_bwd_cg(int CG, dim3 grid, size_t smb, cudaStream_t st, const T* x, long long xp, const float* w,
                              const float* dl, T* dx, long long dxp, float* dw, float* db, long long V, int C) {
  if (C <= 2) return launch_head_bwd_cg2<T, 2>(CG, grid, smb, st, x, xp, w, dl, dx, dxp, dw, db, V, C);
  if (C <= 4) return launch_head_bwd_cg2<T, 4>(CG, grid, smb, st, x, xp, w, dl, dx, dxp, dw, db, V, C);
  return launch_head_bwd_cg2<T, 8>(CG, grid, smb, st, x, xp, w, dl, dx, dxp, dw, db, V, C);
}

}  // namespace rsb

using namespace rsb;

extern "C" int rsb_stem_conv_forward(const float* x, const float* w_oidhw, void* y, int y_pitch, int dtype,
                                     float* out_stats, int N, int D, int H, int W, int Cout, void* stream) {
  RSB_REQUIRE(x && w_oidhw && y, "stem_conv_forward: null pointer");
  RSB_REQUIRE(Cout > 0 && Cout % 8 == 0 && Cout <= 256, "stem_conv: Cout must be a multiple of 8 <= 256 (got %d)", Cout);
  RSB_REQUIRE(N > 0 && N <= 65535 && D > 0 && H > 0 && W > 0, "stem_conv: bad geometry");
  const long long V = static_cast<long long>(D) * H * W;
  const int cpad = (Cout + 15) / 16 * 16;
  const size_t sm = sizeof(float) * (27 * cpad + 2 * Cout);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (W % kStemVX == 0 && (dtype == RSB_BF16 || dtype == RSB_F32)) {
    const long long groups = V / kStemVX;
    dim3 gridt(static_cast<unsigned>((groups + kStemThreads - 1) / kStemThreads), N);
    if (dtype == RSB_BF16)
      stem_fwd_tiled_kernel<__nv_bfloat16><<<gridt, kStemThreads, sm, st>>>(x, w_oidhw, (__nv_bfloat16*)y, y_pitch, out_stats, D, H, W, Cout);
    else
      stem_fwd_tiled_kernel<float><<<gridt, kStemThreads, sm, st>>>(x, w_oidhw, (float*)y, y_pitch, out_stats, D, H, W, Cout);
    return check_launch("stem_conv_forward");
  }
  dim3 grid(static_cast<unsigned>((V + kStemThreads - 1) / kStemThreads), N);
  if (dtype == RSB_BF16)
    stem_fwd_kernel<__nv_bfloat16><<<grid, kStemThreads, sm, st>>>(x, w_oidhw, (__nv_bfloat16*)y, y_pitch, out_stats, D, H, W, Cout);
  else if (dtype == RSB_F32)
    stem_fwd_kernel<float><<<grid, kStemThreads, sm, st>>>(x, w_oidhw, (float*)y, y_pitch, out_stats, D, H, W, Cout);
  else { set_last_error("bad dtype %d", dtype); return -1; }
  return check_launch("stem_conv_forward");
}

extern "C" int rsb_stem_conv_wgrad(const float* x, const void* dy, int dy_pitch, int dtype, float* dw_oidhw,
                                   int N, int D, int H, int W, int Cout, void* stream) {
  RSB_REQUIRE(x && dy && dw_oidhw, "stem_conv_wgrad: null pointer");
  RSB_REQUIRE(Cout > 0 && Cout % 8 == 0 && 27 * (Cout / 8) <= 1024, "stem_conv_wgrad: unsupported Cout %d", Cout);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(dw_oidhw, 0, sizeof(float) * Cout * 27, st);
  RSB_REQUIRE(e == cudaSuccess, "stem_conv_wgrad: memset failed: %s", cudaGetErrorString(e));
  const long long total = static_cast<long long>(N) * D * H * W;
  const int sms = rsb_num_sms();
  RSB_REQUIRE(sms > 0, "no CUDA device");
  if (Cout <= 64) {
    const int tiles_x = (W + kSwTX - 1) / kSwTX, tiles_y = (H + kSwTY - 1) / kSwTY;
    const long long num_tiles = static_cast<long long>(N) * D * tiles_y * tiles_x;
    const size_t smb = sizeof(float) * (3 * 10 * 34 + kSwVox * Cout + 27 * Cout);
    const int threads = 8 * 9 * (Cout / 8);
    long long blocks = static_cast<long long>(sms) * (Cout <= 32 ? 4 : 2);
    if (blocks > num_tiles) blocks = num_tiles;
    cudaError_t ea;
    if (dtype == RSB_BF16) {
      ea = cudaFuncSetAttribute(stem_wgrad_tiled_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smb));
      RSB_REQUIRE(ea == cudaSuccess, "stem_conv_wgrad: smem opt-in failed: %s", cudaGetErrorString(ea));
      stem_wgrad_tiled_kernel<__nv_bfloat16><<<static_cast<unsigned>(blocks), threads, smb, st>>>(x, (const __nv_bfloat16*)dy, dy_pitch, dw_oidhw, N, D, H, W, Cout, tiles_x, tiles_y, num_tiles);
    } else if (dtype == RSB_F32) {
      ea = cudaFuncSetAttribute(stem_wgrad_tiled_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smb));
      RSB_REQUIRE(ea == cudaSuccess, "stem_conv_wgrad: smem opt-in failed: %s", cudaGetErrorString(ea));
      stem_wgrad_tiled_kernel<float><<<static_cast<unsigned>(blocks), threads, smb, st>>>(x, (const float*)dy, dy_pitch, dw_oidhw, N, D, H, W, Cout, tiles_x, tiles_y, num_tiles);
    } else { set_last_error("bad dtype %d", dtype); return -1; }
    return check_launch("stem_wgrad_tiled_kernel");
  }
  long long blocks = static_cast<long long>(sms) * 8;
  long long vpb = (total + blocks - 1) / blocks;
  if (vpb < 64) vpb = 64;
  blocks = (total + vpb - 1) / vpb;
  const int threads = ((27 * (Cout / 8)) + 31) / 32 * 32;
  if (dtype == RSB_BF16)
    stem_wgrad_kernel<__nv_bfloat16><<<static_cast<unsigned>(blocks), threads, 0, st>>>(x, (const __nv_bfloat16*)dy, dy_pitch, dw_oidhw, N, D, H, W, Cout, static_cast<int>(vpb));
  else if (dtype == RSB_F32)
    stem_wgrad_kernel<float><<<static_cast<unsigned>(blocks), threads, 0, st>>>(x, (const float*)dy, dy_pitch, dw_oidhw, N, D, H, W, Cout, static_cast<int>(vpb));
  else { set_last_error("bad dtype %d", dtype); return -1; }
  return check_launch("stem_conv_wgrad");
}

extern "C" int rsb_head_forward(const void* x, int x_pitch, int dtype, const float* w, const float* bias,
                                float* logits_ncdhw, int N, int D, int H, int W, int Cin, int C, void* stream) {
  RSB_REQUIRE(x && w && logits_ncdhw, "head_forward: null pointer");
  RSB_REQUIRE(Cin > 0 && Cin % 8 == 0 && Cin <= kHeadMaxCin, "head: Cin must be a multiple of 8 <= %d (got %d)", kHeadMaxCin, Cin);
  RSB_REQUIRE(C > 0 && C <= 128 && N > 0 && N <= 65535, "head: bad class count / batch");
  const long long V = static_cast<long long>(D) * H * W;
  const int sms = rsb_num_sms();
  RSB_REQUIRE(sms > 0, "no CUDA device");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t sm = sizeof(float) * (C * Cin + C);
  const int CG = Cin / 8;
  if ((CG == 1 || CG == 2 || CG == 4 || CG == 8) && (dtype == RSB_BF16 || dtype == RSB_F32)) {
    const int vpb = 256 / CG;
    long long gx = (V + vpb - 1) / vpb;
    if (gx > sms * 8LL) gx = sms * 8LL;
    dim3 g(static_cast<unsigned>(gx), N);
    if (dtype == RSB_BF16) launch_head_fwd_cg<__nv_bfloat16>(CG, g, sm, st, (const __nv_bfloat16*)x, x_pitch, w, bias, logits_ncdhw, V, C);
    else launch_head_fwd_cg<float>(CG, g, sm, st, (const float*)x, x_pitch, w, bias, logits_ncdhw, V, C);
    return check_launch("head_fwd_cg_kernel");
  }
  long long bx = (V + 255) / 256;
  if (bx > sms * 16LL) bx = sms * 16LL;
  dim3 grid(static_cast<unsigned>(bx), N);
  if (dtype == RSB_BF16)
    head_fwd_kernel<__nv_bfloat16><<<grid, 256, sm, st>>>((const __nv_bfloat16*)x, x_pitch, w, bias, logits_ncdhw, V, Cin, C);
  else if (dtype == RSB_F32)
    head_fwd_kernel<float><<<grid, 256, sm, st>>>((const float*)x, x_pitch, w, bias, logits_ncdhw, V, Cin, C);
  else { set_last_error("bad dtype %d", dtype); return -1; }
  return check_launch("head_forward");
}

extern "C" int rsb_head_backward(const void* x, int x_pitch, int dtype, const float* w,
                                 const float* dlogits_ncdhw, void* dx, int dx_pitch, float* dw, float* db,
                                 int N, int D, int H, int W, int Cin, int C, void* stream) {
  RSB_REQUIRE(x && w && dlogits_ncdhw && dx && dw && db, "head_backward: null pointer");
  RSB_REQUIRE(Cin > 0 && Cin % 8 == 0 && Cin <= kHeadMaxCin, "head: Cin must be a multiple of 8 <= %d (got %d)", kHeadMaxCin, Cin);
  RSB_REQUIRE(C > 0 && C <= 128 && N > 0 && N <= 65535, "head: bad class count / batch");
  RSB_REQUIRE(C * Cin + C <= 256 * 8, "head_backward: too many weights");
  const long long V = static_cast<long long>(D) * H * W;
  const int sms = rsb_num_sms();
  RSB_REQUIRE(sms > 0, "no CUDA device");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * C * Cin, st);
  RSB_REQUIRE(e == cudaSuccess, "head_backward: memset failed: %s", cudaGetErrorString(e));
  e = cudaMemsetAsync(db, 0, sizeof(float) * C, st);
  RSB_REQUIRE(e == cudaSuccess, "head_backward: memset failed: %s", cudaGetErrorString(e));
  {
    const int CG = Cin / 8;
    if ((CG == 1 || CG == 2 || CG == 4 || CG == 8) && C <= kHeadCgMaxC && (dtype == RSB_BF16 || dtype == RSB_F32)) {
      const int vpb = 256 / CG;
      long long gx = (V + vpb - 1) / vpb;
      if (gx > sms * 6LL) gx = sms * 6LL;
      dim3 g(static_cast<unsigned>(gx), N);
      const size_t smb = sizeof(float) * (2 * C * Cin + C);
      if (dtype == RSB_BF16) launch_head_bwd_cg<__nv_bfloat16>(CG, g, smb, st, (const __nv_bfloat16*)x, x_pitch, w, dlogits_ncdhw, (__nv_bfloat16*)dx, dx_pitch, dw, db, V, C);
      else launch_head_bwd_cg<float>(CG, g, smb, st, (const float*)x, x_pitch, w, dlogits_ncdhw, (float*)dx, dx_pitch, dw, db, V, C);
      return check_launch("head_bwd_cg_kernel");
    }
  }
  {
    long long bx = (V + 255) / 256;
    if (bx > sms * 16LL) bx = sms * 16LL;
    dim3 grid(static_cast<unsigned>(bx), N);
    const size_t sm = sizeof(float) * (C * Cin);
    if (dtype == RSB_BF16)
      head_bwd_data_kernel<__nv_bfloat16><<<grid, 256, sm, st>>>(dlogits_ncdhw, w, (__nv_bfloat16*)dx, dx_pitch, V, Cin, C);
    else if (dtype == RSB_F32)
      head_bwd_data_kernel<float><<<grid, 256, sm, st>>>(dlogits_ncdhw, w, (float*)dx, dx_pitch, V, Cin, C);
    else { set_last_error("bad dtype %d", dtype); return -1; }
    int rc = check_launch("head_bwd_data_kernel");
    if (rc) return rc;
  }
  {
    const long long tiles = (V + kHeadWgTile - 1) / kHeadWgTile;
    long long bx = sms * 4LL;
    if (bx > tiles) bx = tiles;
    const int tpb = static_cast<int>((tiles + bx - 1) / bx);
    bx = (tiles + tpb - 1) / tpb;
    dim3 grid(static_cast<unsigned>(bx), N);
    const size_t sm = sizeof(float) * (kHeadWgTile * (Cin + 1) + C * kHeadWgTile);
    RSB_REQUIRE(sm <= 48 * 1024, "head_backward: shared memory budget exceeded");
    if (dtype == RSB_BF16)
      head_bwd_weight_kernel<__nv_bfloat16><<<grid, 256, sm, st>>>((const __nv_bfloat16*)x, x_pitch, dlogits_ncdhw, dw, db, V, Cin, C, tpb);
    else
      head_bwd_weight_kernel<float><<<grid, 256, sm, st>>>((const float*)x, x_pitch, dlogits_ncdhw, dw, db, V, Cin, C, tpb);
  }
  return check_launch("head_bwd_weight_kernel");
}
