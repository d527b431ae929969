// report_loss.cu — device side of the report-supervised losses of R-Super:
//   volume_loss_basic / dice_based_volume_loss        rsuper_train/training/losses_foundation.py:250-395
//   ball_loss / isolate_tumor / insert_ball / GWRP    losses_foundation.py:442-537, 1336-1864
// The reference builds these from dozens of full-volume torch temporaries, a dense F.conv3d with a ball kernel
// of up to 39^3 taps, torch.topk x3 and a full-volume sort per tumour.  Here every per-voxel step is one kernel:
//   rows_gather / rows_scatter_add     lesion-channel selection (get_lesion_channels, :204-248) and its adjoint
//   u8_binary / u8_row_count           mask algebra (to_penalize, pseudo / border masks) and .sum() > 0 tests
//   masked_sigmoid_sum / _grad         Volume loss reduction  V^ = sum sigmoid(x) * mask  and its backward
//   ball_occupancy + ball_correlate    Gaussian-ball cross-correlation restricted to tiles that can be non-zero,
//                                      fused with the argmax (first maximum in flattened order)
//   ball_candidates + ball_rank_select top-{t, t_small, t_big} membership inside the ball by exact ranking
//                                      (value descending, index ascending) instead of three radix selects
//   ball_rank_gwrp                     GlobalWeightedRankPooling weights d^rank / sum, scattered to voxel order
// Scalar glue (the 10-tumour loop, volume formulas on [B,L] values) stays on the host like in the reference.
#include "rsb_common.cuh"

#include <cmath>

#include "../../include/rsuper_b200.h"

namespace rsb {

RSB_DEVICE float sigmoidf_r(float x) { return 1.f / (1.f + expf(-x)); }

static inline int grid_for(long long n, int block, int waves = 8) {
  const int sms = rsb_num_sms();
  long long want = (n + block - 1) / block;
  const long long cap = static_cast<long long>(sms > 0 ? sms : 148) * waves;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return static_cast<int>(want);
}

// ---------------------------------------------------------------------------------------------
// row gather / scatter-add: dst[r][v] = src[row_map[r]][v]   (rows of V elements)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void rows_gather_kernel(const T* __restrict__ src, const int* __restrict__ row_map, T* __restrict__ dst, long long V) {
  const long long r = blockIdx.y;
  const T* s = src + static_cast<long long>(row_map[r]) * V;
  T* d = dst + r * V;
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < V; v += static_cast<long long>(gridDim.x) * blockDim.x)
    d[v] = s[v];
}
__global__ void rows_scatter_add_kernel(const float* __restrict__ src, const int* __restrict__ row_map, float* __restrict__ dst,
                                        long long V) {
  const long long r = blockIdx.y;
  const float* s = src + r * V;
  float* d = dst + static_cast<long long>(row_map[r]) * V;
  // atomic: a channel that belongs to two lesion groups ('kidney_cyst_lesion' matches both suffixes) is the destination of two
  // rows, and rows run concurrently
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < V; v += static_cast<long long>(gridDim.x) * blockDim.x)
    atomicAdd(&d[v], s[v]);
}

// Lesion groups that merge several channels of one organ (get_lesion_channels, :215-220: torch.stack(...).max(dim=0)):
// gmap[r * G + k] = source row of member k of output row r (-1 = no such member).  uint8 masks: max == OR.
template <typename T>
__global__ void rows_gather_max_kernel(const T* __restrict__ src, const int* __restrict__ gmap, int G, T* __restrict__ dst, long long V) {
  const long long r = blockIdx.y;
  const int* g = gmap + r * G;
  T* d = dst + r * V;
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < V; v += static_cast<long long>(gridDim.x) * blockDim.x) {
    T best = src[static_cast<long long>(g[0]) * V + v];
    for (int k = 1; k < G; ++k)
      if (g[k] >= 0) {
        const T o = src[static_cast<long long>(g[k]) * V + v];
        best = o > best ? o : best;
      }
    d[v] = best;
  }
}
// backward of the max-merge: the gradient of output row r goes to the member that attained the maximum (first one on ties)
__global__ void rows_scatter_add_max_kernel(const float* __restrict__ grad, const float* __restrict__ x, const int* __restrict__ gmap, int G,
                                            float* __restrict__ dst, long long V) {
  const long long r = blockIdx.y;
  const int* g = gmap + r * G;
  const float* s = grad + r * V;
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < V; v += static_cast<long long>(gridDim.x) * blockDim.x) {
    int arg = g[0];
    float best = x[static_cast<long long>(arg) * V + v];
    for (int k = 1; k < G; ++k)
      if (g[k] >= 0) {
        const float o = x[static_cast<long long>(g[k]) * V + v];
        if (o > best) { best = o; arg = g[k]; }
      }
    atomicAdd(&dst[static_cast<long long>(arg) * V + v], s[v]);
  }
}

// ---------------------------------------------------------------------------------------------
// mask algebra on uint8 0/1 volumes; op: 0 OR, 1 AND, 2 a AND NOT b, 3 NOR, 4 NOT a
// ---------------------------------------------------------------------------------------------
__global__ void u8_binary_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, uint8_t* __restrict__ out, int op,
                                 long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const bool x = a[i] != 0, y = b ? (b[i] != 0) : false;
    bool r;
    switch (op) {
      case 0: r = x || y; break;
      case 1: r = x && y; break;
      case 2: r = x && !y; break;
      case 3: r = !(x || y); break;
      default: r = !x; break;
    }
    out[i] = r ? 1 : 0;
  }
}

__global__ void u8_row_count_kernel(const uint8_t* __restrict__ a, long long* __restrict__ counts, long long V) {
  const long long r = blockIdx.y;
  const uint8_t* s = a + r * V;
  long long c = 0;
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < V; v += static_cast<long long>(gridDim.x) * blockDim.x)
    c += s[v] ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(reinterpret_cast<unsigned long long*>(&counts[r]), static_cast<unsigned long long>(c));
}

// ---------------------------------------------------------------------------------------------
// Volume loss reduction: sums[r] = sum_v sigmoid(x[r][v]) * scale[r] * mask[r][v]   and its backward
//   dx[r][v] (+)= coef[r] * scale[r] * mask * sig * (1 - sig)
// ---------------------------------------------------------------------------------------------
__global__ void masked_sigmoid_sum_kernel(const float* __restrict__ x, const uint8_t* __restrict__ mask, const float* __restrict__ scale,
                                          float* __restrict__ sums, long long V) {
  __shared__ float red[8];
  const long long r = blockIdx.y;
  const float* xr = x + r * V;
  const uint8_t* mr = mask + r * V;
  float s = 0.f;
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < V; v += static_cast<long long>(gridDim.x) * blockDim.x)
    if (mr[v]) s += sigmoidf_r(xr[v]);
  s = warp_sum(s);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (warp == 0) {
    float t = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0 && t != 0.f) atomicAdd(&sums[r], t * (scale ? scale[r] : 1.f));
  }
}
__global__ void masked_sigmoid_grad_kernel(const float* __restrict__ x, const uint8_t* __restrict__ mask, const float* __restrict__ scale,
                                           const float* __restrict__ coef, float* __restrict__ dx, int accumulate, long long V) {
  const long long r = blockIdx.y;
  const float c = coef[r] * (scale ? scale[r] : 1.f);
  const float* xr = x + r * V;
  const uint8_t* mr = mask + r * V;
  float* dr = dx + r * V;
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < V; v += static_cast<long long>(gridDim.x) * blockDim.x) {
    float g = 0.f;
    if (mr[v]) {
      const float sg = sigmoidf_r(xr[v]);
      g = c * sg * (1.f - sg);
    }
    dr[v] = accumulate ? dr[v] + g : g;
  }
}

// ---------------------------------------------------------------------------------------------
// Ball loss: x_iter = sigmoid(x) * seg   and   x_iter *= (1 - mask)
// ---------------------------------------------------------------------------------------------
__global__ void ball_prepare_kernel(const float* __restrict__ x, const uint8_t* __restrict__ seg, float* __restrict__ out, long long V) {
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < V; v += static_cast<long long>(gridDim.x) * blockDim.x)
    out[v] = seg[v] ? sigmoidf_r(x[v]) : 0.f;
}
__global__ void ball_remove_kernel(float* __restrict__ x_iter, const uint8_t* __restrict__ mask, long long V) {
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < V; v += static_cast<long long>(gridDim.x) * blockDim.x)
    if (mask[v]) x_iter[v] = 0.f;
}

// occupancy of 8^3 cells: occ[cell] = any(x_iter > 0)
__global__ void ball_occupancy_kernel(const float* __restrict__ x, uint8_t* __restrict__ occ, int D, int H, int W, int cd, int ch, int cw) {
  const int cell = blockIdx.x;
  const int cx = cell % cw, cy = (cell / cw) % ch, cz = cell / (cw * ch);
  bool any = false;
  for (int i = threadIdx.x; i < 512; i += blockDim.x) {
    const int z = cz * 8 + (i >> 6), y = cy * 8 + ((i >> 3) & 7), xx = cx * 8 + (i & 7);
    if (z < D && y < H && xx < W && x[(static_cast<long long>(z) * H + y) * W + xx] > 0.f) any = true;
  }
  any = __syncthreads_or(any);
  if (threadIdx.x == 0) occ[cell] = any ? 1 : 0;
}

// score = cross-correlation of x_iter with the (Gaussian) ball, zero padding; fused argmax.
// taps: int4 {dz, dy, dx, float bits of the weight}.  One block per 8^3 output tile; tiles whose reach holds no
// occupied cell score exactly 0 and are skipped.  best[0] = packed (score bits << 32 | ~index) maximised with atomicMax:
// scores are >= 0 so their float bits order like unsigned ints, and ~index makes the FIRST flattened index win ties.
__global__ void __launch_bounds__(512) ball_correlate_kernel(const float* __restrict__ x, const uint8_t* __restrict__ occ,
                                                             const int4* __restrict__ taps, int n_taps, int reach_cells,
                                                             unsigned long long* __restrict__ best, int D, int H, int W, int cd,
                                                             int ch, int cw) {
  const int cell = blockIdx.x;
  const int cx = cell % cw, cy = (cell / cw) % ch, cz = cell / (cw * ch);
  __shared__ int s_any;
  if (threadIdx.x == 0) s_any = 0;
  __syncthreads();
  const int span = 2 * reach_cells + 1;
  for (int i = threadIdx.x; i < span * span * span; i += blockDim.x) {
    const int z = cz - reach_cells + i / (span * span), y = cy - reach_cells + (i / span) % span, xx = cx - reach_cells + i % span;
    if (z >= 0 && z < cd && y >= 0 && y < ch && xx >= 0 && xx < cw && occ[(z * ch + y) * cw + xx]) s_any = 1;
  }
  __syncthreads();
  const int z = cz * 8 + (threadIdx.x >> 6), y = cy * 8 + ((threadIdx.x >> 3) & 7), xq = cx * 8 + (threadIdx.x & 7);
  const bool inside = z < D && y < H && xq < W;
  float acc = 0.f;
  if (s_any && inside) {
    for (int t = 0; t < n_taps; ++t) {
      const int4 tp = taps[t];
      const int zz = z + tp.x, yy = y + tp.y, xx = xq + tp.z;
      if (zz >= 0 && zz < D && yy >= 0 && yy < H && xx >= 0 && xx < W)
        acc = fmaf(x[(static_cast<long long>(zz) * H + yy) * W + xx], __int_as_float(tp.w), acc);
    }
  }
  // every lane takes part in the warp reduction (lanes outside the volume carry key 0)
  unsigned long long key = 0ull;
  if (inside) {
    const unsigned long long idx = (static_cast<unsigned long long>(z) * H + y) * W + xq;
    key = (static_cast<unsigned long long>(__float_as_uint(acc)) << 32) | (0xFFFFFFFFull - idx);
  }
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
    key = other > key ? other : key;
  }
  if ((threadIdx.x & 31) == 0 && key != 0ull) atomicMax(best, key);
}

// ---------------------------------------------------------------------------------------------
// The same correlation in three separable stages (rsb_ball_correlate_argmax_sep).  The reference's kernel is a Gaussian
// truncated to a ball, k(d) = exp(-|d|^2 / 2 sigma^2) [|d|^2 <= r^2] / Z (create_ball_kernel, :1161-1232).  The Gaussian
// factorises, the ball does not — but the ball is, for every (dz, dy), the x range |dx| <= w(dz, dy), so
//   rows    R_w(z, y, x) = sum_{|dx| <= w} g(dx) x(z, y, x + dx)                       w = 0 .. R   (cumulative in w)
//   discs   D_a(z, y, x) = sum_{dy : a^2 + dy^2 <= r^2} g(dy) R_{w(a, |dy|)}(z, y + dy, x)   a = |dz| = 0 .. R
//   score   S(z, y, x)   = sum_{|dz| <= R} g(dz) D_{|dz|}(z + dz, y, x)
// is the identical sum with (R+1) + ~1.6 R^2 + (2R+1) multiply-adds per voxel instead of ~4.2 R^3 (R = 15: ~400 vs 15.6 k).
// 1 / Z is dropped: only the argmax is used.  Rows of x that are empty are skipped by all stages (row occupancy).
// ---------------------------------------------------------------------------------------------
constexpr int kSepMaxR = 64;

// rowocc[z * H + y] = any(x[z, y, :] > 0)
__global__ void ball_row_occupancy_kernel(const float* __restrict__ x, uint8_t* __restrict__ rowocc, int W, long long rows) {
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * W;
  bool any = false;
  for (int i = threadIdx.x & 31; i < W; i += 32) any |= xr[i] > 0.f;
  any = __any_sync(0xffffffffu, any);
  if ((threadIdx.x & 31) == 0) rowocc[row] = any ? 1 : 0;
}

// one block per occupied row: Rw[w][row][x] for w = 0..R
__global__ void __launch_bounds__(128) ball_rows_kernel(const float* __restrict__ x, const uint8_t* __restrict__ rowocc, const float* __restrict__ g,
                                                        float* __restrict__ Rw, int R, int W, long long V) {
  extern __shared__ float srow[];   // W + 2R, zero padded
  const long long row = blockIdx.x;
  if (!rowocc[row]) return;
  for (int i = threadIdx.x; i < W + 2 * R; i += blockDim.x) {
    const int xx = i - R;
    srow[i] = (xx >= 0 && xx < W) ? x[row * W + xx] : 0.f;
  }
  __syncthreads();
  for (int xq = threadIdx.x; xq < W; xq += blockDim.x) {
    float acc = g[0] * srow[xq + R];
    float* out = Rw + row * W + xq;
    out[0] = acc;
    for (int w = 1; w <= R; ++w) {
      acc = fmaf(g[w], srow[xq + R + w] + srow[xq + R - w], acc);
      out[static_cast<long long>(w) * V] = acc;
    }
  }
}

// one block per (z, y): Da[a][z][y][x] for a = 0..R; docc[z*H+y] = any contributing row occupied
__global__ void __launch_bounds__(128) ball_discs_kernel(const float* __restrict__ Rw, const uint8_t* __restrict__ rowocc, const float* __restrict__ g,
                                                         const int* __restrict__ wtab, float* __restrict__ Da, uint8_t* __restrict__ docc, int R,
                                                         int H, int W, long long V) {
  __shared__ int s_any;
  const long long row = blockIdx.x;
  const int y = static_cast<int>(row % H);
  const long long zrow0 = row - y;   // z * H
  if (threadIdx.x == 0) {
    int any = 0;
    for (int dy = -R; dy <= R && !any; ++dy) {
      const int yy = y + dy;
      if (yy >= 0 && yy < H && rowocc[zrow0 + yy]) any = 1;
    }
    s_any = any;
    docc[row] = static_cast<uint8_t>(any);
  }
  __syncthreads();
  if (!s_any) return;
  for (int xq = threadIdx.x; xq < W; xq += blockDim.x) {
    for (int a = 0; a <= R; ++a) {
      float acc = 0.f;
      for (int dy = -R; dy <= R; ++dy) {
        const int ady = dy < 0 ? -dy : dy;
        const int w = wtab[a * (R + 1) + ady];
        const int yy = y + dy;
        if (w < 0 || yy < 0 || yy >= H || !rowocc[zrow0 + yy]) continue;
        acc = fmaf(g[ady], Rw[static_cast<long long>(w) * V + (zrow0 + yy) * W + xq], acc);
      }
      Da[static_cast<long long>(a) * V + row * W + xq] = acc;
    }
  }
}

// Tiled disc stage for R <= 16 (tumour diameters up to 33 voxels).  The per-row kernel above reads every R_w value of its
// ~1.6 R^2 taps from L2 (ncu: 1.27 ms at 128^3, R = 15, 2.7 % of HBM, all warps waiting on loads); here a block stages the
// R_w rows of a (16 y + halo) x 32 x tile of one z-plane in shared memory once and serves all taps from there.
constexpr int kDiscTY = 16, kDiscTX = 32, kDiscMaxR = 16;
struct DiscTab {                               // by value (constant bank): uniform reads that cost no shared-memory issue slot
  signed char w[kDiscMaxR + 1][kDiscMaxR + 1];   // w[a][|dy|], -1 = outside the ball
  float g[kDiscMaxR + 1];
};
__global__ void __launch_bounds__(256) ball_discs_tiled_kernel(const float* __restrict__ Rw, const uint8_t* __restrict__ rowocc,
                                                               const DiscTab tab, float* __restrict__ Da, uint8_t* __restrict__ docc,
                                                               int R, int H, int W, long long V) {
  extern __shared__ float s_tile[];                       // [R + 1][TY + 2R][TX]
  __shared__ uint8_t s_occ[kDiscTY + 2 * kDiscMaxR];
  __shared__ int s_any;
  const int rows = kDiscTY + 2 * R;
  const int x0 = blockIdx.x * kDiscTX, y0 = blockIdx.y * kDiscTY, z = blockIdx.z;
  const long long zrow0 = static_cast<long long>(z) * H;
  if (threadIdx.x == 0) s_any = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < rows; i += blockDim.x) {
    const int yy = y0 - R + i;
    const uint8_t o = (yy >= 0 && yy < H) ? rowocc[zrow0 + yy] : 0;
    s_occ[i] = o;
    if (o) s_any = 1;
  }
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 8 row groups x 32 x
  const int xq = x0 + tx;
  // docc: any contributing row occupied (per output row)
  if (threadIdx.x < kDiscTY && y0 + threadIdx.x < H) {
    int any = 0;
    for (int d = 0; d <= 2 * R && !any; ++d) any = s_occ[threadIdx.x + d];
    docc[zrow0 + y0 + threadIdx.x] = static_cast<uint8_t>(any);
  }
  if (!s_any) return;                                         // nothing within reach of this tile: D rows are never read
  for (int w = 0; w <= R; ++w)
    for (int i = ty; i < rows; i += 8) {
      const int yy = y0 - R + i;
      float v = 0.f;
      if (s_occ[i] && xq < W) v = Rw[static_cast<long long>(w) * V + (zrow0 + yy) * W + xq];
      s_tile[(w * rows + i) * kDiscTX + tx] = v;
    }
  __syncthreads();
#pragma unroll 1
  for (int k = 0; k < kDiscTY / 8; ++k) {
    const int yl = ty + 8 * k, y = y0 + yl;
    if (y >= H || xq >= W) continue;
    float acc[kDiscMaxR + 1];
#pragma unroll
    for (int a = 0; a <= kDiscMaxR; ++a) acc[a] = 0.f;
    for (int dy = -R; dy <= R; ++dy) {
      const int rl = yl + dy + R;
      if (!s_occ[rl]) continue;
      const int ady = dy < 0 ? -dy : dy;
      const float gy = tab.g[ady];
      // w(a, |dy|) does not increase with a: the row value is re-read from shared memory only when w changes (~9 of 17 times)
      int wprev = -2;
      float val = 0.f;
#pragma unroll
      for (int a = 0; a <= kDiscMaxR; ++a) {
        if (a <= R) {
          const int w = tab.w[a][ady];
          if (w >= 0) {
            if (w != wprev) { val = s_tile[(w * rows + rl) * kDiscTX + tx]; wprev = w; }
            acc[a] = fmaf(gy, val, acc[a]);
          }
        }
      }
    }
    const long long o = (zrow0 + y) * W + xq;
#pragma unroll
    for (int a = 0; a <= kDiscMaxR; ++a)
      if (a <= R) Da[static_cast<long long>(a) * V + o] = acc[a];
  }
}

// one block per (z, y): score = sum_dz g(dz) D_|dz|(z + dz, y, x), fused first-maximum argmax (same packed key as above)
__global__ void __launch_bounds__(128) ball_planes_argmax_kernel(const float* __restrict__ Da, const uint8_t* __restrict__ docc,
                                                                 const float* __restrict__ g, unsigned long long* __restrict__ best, int R, int D,
                                                                 int H, int W, long long V) {
  const long long row = blockIdx.x;
  const int z = static_cast<int>(row / H), y = static_cast<int>(row % H);
  unsigned long long key = 0ull;
  for (int xq = threadIdx.x; xq < W; xq += blockDim.x) {
    float acc = 0.f;
    for (int dz = -R; dz <= R; ++dz) {
      const int zz = z + dz, a = dz < 0 ? -dz : dz;
      if (zz < 0 || zz >= D) continue;
      const long long r2 = static_cast<long long>(zz) * H + y;
      if (!docc[r2]) continue;
      acc = fmaf(g[a], Da[static_cast<long long>(a) * V + r2 * W + xq], acc);
    }
    const unsigned long long idx = static_cast<unsigned long long>(row) * W + xq;
    const unsigned long long k = (static_cast<unsigned long long>(__float_as_uint(acc)) << 32) | (0xFFFFFFFFull - idx);
    key = k > key ? k : key;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
    key = other > key ? other : key;
  }
  if ((threadIdx.x & 31) == 0 && key != 0ull) atomicMax(best, key);
}

// candidates inside the clipped ball: cand[i] = {value bits, voxel index}; ball = centre + odd(ceil(d)) rule of
// create_ball_kernel / insert_ball (d2 <= radius^2 inside a grid of half-width `half`).
//   mode 0: value = x_iter (only > 0 kept)            mode 1: candidates = voxels with mask != 0, value = sigmoid(x)
__global__ void ball_candidates_kernel(const float* __restrict__ x, const uint8_t* __restrict__ mask, int mode, int cz, int cy, int cx,
                                       int half, float radius2, uint2* __restrict__ cand, int* __restrict__ n_cand, int max_cand,
                                       uint8_t* __restrict__ ball_out, int D, int H, int W) {
  const long long V = static_cast<long long>(D) * H * W;
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < V; v += static_cast<long long>(gridDim.x) * blockDim.x) {
    float val;
    bool take;
    if (mode == 0) {
      const int xx = static_cast<int>(v % W), y = static_cast<int>((v / W) % H), z = static_cast<int>(v / (static_cast<long long>(W) * H));
      const int dz = z - cz, dy = y - cy, dx = xx - cx;
      const bool in_ball = abs(dz) <= half && abs(dy) <= half && abs(dx) <= half &&
                           static_cast<float>(dz * dz + dy * dy + dx * dx) <= radius2;
      if (ball_out) ball_out[v] = in_ball ? 1 : 0;
      val = x[v];
      take = in_ball && val > 0.f;
    } else {
      take = mask[v] != 0;
      val = take ? sigmoidf_r(x[v]) : 0.f;
    }
    if (take) {
      const int slot = atomicAdd(n_cand, 1);
      if (slot < max_cand) cand[slot] = make_uint2(__float_as_uint(val), static_cast<unsigned>(v));
    }
  }
}

// exact rank of every candidate: #{j : v_j > v_i or (v_j == v_i and idx_j < idx_i)}   (values >= 0: bit order = value order)
RSB_DEVICE int cand_rank(const uint2* __restrict__ cand, int n, uint2 me, uint2* tile) {
  int rank = 0;
  for (int base = 0; base < n; base += blockDim.x) {
    const int j = base + threadIdx.x;
    tile[threadIdx.x] = j < n ? cand[j] : make_uint2(0u, 0xFFFFFFFFu);
    __syncthreads();
    const int lim = min(static_cast<int>(blockDim.x), n - base);
    for (int q = 0; q < lim; ++q) {
      const uint2 o = tile[q];
      rank += (o.x > me.x || (o.x == me.x && o.y < me.y)) ? 1 : 0;
    }
    __syncthreads();
  }
  return rank;
}

// masks[0..2][idx] = rank < k[0..2]  (mask / small / big of isolate_tumor); volumes pre-zeroed by the caller
__global__ void __launch_bounds__(256) ball_rank_select_kernel(const uint2* __restrict__ cand, const int* __restrict__ n_cand, int max_cand,
                                                               int k0, int k1, int k2, uint8_t* __restrict__ m0, uint8_t* __restrict__ m1,
                                                               uint8_t* __restrict__ m2) {
  __shared__ uint2 tile[256];
  const int n = min(*n_cand, max_cand);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (blockIdx.x * blockDim.x >= n) return;
  const uint2 me = i < n ? cand[i] : make_uint2(0u, 0xFFFFFFFFu);
  const int rank = cand_rank(cand, n, me, tile);
  if (i < n) {
    if (rank < k0) m0[me.y] = 1;
    if (rank < k1) m1[me.y] = 1;
    if (rank < k2) m2[me.y] = 1;
  }
}

// GWRP (losses_foundation.py:442-537, hard_cutoff): w_rank = d^rank / sum_{i<N} d^i, d = (1-c)^(1/N); written as
// wmap[idx] = w_rank * N  (the reference multiplies the weights by pseudo.sum()), evaluated in double.
__global__ void __launch_bounds__(256) ball_rank_gwrp_kernel(const uint2* __restrict__ cand, const int* __restrict__ n_cand, int max_cand,
                                                             float concentration, float* __restrict__ wmap) {
  __shared__ uint2 tile[256];
  const int n = min(*n_cand, max_cand);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (blockIdx.x * blockDim.x >= n) return;
  const uint2 me = i < n ? cand[i] : make_uint2(0u, 0xFFFFFFFFu);
  const int rank = cand_rank(cand, n, me, tile);
  if (i < n) {
    const double N = static_cast<double>(n);
    const double logd = log(1.0 - static_cast<double>(concentration)) / N;
    const double norm = (1.0 - exp(logd * N)) / (1.0 - exp(logd));
    wmap[me.y] = static_cast<float>(exp(logd * rank) / norm * N);
  }
}

// wmap[v] = (pseudo ? wmap[v] : 0) + (1 - dilated)      (fg weights + background indicator of ball_loss, :1775-1811)
__global__ void ball_weight_map_kernel(float* __restrict__ wmap, const uint8_t* __restrict__ pseudo, const uint8_t* __restrict__ dilated,
                                       long long V) {
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < V; v += static_cast<long long>(gridDim.x) * blockDim.x)
    wmap[v] = (pseudo[v] ? wmap[v] : 0.f) + (dilated[v] ? 0.f : 1.f);
}

}  // namespace rsb

using namespace rsb;

#define RSB_ST static_cast<cudaStream_t>(stream)

extern "C" int rsb_rows_gather(const void* src, const int* row_map, void* dst, int n_rows, long long V, int elem_bytes, void* stream) {
  RSB_REQUIRE(src && row_map && dst && n_rows > 0 && n_rows <= 65535 && V > 0, "rows_gather: bad arguments");
  dim3 grid(grid_for(V, 256), n_rows);
  if (elem_bytes == 4) rows_gather_kernel<float><<<grid, 256, 0, RSB_ST>>>((const float*)src, row_map, (float*)dst, V);
  else if (elem_bytes == 1) rows_gather_kernel<uint8_t><<<grid, 256, 0, RSB_ST>>>((const uint8_t*)src, row_map, (uint8_t*)dst, V);
  else { set_last_error("rows_gather: element size must be 1 or 4"); return -1; }
  return check_launch("rows_gather_kernel");
}

extern "C" int rsb_rows_scatter_add(const float* src, const int* row_map, float* dst, int n_rows, long long V, void* stream) {
  RSB_REQUIRE(src && row_map && dst && n_rows > 0 && n_rows <= 65535 && V > 0, "rows_scatter_add: bad arguments");
  dim3 grid(grid_for(V, 256), n_rows);
  rows_scatter_add_kernel<<<grid, 256, 0, RSB_ST>>>(src, row_map, dst, V);
  return check_launch("rows_scatter_add_kernel");
}

extern "C" int rsb_rows_gather_max(const void* src, const int* group_map, int group_size, void* dst, int n_rows, long long V, int elem_bytes,
                                   void* stream) {
  RSB_REQUIRE(src && group_map && dst && n_rows > 0 && n_rows <= 65535 && V > 0 && group_size >= 1, "rows_gather_max: bad arguments");
  dim3 grid(grid_for(V, 256), n_rows);
  if (elem_bytes == 4) rows_gather_max_kernel<float><<<grid, 256, 0, RSB_ST>>>((const float*)src, group_map, group_size, (float*)dst, V);
  else if (elem_bytes == 1) rows_gather_max_kernel<uint8_t><<<grid, 256, 0, RSB_ST>>>((const uint8_t*)src, group_map, group_size, (uint8_t*)dst, V);
  else { set_last_error("rows_gather_max: element size must be 1 or 4"); return -1; }
  return check_launch("rows_gather_max_kernel");
}

extern "C" int rsb_rows_scatter_add_max(const float* grad_rows, const float* x, const int* group_map, int group_size, float* dst, int n_rows,
                                        long long V, void* stream) {
  RSB_REQUIRE(grad_rows && x && group_map && dst && n_rows > 0 && n_rows <= 65535 && V > 0 && group_size >= 1,
              "rows_scatter_add_max: bad arguments");
  dim3 grid(grid_for(V, 256), n_rows);
  rows_scatter_add_max_kernel<<<grid, 256, 0, RSB_ST>>>(grad_rows, x, group_map, group_size, dst, V);
  return check_launch("rows_scatter_add_max_kernel");
}

extern "C" int rsb_u8_binary(const uint8_t* a, const uint8_t* b, uint8_t* out, int op, long long n, void* stream) {
  RSB_REQUIRE(a && out && n > 0 && op >= 0 && op <= 4 && (b || op == 4), "u8_binary: bad arguments");
  u8_binary_kernel<<<grid_for(n, 256), 256, 0, RSB_ST>>>(a, b, out, op, n);
  return check_launch("u8_binary_kernel");
}

extern "C" int rsb_u8_row_count(const uint8_t* a, long long* counts, int n_rows, long long V, void* stream) {
  RSB_REQUIRE(a && counts && n_rows > 0 && n_rows <= 65535 && V > 0, "u8_row_count: bad arguments");
  cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(long long) * n_rows, RSB_ST);
  RSB_REQUIRE(e == cudaSuccess, "u8_row_count: memset failed");
  dim3 grid(grid_for(V, 256, 2), n_rows);
  u8_row_count_kernel<<<grid, 256, 0, RSB_ST>>>(a, counts, V);
  return check_launch("u8_row_count_kernel");
}

extern "C" int rsb_masked_sigmoid_sum(const float* x, const uint8_t* mask, const float* scale, float* sums, int n_rows, long long V,
                                      void* stream) {
  RSB_REQUIRE(x && mask && sums && n_rows > 0 && n_rows <= 65535 && V > 0, "masked_sigmoid_sum: bad arguments");
  cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(float) * n_rows, RSB_ST);
  RSB_REQUIRE(e == cudaSuccess, "masked_sigmoid_sum: memset failed");
  dim3 grid(grid_for(V, 256, 2), n_rows);
  masked_sigmoid_sum_kernel<<<grid, 256, 0, RSB_ST>>>(x, mask, scale, sums, V);
  return check_launch("masked_sigmoid_sum_kernel");
}

extern "C" int rsb_masked_sigmoid_grad(const float* x, const uint8_t* mask, const float* scale, const float* coef, float* dx,
                                       int accumulate, int n_rows, long long V, void* stream) {
  RSB_REQUIRE(x && mask && coef && dx && n_rows > 0 && n_rows <= 65535 && V > 0, "masked_sigmoid_grad: bad arguments");
  dim3 grid(grid_for(V, 256), n_rows);
  masked_sigmoid_grad_kernel<<<grid, 256, 0, RSB_ST>>>(x, mask, scale, coef, dx, accumulate, V);
  return check_launch("masked_sigmoid_grad_kernel");
}

extern "C" int rsb_ball_prepare(const float* x, const uint8_t* seg, float* x_iter, long long V, void* stream) {
  RSB_REQUIRE(x && seg && x_iter && V > 0, "ball_prepare: bad arguments");
  ball_prepare_kernel<<<grid_for(V, 256), 256, 0, RSB_ST>>>(x, seg, x_iter, V);
  return check_launch("ball_prepare_kernel");
}

extern "C" int rsb_ball_remove(float* x_iter, const uint8_t* mask, long long V, void* stream) {
  RSB_REQUIRE(x_iter && mask && V > 0, "ball_remove: bad arguments");
  ball_remove_kernel<<<grid_for(V, 256), 256, 0, RSB_ST>>>(x_iter, mask, V);
  return check_launch("ball_remove_kernel");
}

extern "C" size_t rsb_ball_workspace_bytes(int D, int H, int W) {
  const size_t cells = static_cast<size_t>((D + 7) / 8) * ((H + 7) / 8) * ((W + 7) / 8);
  return (cells + 255) / 256 * 256 + 256;  // occupancy grid + the packed argmax key
}

extern "C" int rsb_ball_correlate_argmax(const float* x_iter, const void* taps, int n_taps, int kernel_half, void* workspace,
                                         long long* argmax_out, int D, int H, int W, void* stream) {
  RSB_REQUIRE(x_iter && taps && workspace && argmax_out && n_taps > 0 && D > 0 && H > 0 && W > 0, "ball_correlate: bad arguments");
  const int cd = (D + 7) / 8, ch = (H + 7) / 8, cw = (W + 7) / 8;
  const int cells = cd * ch * cw;
  uint8_t* occ = reinterpret_cast<uint8_t*>(workspace);
  unsigned long long* best = reinterpret_cast<unsigned long long*>(occ + (static_cast<size_t>(cells) + 255) / 256 * 256);
  cudaError_t e = cudaMemsetAsync(best, 0, 8, RSB_ST);
  RSB_REQUIRE(e == cudaSuccess, "ball_correlate: memset failed");
  ball_occupancy_kernel<<<cells, 128, 0, RSB_ST>>>(x_iter, occ, D, H, W, cd, ch, cw);
  int rc = check_launch("ball_occupancy_kernel");
  if (rc) return rc;
  const int reach = (kernel_half + 7) / 8;  // cells a tile's taps can reach beyond its own cell
  ball_correlate_kernel<<<cells, 512, 0, RSB_ST>>>(x_iter, occ, reinterpret_cast<const int4*>(taps), n_taps, reach, best, D, H, W, cd, ch, cw);
  rc = check_launch("ball_correlate_kernel");
  if (rc) return rc;
  // unpack on device into the caller's int64: index = 0xFFFFFFFF - low word
  e = cudaMemcpyAsync(argmax_out, best, 8, cudaMemcpyDeviceToDevice, RSB_ST);
  RSB_REQUIRE(e == cudaSuccess, "ball_correlate: copy failed");
  return 0;
}

extern "C" size_t rsb_ball_sep_workspace_bytes(int D, int H, int W, int R) {
  const size_t V = static_cast<size_t>(D) * H * W, rows = static_cast<size_t>(D) * H;
  const size_t occ = (rows + 255) / 256 * 256;
  return 2 * static_cast<size_t>(R + 1) * V * sizeof(float) + 2 * occ + 256;
}

extern "C" int rsb_ball_correlate_argmax_sep(const float* x_iter, const float* gauss, const int* wtab, const float* gauss_host,
                                             const int* wtab_host, int R, void* workspace, long long* argmax_out, int D, int H, int W,
                                             void* stream) {
  RSB_REQUIRE(x_iter && gauss && wtab && workspace && argmax_out && D > 0 && H > 0 && W > 0, "ball_correlate_sep: bad arguments");
  RSB_REQUIRE(R >= 0 && R <= kSepMaxR, "ball_correlate_sep: reach %d outside [0, %d]", R, kSepMaxR);
  const long long V = static_cast<long long>(D) * H * W, rows = static_cast<long long>(D) * H;
  RSB_REQUIRE(V < (1LL << 32) && rows < (1LL << 31), "ball_correlate_sep: volume too large");
  const size_t occ_bytes = (static_cast<size_t>(rows) + 255) / 256 * 256;
  float* Rw = reinterpret_cast<float*>(workspace);
  float* Da = Rw + static_cast<size_t>(R + 1) * V;
  uint8_t* rowocc = reinterpret_cast<uint8_t*>(Da + static_cast<size_t>(R + 1) * V);
  uint8_t* docc = rowocc + occ_bytes;
  unsigned long long* best = reinterpret_cast<unsigned long long*>(docc + occ_bytes);
  cudaError_t e = cudaMemsetAsync(best, 0, 8, RSB_ST);
  RSB_REQUIRE(e == cudaSuccess, "ball_correlate_sep: memset failed");
  ball_row_occupancy_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, RSB_ST>>>(x_iter, rowocc, W, rows);
  int rc = check_launch("ball_row_occupancy_kernel");
  if (rc) return rc;
  ball_rows_kernel<<<static_cast<unsigned>(rows), 128, (W + 2 * R) * sizeof(float), RSB_ST>>>(x_iter, rowocc, gauss, Rw, R, W, V);
  rc = check_launch("ball_rows_kernel");
  if (rc) return rc;
  const size_t tile_bytes = static_cast<size_t>(R + 1) * (kDiscTY + 2 * R) * kDiscTX * sizeof(float);
  DiscTab tab_storage;
  const DiscTab* tab_host = nullptr;
  if (R <= kDiscMaxR && gauss_host != nullptr && wtab_host != nullptr) {
    for (int a = 0; a <= kDiscMaxR; ++a) {
      tab_storage.g[a] = a <= R ? gauss_host[a] : 0.f;
      for (int b = 0; b <= kDiscMaxR; ++b) tab_storage.w[a][b] = (a <= R && b <= R) ? static_cast<signed char>(wtab_host[a * (R + 1) + b]) : -1;
    }
    tab_host = &tab_storage;
  }
  if (tab_host != nullptr && tile_bytes <= 200 * 1024 && D <= 65535) {
    cudaError_t ea = cudaFuncSetAttribute(ball_discs_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(tile_bytes));
    RSB_REQUIRE(ea == cudaSuccess, "ball_correlate_sep: shared-memory opt-in failed: %s", cudaGetErrorString(ea));
    dim3 gridt((W + kDiscTX - 1) / kDiscTX, (H + kDiscTY - 1) / kDiscTY, D);
    ball_discs_tiled_kernel<<<gridt, 256, tile_bytes, RSB_ST>>>(Rw, rowocc, *tab_host, Da, docc, R, H, W, V);
  } else {
    ball_discs_kernel<<<static_cast<unsigned>(rows), 128, 0, RSB_ST>>>(Rw, rowocc, gauss, wtab, Da, docc, R, H, W, V);
  }
  rc = check_launch("ball_discs_kernel");
  if (rc) return rc;
  ball_planes_argmax_kernel<<<static_cast<unsigned>(rows), 128, 0, RSB_ST>>>(Da, docc, gauss, best, R, D, H, W, V);
  rc = check_launch("ball_planes_argmax_kernel");
  if (rc) return rc;
  e = cudaMemcpyAsync(argmax_out, best, 8, cudaMemcpyDeviceToDevice, RSB_ST);
  RSB_REQUIRE(e == cudaSuccess, "ball_correlate_sep: copy failed");
  return 0;
}

extern "C" int rsb_ball_candidates(const float* x, const uint8_t* mask, int mode, int cz, int cy, int cx, int half, float radius2,
                                   void* cand, int* n_cand, int max_cand, uint8_t* ball_out, int D, int H, int W, void* stream) {
  RSB_REQUIRE(x && cand && n_cand && max_cand > 0 && D > 0 && H > 0 && W > 0 && (mode == 0 || mask), "ball_candidates: bad arguments");
  cudaError_t e = cudaMemsetAsync(n_cand, 0, sizeof(int), RSB_ST);
  RSB_REQUIRE(e == cudaSuccess, "ball_candidates: memset failed");
  const long long V = static_cast<long long>(D) * H * W;
  ball_candidates_kernel<<<grid_for(V, 256), 256, 0, RSB_ST>>>(x, mask, mode, cz, cy, cx, half, radius2, reinterpret_cast<uint2*>(cand), n_cand,
                                                                 max_cand, ball_out, D, H, W);
  return check_launch("ball_candidates_kernel");
}

extern "C" int rsb_ball_rank_select(const void* cand, const int* n_cand, int max_cand, int k0, int k1, int k2, uint8_t* m0, uint8_t* m1,
                                    uint8_t* m2, void* stream) {
  RSB_REQUIRE(cand && n_cand && m0 && m1 && m2 && max_cand > 0, "ball_rank_select: bad arguments");
  ball_rank_select_kernel<<<(max_cand + 255) / 256, 256, 0, RSB_ST>>>(reinterpret_cast<const uint2*>(cand), n_cand, max_cand, k0, k1, k2, m0, m1, m2);
  return check_launch("ball_rank_select_kernel");
}

extern "C" int rsb_ball_rank_gwrp(const void* cand, const int* n_cand, int max_cand, float concentration, float* wmap, void* stream) {
  RSB_REQUIRE(cand && n_cand && wmap && max_cand > 0, "ball_rank_gwrp: bad arguments");
  ball_rank_gwrp_kernel<<<(max_cand + 255) / 256, 256, 0, RSB_ST>>>(reinterpret_cast<const uint2*>(cand), n_cand, max_cand, concentration, wmap);
  return check_launch("ball_rank_gwrp_kernel");
}

extern "C" int rsb_ball_weight_map(float* wmap, const uint8_t* pseudo, const uint8_t* dilated, long long V, void* stream) {
  RSB_REQUIRE(wmap && pseudo && dilated && V > 0, "ball_weight_map: bad arguments");
  ball_weight_map_kernel<<<grid_for(V, 256), 256, 0, RSB_ST>>>(wmap, pseudo, dilated, V);
  return check_launch("ball_weight_map_kernel");
}
