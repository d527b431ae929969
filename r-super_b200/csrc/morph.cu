// morph.cu — binary dilation by a ball, bit-exact with the reference's conv-based dilation:
//   dilate_volume / dilate_volume_conv / create_ball_kernel
//   rsuper_train/training/losses_foundation.py:22-99, 1161-1232
// The reference runs a depthwise fp32 F.conv3d with a 0/1 ball kernel and thresholds (> 0); on 0/1
// inputs that is exactly "OR over the ball offsets with zero padding", which is what this kernel
// computes on uint8 volumes.  The pass schedule (k <= 7: one pass; else floor(r/3) passes of k=7
// plus a remainder pass) and the structuring elements (grid = odd(ceil(1.2*odd(ceil(d)))),
// radius = odd(ceil(d))/2) follow the reference exactly, so results are identical.
#include "rsb_common.cuh"

#include <cmath>

#include "../../include/rsuper_b200.h"

namespace rsb {

constexpr int kMaxBallOffsets = 256;
struct BallOffsets {
  int count;
  signed char off[kMaxBallOffsets][3];
};

// create_ball_kernel(diameter, gaussian=False) -> list of (dz,dy,dx) with kernel value 1
static int make_ball(int diameter, BallOffsets& b) {
  int d_odd = diameter;
  if (d_odd % 2 == 0) d_odd += 1;
  int ks = static_cast<int>(std::ceil(1.2 * static_cast<double>(d_odd)));
  if (ks % 2 == 0) ks += 1;
  const double radius = d_odd / 2.0;
  const double center = (ks - 1) / 2.0;
  b.count = 0;
  for (int z = 0; z < ks; ++z)
    for (int y = 0; y < ks; ++y)
      for (int x = 0; x < ks; ++x) {
        // the reference evaluates this in float32 (torch.arange(dtype=float32)); all terms are small
        // integers or half-integers, exact in both precisions
        const float fz = static_cast<float>(z) - static_cast<float>(center);
        const float fy = static_cast<float>(y) - static_cast<float>(center);
        const float fx = static_cast<float>(x) - static_cast<float>(center);
        const float d2 = fz * fz + fy * fy + fx * fx;
        if (d2 <= static_cast<float>(radius * radius)) {
          if (b.count >= kMaxBallOffsets) return -1;
          b.off[b.count][0] = static_cast<signed char>(z - static_cast<int>(center));
          b.off[b.count][1] = static_cast<signed char>(y - static_cast<int>(center));
          b.off[b.count][2] = static_cast<signed char>(x - static_cast<int>(center));
          ++b.count;
        }
      }
  return 0;
}

// conv3d (cross-correlation) with a symmetric ball == OR over offsets; zero padding.
//
// Bit-sliced tile kernel.  A block owns an 8 (z) x 8 (y) x 32 (x) output tile.  Every haloed input row (38 voxels: reach
// <= 3 on both sides) is packed into one 64-bit word with two warp ballots, then spread along x for the four possible half
// widths (S_w[row] = OR_{|dx| <= w} row << dx).  A ball is, for every (dz, dy), the x range [-w(dz,dy), w(dz,dy)], so an
// output row is the OR of <= 49 pre-spread words — ~1.5 instructions per voxel instead of up to 123 byte loads.  A tile
// whose haloed input is empty (most of a lesion channel, all of an unused one) writes zeros without touching the table.
constexpr int kDilReach = 3;                 // k <= 7  =>  |offset| <= 3
constexpr int kDilTZ = 8, kDilTY = 8, kDilTX = 32;
constexpr int kDilRows = (kDilTZ + 2 * kDilReach) * (kDilTY + 2 * kDilReach);   // 196 haloed rows
struct BallRows {
  signed char w[2 * kDilReach + 1][2 * kDilReach + 1];   // half width of the x range at (dz, dy), -1 = empty
};

__global__ void __launch_bounds__(256) dilate_tile_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, const BallRows b,
                                                          int D, int H, int W, int tiles_x, int tiles_y) {
  __shared__ unsigned long long S[kDilReach + 1][kDilRows];
  __shared__ unsigned int out_bits[kDilTZ * kDilTY];
  const long long V = static_cast<long long>(D) * H * W;
  const uint8_t* s = src + blockIdx.y * V;
  uint8_t* d = dst + blockIdx.y * V;
  int t = blockIdx.x;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y; t /= tiles_y;
  const int z0 = t * kDilTZ, y0 = ty * kDilTY, x0 = tx * kDilTX;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int HY = kDilTY + 2 * kDilReach;
  int any = 0;
  for (int r = warp; r < kDilRows; r += 8) {
    const int z = z0 - kDilReach + r / HY, y = y0 - kDilReach + r % HY;
    unsigned lo = 0u, hi = 0u;
    if (z >= 0 && z < D && y >= 0 && y < H) {          // warp-uniform
      const uint8_t* row = s + (static_cast<long long>(z) * H + y) * W;
      const int xa = x0 - kDilReach + lane, xb = xa + 32;
      lo = __ballot_sync(0xffffffffu, xa >= 0 && xa < W && row[xa] != 0);
      hi = __ballot_sync(0xffffffffu, lane < 2 * kDilReach && xb < W && row[xb] != 0);
    }
    if (lane == 0) {
      const unsigned long long w0 = static_cast<unsigned long long>(lo) | (static_cast<unsigned long long>(hi) << 32);
      unsigned long long acc = w0;
      S[0][r] = acc;
#pragma unroll
      for (int w = 1; w <= kDilReach; ++w) {
        acc |= (w0 << w) | (w0 >> w);
        S[w][r] = acc;
      }
    }
    any |= (lo | hi) != 0u;
  }
  any = __syncthreads_or(any);
  if (threadIdx.x < kDilTZ * kDilTY) {
    unsigned long long acc = 0ull;
    if (any) {
      const int oz = threadIdx.x / kDilTY, oy = threadIdx.x % kDilTY;
#pragma unroll
      for (int dz = 0; dz <= 2 * kDilReach; ++dz)
#pragma unroll
        for (int dy = 0; dy <= 2 * kDilReach; ++dy) {
          const int w = b.w[dz][dy];
          if (w >= 0) acc |= S[w][(oz + dz) * HY + oy + dy];
        }
    }
    out_bits[threadIdx.x] = static_cast<unsigned int>(acc >> kDilReach);   // bit i = output voxel x0 + i
  }
  __syncthreads();
  // 256 threads x 8 voxels: thread -> (row, 8-voxel group)
  const int r = threadIdx.x >> 2, g = threadIdx.x & 3;
  const int z = z0 + r / kDilTY, y = y0 + r % kDilTY, x = x0 + g * 8;
  if (z < D && y < H && x < W) {
    const unsigned bits = (out_bits[r] >> (g * 8)) & 0xFFu;
    uint8_t* o = d + (static_cast<long long>(z) * H + y) * W + x;
    if (x + 8 <= W && (reinterpret_cast<uintptr_t>(o) & 7) == 0) {
      // bit j -> byte j (0/1): multiply spreads the 8 bits, the mask keeps bit j of byte j, the second step normalises to 1
      unsigned long long v = (static_cast<unsigned long long>(bits) * 0x0101010101010101ull) & 0x8040201008040201ull;
      v = ((v + 0x7F7F7F7F7F7F7F7Full) >> 7) & 0x0101010101010101ull;
      *reinterpret_cast<unsigned long long*>(o) = v;
    } else {
      for (int j = 0; j < 8 && x + j < W; ++j) o[j] = (bits >> j) & 1u;
    }
  }
}

}  // namespace rsb

using namespace rsb;

extern "C" int rsb_dilate_ball(const uint8_t* src, uint8_t* dst, uint8_t* tmp, int n_vol, int D, int H, int W,
                               int kernel_size, void* stream) {
  RSB_REQUIRE(src && dst && tmp, "dilate_ball: null pointer");
  RSB_REQUIRE(n_vol > 0 && n_vol <= 65535 && D > 0 && H > 0 && W > 0, "dilate_ball: bad geometry");
  RSB_REQUIRE(kernel_size >= 1, "dilate_ball: bad kernel size %d", kernel_size);
  // dilate_volume (losses_foundation.py:22-46), full_pass_radius = 3
  if (kernel_size % 2 == 0) kernel_size += 1;
  int passes[64];
  int np = 0;
  if (kernel_size <= 7) {
    passes[np++] = kernel_size;
  } else {
    const int radius = (kernel_size - 1) / 2;
    const int num_full = radius / 3, rem = radius % 3;
    RSB_REQUIRE(num_full + 1 <= 64, "dilate_ball: kernel too large");
    for (int i = 0; i < num_full; ++i) passes[np++] = 7;
    if (rem > 0) passes[np++] = 2 * rem + 1;
  }
  const int tiles_x = (W + kDilTX - 1) / kDilTX, tiles_y = (H + kDilTY - 1) / kDilTY, tiles_z = (D + kDilTZ - 1) / kDilTZ;
  const long long tiles = static_cast<long long>(tiles_x) * tiles_y * tiles_z;
  RSB_REQUIRE(tiles < (1LL << 31), "dilate_ball: volume too large");
  dim3 grid(static_cast<unsigned>(tiles), n_vol);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // ping-pong so that the last pass lands in dst
  const uint8_t* cur = src;
  for (int i = 0; i < np; ++i) {
    BallOffsets b;
    RSB_REQUIRE(make_ball(passes[i], b) == 0, "dilate_ball: structuring element too large");
    // the ball as x ranges: for every (dz, dy) the offsets form the symmetric run [-w, w] (checked)
    BallRows rows;
    int cnt[2 * kDilReach + 1][2 * kDilReach + 1] = {};
    for (int z = 0; z <= 2 * kDilReach; ++z)
      for (int y = 0; y <= 2 * kDilReach; ++y) rows.w[z][y] = -1;
    for (int k = 0; k < b.count; ++k) {
      const int dz = b.off[k][0], dy = b.off[k][1], dx = b.off[k][2];
      RSB_REQUIRE(abs(dz) <= kDilReach && abs(dy) <= kDilReach && abs(dx) <= kDilReach, "dilate_ball: pass reach exceeds %d", kDilReach);
      signed char& w = rows.w[dz + kDilReach][dy + kDilReach];
      if (abs(dx) > w) w = static_cast<signed char>(abs(dx));
      ++cnt[dz + kDilReach][dy + kDilReach];
    }
    for (int z = 0; z <= 2 * kDilReach; ++z)
      for (int y = 0; y <= 2 * kDilReach; ++y)
        RSB_REQUIRE(rows.w[z][y] < 0 || cnt[z][y] == 2 * rows.w[z][y] + 1, "dilate_ball: structuring element is not a union of x runs");
    uint8_t* out = ((np - 1 - i) % 2 == 0) ? dst : tmp;
    dilate_tile_kernel<<<grid, 256, 0, st>>>(cur, out, rows, D, H, W, tiles_x, tiles_y);
    int rc = check_launch("dilate_tile_kernel");
    if (rc) return rc;
    cur = out;
  }
  return 0;
}
