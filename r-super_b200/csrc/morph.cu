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
__global__ void dilate_pass_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                   const BallOffsets b, int D, int H, int W) {
  const long long V = static_cast<long long>(D) * H * W;
  const long long vol = blockIdx.y;
  const uint8_t* s = src + vol * V;
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < V;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(v % W), y = static_cast<int>((v / W) % H), z = static_cast<int>(v / (static_cast<long long>(W) * H));
    uint8_t out = 0;
    for (int i = 0; i < b.count && !out; ++i) {
      const int zz = z + b.off[i][0], yy = y + b.off[i][1], xx = x + b.off[i][2];
      if (zz >= 0 && zz < D && yy >= 0 && yy < H && xx >= 0 && xx < W)
        out = s[(static_cast<long long>(zz) * H + yy) * W + xx] ? 1 : 0;
    }
    dst[vol * V + v] = out;
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
  const int sms = rsb_num_sms();
  RSB_REQUIRE(sms > 0, "no CUDA device");
  const long long V = static_cast<long long>(D) * H * W;
  long long bx = (V + 255) / 256;
  if (bx > sms * 8LL) bx = sms * 8LL;
  dim3 grid(static_cast<unsigned>(bx), n_vol);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // ping-pong so that the last pass lands in dst
  const uint8_t* cur = src;
  for (int i = 0; i < np; ++i) {
    BallOffsets b;
    RSB_REQUIRE(make_ball(passes[i], b) == 0, "dilate_ball: structuring element too large");
    uint8_t* out = ((np - 1 - i) % 2 == 0) ? dst : tmp;
    dilate_pass_kernel<<<grid, 256, 0, st>>>(cur, out, b, D, H, W);
    int rc = check_launch("dilate_pass_kernel");
    if (rc) return rc;
    cur = out;
  }
  return 0;
}
