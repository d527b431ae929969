// seg_loss.cu — fused segmentation loss of calculate_loss:
//   loss = mean(BCEWithLogits(r, l, weight=cw) * known) + DiceLossMultiClass(r, l, known, cw)
//   rsuper_train/training/losses_foundation.py:945-956 (deep supervision) / 1030-1035, 541-607.
// The reference materialises >= 16 full [B,C,V] fp32 temporaries for this; here the forward is ONE
// read of (logits, label, known) producing 4 partial sums per (b,c), a one-block finalize that also
// derives the backward coefficients (including the gradient that flows through the batch-coupled,
// clamped alpha_c — losses_foundation.py:581-584), and the backward is one more read + one write.
//   bytes per voxel*class: fwd 4+1+1, bwd 4+1+1+4  => 16 B.
#include "rsb_common.cuh"

#include "../../include/rsuper_b200.h"

namespace rsb {

constexpr int kSegThreads = 256;
constexpr float kSmooth = 1e-5f;  // losses_foundation.py:569

RSB_DEVICE float sigmoidf_acc(float x) { return 1.f / (1.f + expf(-x)); }
// max(x,0) - x*l + log1p(exp(-|x|))   (ATen binary_cross_entropy_with_logits formulation)
RSB_DEVICE float bce_logits(float x, float l) { return fmaxf(x, 0.f) - x * l + log1pf(expf(-fabsf(x))); }

RSB_DEVICE float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = 0.f;
  if (warp == 0) {
    r = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
    r = warp_sum(r);
  }
  return r;  // valid in warp 0
}

// grid = (blocks, B*C).  partials[bc*4 + {0:bce, 1:TP, 2:FP, 3:FN}]
__global__ void __launch_bounds__(kSegThreads) seg_loss_pass1_kernel(const float* __restrict__ logits,
                                                                     const uint8_t* __restrict__ label,
                                                                     const uint8_t* __restrict__ known,
                                                                     const float* __restrict__ wmap,
                                                                     float* __restrict__ partials, long long V) {
  __shared__ float red[8];
  const int bc = blockIdx.y;
  const float* r = logits + static_cast<long long>(bc) * V;
  const uint8_t* l = label + static_cast<long long>(bc) * V;
  const uint8_t* k = known ? known + static_cast<long long>(bc) * V : nullptr;
  const float* wm = wmap ? wmap + static_cast<long long>(bc) * V : nullptr;
  float s_bce = 0.f, s_tp = 0.f, s_fp = 0.f, s_fn = 0.f;
  const long long nvec = V / 16;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4* rp = reinterpret_cast<const float4*>(r + i * 16);
    float4 q0 = rp[0], q1 = rp[1], q2 = rp[2], q3 = rp[3];
    const uint4 lv = *reinterpret_cast<const uint4*>(l + i * 16);
    uint4 kv = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
    if (k) kv = *reinterpret_cast<const uint4*>(k + i * 16);
    const float xs[16] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w,
                          q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
    const uint32_t lw[4] = {lv.x, lv.y, lv.z, lv.w};
    const uint32_t kw[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float lab = ((lw[j >> 2] >> ((j & 3) * 8)) & 0xFFu) ? 1.f : 0.f;
      const float kn = ((kw[j >> 2] >> ((j & 3) * 8)) & 0xFFu) ? 1.f : 0.f;
      const float x = xs[j];
      const float sg = sigmoidf_acc(x);
      s_bce += kn * bce_logits(x, lab) * (wm ? wm[i * 16 + j] : 1.f);
      s_tp += kn * lab * sg;
      s_fp += kn * (1.f - lab) * sg;
      s_fn += kn * lab * (1.f - sg);
    }
  }
  // scalar tail (V % 16)
  if (blockIdx.x == 0) {
    for (long long i = nvec * 16 + threadIdx.x; i < V; i += blockDim.x) {
      const float lab = l[i] ? 1.f : 0.f;
      const float kn = k ? (k[i] ? 1.f : 0.f) : 1.f;
      const float x = r[i];
      const float sg = sigmoidf_acc(x);
      s_bce += kn * bce_logits(x, lab) * (wm ? wm[i] : 1.f);
      s_tp += kn * lab * sg;
      s_fp += kn * (1.f - lab) * sg;
      s_fn += kn * lab * (1.f - sg);
    }
  }
  const float t0 = block_sum(s_bce, red);
  const float t1 = block_sum(s_tp, red);
  const float t2 = block_sum(s_fp, red);
  const float t3 = block_sum(s_fn, red);
  if (threadIdx.x == 0) {
    atomicAdd(&partials[bc * 4 + 0], t0);
    atomicAdd(&partials[bc * 4 + 1], t1);
    atomicAdd(&partials[bc * 4 + 2], t2);
    atomicAdd(&partials[bc * 4 + 3], t3);
  }
}

// one block; loss_out = {total, bce, dice}; coef[bc*4] = {c0, c1, c2, alpha}
//   dL/dr = c0*k*(sig-l) + k*sig*(1-sig) * (l ? c1 : c2)
__global__ void seg_loss_finalize_kernel(const float* __restrict__ partials, const float* __restrict__ cw,
                                         float* __restrict__ coef, float* __restrict__ loss_out, int B, int C,
                                         long long V) {
  __shared__ float red[8];
  __shared__ float s_alpha[256], s_dalpha[256];
  const float inv_bc = 1.f / (static_cast<float>(B) * C);
  const float inv_bcv = inv_bc / static_cast<float>(V);
  // per-class alpha and d(loss)/d(alpha)
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float F = 0.f, G = 0.f;
    for (int b = 0; b < B; ++b) {
      F += partials[(b * C + c) * 4 + 2];
      G += partials[(b * C + c) * 4 + 3];
    }
    const float araw = F / (F + G + kSmooth);
    const float alpha = fminf(fmaxf(araw, 0.2f), 0.8f);
    float dl_da = 0.f;
    for (int b = 0; b < B; ++b) {
      const float w = cw ? cw[b * C + c] : 1.f;
      const float tp = partials[(b * C + c) * 4 + 1], fp = partials[(b * C + c) * 4 + 2],
                  fn = partials[(b * C + c) * 4 + 3];
      const float den = tp + alpha * fp + (1.f - alpha) * fn + kSmooth;
      dl_da += w * inv_bc * tp * (fp - fn) / (den * den);
    }
    const bool pass = (araw >= 0.2f) && (araw <= 0.8f);
    const float s = F + G + kSmooth;
    s_alpha[c] = alpha;
    // d alpha/dF = (G+s0)/s^2 ; d alpha/dG = -F/s^2  (folded below)
    s_dalpha[c] = pass ? dl_da / (s * s) : 0.f;
    // stash F, G scaled terms through coef of b = 0 slots later
  }
  __syncthreads();
  float l_bce = 0.f, l_dice = 0.f;
  for (int i = threadIdx.x; i < B * C; i += blockDim.x) {
    const int c = i % C;
    const float w = cw ? cw[i] : 1.f;
    const float tp = partials[i * 4 + 1], fp = partials[i * 4 + 2], fn = partials[i * 4 + 3];
    const float alpha = s_alpha[c], beta = 1.f - alpha;
    const float den = tp + alpha * fp + beta * fn + kSmooth;
    l_bce += w * partials[i * 4 + 0] * inv_bcv;
    l_dice += w * (1.f - tp / den) * inv_bc;
    float F = 0.f, G = 0.f;
    for (int b = 0; b < B; ++b) {
      F += partials[(b * C + c) * 4 + 2];
      G += partials[(b * C + c) * 4 + 3];
    }
    const float dT = -w * inv_bc * (den - tp) / (den * den);
    const float dFP = w * inv_bc * tp * alpha / (den * den) + s_dalpha[c] * (G + kSmooth);
    const float dFN = w * inv_bc * tp * beta / (den * den) - s_dalpha[c] * F;
    coef[i * 4 + 0] = w * inv_bcv;
    coef[i * 4 + 1] = dT - dFN;  // label == 1
    coef[i * 4 + 2] = dFP;       // label == 0
    coef[i * 4 + 3] = alpha;
  }
  const float tb = block_sum(l_bce, red);
  const float td = block_sum(l_dice, red);
  if (threadIdx.x == 0) {
    loss_out[0] = tb + td;
    loss_out[1] = tb;
    loss_out[2] = td;
  }
}

__global__ void __launch_bounds__(kSegThreads) seg_loss_pass2_kernel(const float* __restrict__ logits,
                                                                     const uint8_t* __restrict__ label,
                                                                     const uint8_t* __restrict__ known,
                                                                     const float* __restrict__ wmap,
                                                                     const float* __restrict__ coef,
                                                                     const float* __restrict__ grad_scale,
                                                                     float* __restrict__ dlogits, int accumulate,
                                                                     long long V) {
  const int bc = blockIdx.y;
  // grad_scale[0] scales the BCE term, grad_scale[1] the Dice term (separate dict entries in the Ball loss)
  const float gb = grad_scale ? grad_scale[0] : 1.f, gd = grad_scale ? grad_scale[1] : 1.f;
  const float c0 = coef[bc * 4 + 0] * gb, c1 = coef[bc * 4 + 1] * gd, c2 = coef[bc * 4 + 2] * gd;
  const float* wm = wmap ? wmap + static_cast<long long>(bc) * V : nullptr;
  const float* r = logits + static_cast<long long>(bc) * V;
  const uint8_t* l = label + static_cast<long long>(bc) * V;
  const uint8_t* k = known ? known + static_cast<long long>(bc) * V : nullptr;
  float* d = dlogits + static_cast<long long>(bc) * V;
  const long long nvec = V / 16;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4* rp = reinterpret_cast<const float4*>(r + i * 16);
    float4 q0 = rp[0], q1 = rp[1], q2 = rp[2], q3 = rp[3];
    const uint4 lv = *reinterpret_cast<const uint4*>(l + i * 16);
    uint4 kv = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
    if (k) kv = *reinterpret_cast<const uint4*>(k + i * 16);
    const float xs[16] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w,
                          q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
    const uint32_t lw[4] = {lv.x, lv.y, lv.z, lv.w};
    const uint32_t kw[4] = {kv.x, kv.y, kv.z, kv.w};
    float o[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const bool lab = ((lw[j >> 2] >> ((j & 3) * 8)) & 0xFFu) != 0;
      const float kn = ((kw[j >> 2] >> ((j & 3) * 8)) & 0xFFu) ? 1.f : 0.f;
      const float sg = sigmoidf_acc(xs[j]);
      o[j] = kn * (c0 * (wm ? wm[i * 16 + j] : 1.f) * (sg - (lab ? 1.f : 0.f)) + sg * (1.f - sg) * (lab ? c1 : c2));
    }
    float4* dp = reinterpret_cast<float4*>(d + i * 16);
    if (accumulate) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 t = dp[q];
        t.x += o[q * 4]; t.y += o[q * 4 + 1]; t.z += o[q * 4 + 2]; t.w += o[q * 4 + 3];
        dp[q] = t;
      }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) dp[q] = make_float4(o[q * 4], o[q * 4 + 1], o[q * 4 + 2], o[q * 4 + 3]);
    }
  }
  if (blockIdx.x == 0) {
    for (long long i = nvec * 16 + threadIdx.x; i < V; i += blockDim.x) {
      const bool lab = l[i] != 0;
      const float kn = k ? (k[i] ? 1.f : 0.f) : 1.f;
      const float sg = sigmoidf_acc(r[i]);
      const float v = kn * (c0 * (wm ? wm[i] : 1.f) * (sg - (lab ? 1.f : 0.f)) + sg * (1.f - sg) * (lab ? c1 : c2));
      d[i] = accumulate ? d[i] + v : v;
    }
  }
}

}  // namespace rsb

using namespace rsb;

static int seg_check(const RsbSegLossArgs* p) {
  RSB_REQUIRE(p != nullptr, "seg_loss: null args");
  RSB_REQUIRE(p->logits && p->label && p->partials && p->coef && p->loss_out, "seg_loss: null pointer");
  RSB_REQUIRE(p->B > 0 && p->C > 0 && p->C <= 256 && p->V > 0, "seg_loss: bad shape (B=%d C=%d)", p->B, p->C);
  RSB_REQUIRE(static_cast<long long>(p->B) * p->C <= 65535, "seg_loss: B*C too large");
  return 0;
}

static int seg_grid_x(long long V) {
  const int sms = rsb_num_sms();
  long long want = (V / 16 + kSegThreads - 1) / kSegThreads;
  long long cap = sms > 0 ? sms * 4LL : 512;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return static_cast<int>(want);
}

extern "C" int rsb_seg_loss_forward(const RsbSegLossArgs* p, void* stream) {
  if (seg_check(p)) return -1;
  // vector path needs 16-byte aligned rows
  RSB_REQUIRE(p->V % 16 == 0, "seg_loss: D*H*W must be a multiple of 16 (got %lld)", p->V);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(p->partials, 0, sizeof(float) * 4 * p->B * p->C, st);
  RSB_REQUIRE(e == cudaSuccess, "seg_loss: memset failed: %s", cudaGetErrorString(e));
  dim3 grid(seg_grid_x(p->V), p->B * p->C);
  seg_loss_pass1_kernel<<<grid, kSegThreads, 0, st>>>(p->logits, p->label, p->known, p->bce_weight_map, p->partials, p->V);
  int rc = check_launch("seg_loss_pass1_kernel");
  if (rc) return rc;
  seg_loss_finalize_kernel<<<1, 256, 0, st>>>(p->partials, p->class_weights, p->coef, p->loss_out, p->B, p->C, p->V);
  return check_launch("seg_loss_finalize_kernel");
}

extern "C" int rsb_seg_loss_backward(const RsbSegLossArgs* p, const float* grad_scale, float* dlogits,
                                     int accumulate, void* stream) {
  if (seg_check(p)) return -1;
  RSB_REQUIRE(dlogits != nullptr, "seg_loss_backward: null dlogits");
  RSB_REQUIRE(p->V % 16 == 0, "seg_loss: D*H*W must be a multiple of 16 (got %lld)", p->V);
  dim3 grid(seg_grid_x(p->V), p->B * p->C);
  seg_loss_pass2_kernel<<<grid, kSegThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      p->logits, p->label, p->known, p->bce_weight_map, p->coef, grad_scale, dlogits, accumulate, p->V);
  return check_launch("seg_loss_pass2_kernel");
}
