// train_glue.cu — the optimizer end of the train step as two multi-tensor launches.
//
// Replaces, for every parameter tensor of the network at once,
//   torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)            rsuper_train/train_ddp.py:352
//   torch.optim.AdamW(lr, betas, eps=1e-5, weight_decay).step()       rsuper_train/training/utils.py:46-51, train_ddp.py:353
//   update_ema_variables: ema = a*ema + (1-a)*param                   rsuper_train/training/utils.py:154-158
// which stock PyTorch runs as ~10 multi-tensor launches making ~9 passes over the 40.56 M parameters.
//
//   launch 1  grad_sqnorm_kernel    one read of every gradient -> one partial sum of squares per block (no atomics:
//                                   the order of the additions is fixed, the norm is deterministic)
//   launch 2  clip_adamw_ema_kernel every block re-sums the block partials in the same order (<= 1184 floats from L2),
//                                   derives clip = min(1, max_norm / (norm + 1e-6)) and updates g, p, m, v, ema in ONE
//                                   pass: 5 reads + 5 writes of 4 bytes per parameter (HBM bound: 40 B/param).
//
// Tensors are described by a device table (RsbOptTensor); a block processes chunks of CHUNK elements and finds the
// tensor of a chunk by binary search over the chunk prefix.  128-bit accesses when all five pointers of a tensor are
// 16-byte aligned (DDP bucket views need not be), scalar otherwise and on the ragged tail.
#include "rsb_common.cuh"

#include <cmath>
#include <cstring>

#include "../../include/rsuper_b200.h"

namespace rsb {

constexpr int OPT_THREADS = 256;
constexpr int OPT_CHUNK = 4096;  // elements per (block, iteration): 4 float4 per thread

struct OptHyper {
  float clip_max_norm;  // <= 0: no clipping (clip = 1)
  float decay;          // 1 - lr * weight_decay
  float one_minus_b1, b2, one_minus_b2;
  float step_size;      // lr / (1 - b1^t)
  float bc2_sqrt;       // sqrt(1 - b2^t)
  float eps;
  float ema_alpha, one_minus_ema_alpha;
};

RSB_DEVICE int find_tensor(const RsbOptTensor* __restrict__ tab, int n_tensors, long long chunk) {
  int lo = 0, hi = n_tensors - 1;  // last tensor whose chunk_begin <= chunk
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab[mid].chunk_begin <= chunk) lo = mid; else hi = mid - 1;
  }
  return lo;
}

RSB_DEVICE float block_sum(float v, float* red /* OPT_THREADS / 32 floats */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // red may still be read by a previous call
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < OPT_THREADS / 32; ++w) t += red[w];  // same order in every thread of every block
  return t;
}

__global__ void __launch_bounds__(OPT_THREADS) grad_sqnorm_kernel(const RsbOptTensor* __restrict__ tab, int n_tensors,
                                                                   long long total_chunks, float* __restrict__ partials) {
  __shared__ float red[OPT_THREADS / 32];
  float acc = 0.f;
  for (long long chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
    const RsbOptTensor t = tab[find_tensor(tab, n_tensors, chunk)];
    if (t.g == nullptr) continue;  // EMA-only row (parameter without a gradient this step)
    const long long begin = (chunk - t.chunk_begin) * OPT_CHUNK;
    const long long rem = t.n - begin;
    const int cnt = rem < OPT_CHUNK ? static_cast<int>(rem) : OPT_CHUNK;
    const float* g = t.g + begin;
    if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
      const int nv = cnt >> 2;
      const float4* g4 = reinterpret_cast<const float4*>(g);
      for (int i = threadIdx.x; i < nv; i += OPT_THREADS) {
        const float4 q = __ldg(g4 + i);
        acc += q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
      }
      for (int i = (nv << 2) + threadIdx.x; i < cnt; i += OPT_THREADS) acc += g[i] * g[i];
    } else {
      for (int i = threadIdx.x; i < cnt; i += OPT_THREADS) acc += g[i] * g[i];
    }
  }
  const float s = block_sum(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

RSB_DEVICE void update_one(float& g, float& p, float& m, float& v, float& e, float clip, const OptHyper& h) {
  g *= clip;                                       // clip_grad_norm_: g.mul_(clip_coef_clamped)
  p *= h.decay;                                    // AdamW: param.mul_(1 - lr * weight_decay)
  m += (g - m) * h.one_minus_b1;                   // exp_avg.lerp_(grad, 1 - beta1)
  v = h.b2 * v + h.one_minus_b2 * g * g;           // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
  const float denom = sqrtf(v) / h.bc2_sqrt + h.eps;     // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
  p -= h.step_size * (m / denom);                  // param.addcdiv_(exp_avg, denom, value = -step_size)
  e = h.ema_alpha * e + h.one_minus_ema_alpha * p; // ema.mul_(alpha).add_(param, alpha = 1 - alpha)
}

template <bool HasEma>
__global__ void __launch_bounds__(OPT_THREADS) clip_adamw_ema_kernel(const RsbOptTensor* __restrict__ tab, int n_tensors,
                                                                      long long total_chunks, const float* __restrict__ partials,
                                                                      int n_partials, OptHyper h_value, const OptHyper* __restrict__ h_device,
                                                                      float* __restrict__ norm_out) {
  __shared__ float red[OPT_THREADS / 32];
  // step-dependent scalars either by value (eager launches) or from device memory (CUDA-graph replays: the host rewrites
  // that struct before every replay, the captured launch itself never changes)
  const OptHyper h = h_device != nullptr ? *h_device : h_value;
  float clip = 1.f;
  {
    float acc = 0.f;
    if (partials != nullptr)
      for (int i = threadIdx.x; i < n_partials; i += OPT_THREADS) acc += partials[i];
    const float total = block_sum(acc, red);
    const float norm = sqrtf(total);
    if (h.clip_max_norm > 0.f) clip = fminf(h.clip_max_norm / (norm + 1e-6f), 1.f);
    if (norm != norm) clip = norm;  // NaN gradients propagate like in clip_grad_norm_ (fminf would drop the NaN)
    if (blockIdx.x == 0 && threadIdx.x == 0 && norm_out != nullptr) *norm_out = norm;
  }
  for (long long chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
    const RsbOptTensor t = tab[find_tensor(tab, n_tensors, chunk)];
    const long long begin = (chunk - t.chunk_begin) * OPT_CHUNK;
    const long long rem = t.n - begin;
    const int cnt = rem < OPT_CHUNK ? static_cast<int>(rem) : OPT_CHUNK;
    if (t.g == nullptr) {
      // no gradient this step: torch's AdamW skips the parameter, update_ema_variables (training/utils.py:154-158) does not
      if (HasEma && t.ema != nullptr) {
        const float* p = t.p + begin;
        float* e = t.ema + begin;
        for (int i = threadIdx.x; i < cnt; i += OPT_THREADS) e[i] = h.ema_alpha * e[i] + h.one_minus_ema_alpha * p[i];
      }
      continue;
    }
    float* g = t.g + begin;
    float* p = t.p + begin;
    float* m = t.m + begin;
    float* v = t.v + begin;
    float* e = HasEma ? t.ema + begin : nullptr;
    // the table holds generic pointers; telling the compiler they are global memory turns LD/ST.E into LDG/STG
    __builtin_assume(__isGlobal(g));
    __builtin_assume(__isGlobal(p));
    __builtin_assume(__isGlobal(m));
    __builtin_assume(__isGlobal(v));
    if (HasEma) __builtin_assume(__isGlobal(e));
    uintptr_t align = reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(m) |
                      reinterpret_cast<uintptr_t>(v);
    if (HasEma) align |= reinterpret_cast<uintptr_t>(e);
    int done = 0;
    if ((align & 15) == 0) {
      const int nv = cnt >> 2;
      for (int i = threadIdx.x; i < nv; i += OPT_THREADS) {
        float4 G = reinterpret_cast<float4*>(g)[i], P = reinterpret_cast<float4*>(p)[i], M = reinterpret_cast<float4*>(m)[i],
               V = reinterpret_cast<float4*>(v)[i];
        float4 E = HasEma ? reinterpret_cast<float4*>(e)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        update_one(G.x, P.x, M.x, V.x, E.x, clip, h);
        update_one(G.y, P.y, M.y, V.y, E.y, clip, h);
        update_one(G.z, P.z, M.z, V.z, E.z, clip, h);
        update_one(G.w, P.w, M.w, V.w, E.w, clip, h);
        reinterpret_cast<float4*>(g)[i] = G;
        reinterpret_cast<float4*>(p)[i] = P;
        reinterpret_cast<float4*>(m)[i] = M;
        reinterpret_cast<float4*>(v)[i] = V;
        if (HasEma) reinterpret_cast<float4*>(e)[i] = E;
      }
      done = nv << 2;
    }
    for (int i = done + threadIdx.x; i < cnt; i += OPT_THREADS) {
      float G = g[i], P = p[i], M = m[i], V = v[i], E = HasEma ? e[i] : 0.f;
      update_one(G, P, M, V, E, clip, h);
      g[i] = G; p[i] = P; m[i] = M; v[i] = V;
      if (HasEma) e[i] = E;
    }
  }
}

}  // namespace rsb

using namespace rsb;

// python-double scalar arithmetic like torch's single-tensor AdamW (optim/adamw.py), rounded to fp32 once
static OptHyper make_hyper(double max_norm, double lr, double beta1, double beta2, double eps, double weight_decay, long long step,
                           double ema_alpha) {
  OptHyper h;
  h.clip_max_norm = static_cast<float>(max_norm);
  h.decay = static_cast<float>(1.0 - lr * weight_decay);
  h.one_minus_b1 = static_cast<float>(1.0 - beta1);
  h.b2 = static_cast<float>(beta2);
  h.one_minus_b2 = static_cast<float>(1.0 - beta2);
  const double bc1 = 1.0 - pow(beta1, static_cast<double>(step));
  const double bc2 = 1.0 - pow(beta2, static_cast<double>(step));
  h.step_size = static_cast<float>(lr / bc1);
  h.bc2_sqrt = static_cast<float>(sqrt(bc2));
  h.eps = static_cast<float>(eps);
  h.ema_alpha = static_cast<float>(ema_alpha);
  h.one_minus_ema_alpha = static_cast<float>(1.0 - ema_alpha);
  return h;
}

extern "C" long long rsb_opt_chunk_elems(void) { return OPT_CHUNK; }

extern "C" int rsb_opt_max_blocks(void) {
  const int sms = rsb_num_sms();
  return (sms > 0 ? sms : 148) * 8;
}

extern "C" int rsb_clip_adamw_ema_step(const RsbOptTensor* table_device, int n_tensors, long long total_chunks, int has_ema,
                                       float* partials, float* norm_out, double max_norm, double lr, double beta1, double beta2,
                                       double eps, double weight_decay, long long step, double ema_alpha, const float* hyper_device,
                                       void* stream) {
  RSB_REQUIRE(table_device != nullptr && n_tensors > 0 && total_chunks > 0, "clip_adamw_ema_step: empty tensor table");
  RSB_REQUIRE(step >= 1, "clip_adamw_ema_step: step counts from 1 (got %lld)", step);
  RSB_REQUIRE(max_norm <= 0.0 || partials != nullptr, "clip_adamw_ema_step: clipping needs the partials workspace");
  RSB_REQUIRE(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0, "clip_adamw_ema_step: betas must be in [0, 1)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int max_blocks = rsb_opt_max_blocks();
  const int grid = static_cast<int>(total_chunks < max_blocks ? total_chunks : max_blocks);
  if (partials != nullptr) {
    grad_sqnorm_kernel<<<grid, OPT_THREADS, 0, st>>>(table_device, n_tensors, total_chunks, partials);
    if (int rc = check_launch("grad_sqnorm_kernel")) return rc;
  }
  const OptHyper h = make_hyper(max_norm, lr, beta1, beta2, eps, weight_decay, step, ema_alpha);
  if (has_ema)
    clip_adamw_ema_kernel<true><<<grid, OPT_THREADS, 0, st>>>(table_device, n_tensors, total_chunks, partials, grid, h, reinterpret_cast<const OptHyper*>(hyper_device), norm_out);
  else
    clip_adamw_ema_kernel<false><<<grid, OPT_THREADS, 0, st>>>(table_device, n_tensors, total_chunks, partials, grid, h, reinterpret_cast<const OptHyper*>(hyper_device), norm_out);
  return check_launch("clip_adamw_ema_kernel");
}

extern "C" int rsb_opt_hyper_floats(void) { return static_cast<int>(sizeof(OptHyper) / sizeof(float)); }

extern "C" int rsb_opt_fill_hyper(float* hyper_host, double max_norm, double lr, double beta1, double beta2, double eps,
                                  double weight_decay, long long step, double ema_alpha) {
  RSB_REQUIRE(hyper_host != nullptr && step >= 1, "opt_fill_hyper: null pointer / step < 1");
  const OptHyper h = make_hyper(max_norm, lr, beta1, beta2, eps, weight_decay, step, ema_alpha);
  memcpy(hyper_host, &h, sizeof(h));
  return 0;
}
