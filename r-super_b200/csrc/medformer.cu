// medformer.cu — the voxel-side kernels MedFormer needs beyond the UNet set.  All of them are HBM / latency
// bound CUDA-core kernels on NDHWC tensors (8 channels = one 128-bit access per thread where the access pattern allows);
// the GEMM-shaped parts of the model (3x3x3 and 1x1x1 convolutions) run on the tcgen05 kernels of conv3_fprop.cu /
// conv3_wgrad.cu.
//
// Reference call sites replaced (rsuper_train/model/dim3):
//   nn.Conv3d(C, C, 3, groups=C) in DepthwiseSeparableConv / MBConv            conv_layers.py:125-157, 192-230
//   SEBlock's `x * excitation(squeeze(x))`                                      conv_layers.py:159-173
//   SemanticMapGeneration: softmax over the voxels + einsum('bij,bkj->bik')     medformer_utils.py:222-235
//   BidirectionAttention: the two softmaxes of one logit matrix and both einsums medformer_utils.py:67-103
#include "rsb_common.cuh"

#include "../../include/rsuper_b200.h"

namespace rsb {

// Thread mapping of the per-channel kernels: blockIdx.y = sample, blockIdx.z = chunk of `cgb` channel groups (8 channels
// each), threadIdx.x = (voxel within the block, channel group within the chunk): a thread keeps its channels over the
// grid-stride loop, consecutive threads read consecutive channels of one voxel.
struct MfMap {
  int cg;
  bool active;
  long long v0, vstride;
};
RSB_DEVICE MfMap mf_map(int CG, int cgb) {
  MfMap m;
  const int vpb = blockDim.x / cgb;
  const int cgl = static_cast<int>(threadIdx.x) % cgb;
  m.cg = static_cast<int>(blockIdx.z) * cgb + cgl;
  m.active = static_cast<int>(threadIdx.x) < vpb * cgb && m.cg < CG;
  m.v0 = static_cast<long long>(blockIdx.x) * vpb + threadIdx.x / cgb;
  m.vstride = static_cast<long long>(gridDim.x) * vpb;
  return m;
}

constexpr int kMfBlock = 256;

RSB_DEVICE float mf_ld(const float* p) { return *p; }
RSB_DEVICE float mf_ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
RSB_DEVICE void mf_st(float* p, float v) { *p = v; }
RSB_DEVICE void mf_st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

struct MfGrid {
  int cgb, chunks, gx;
};
// voxels_per_thread > 1: reduction kernels (every block ends with one round of global atomics per channel and tap — with
// few voxels and many channels one voxel per thread would spend the launch on contended atomics)
static inline MfGrid mf_grid(int C, long long V, int sms, int zmul = 1, int voxels_per_thread = 1) {
  MfGrid g;
  const int CG = C / 8;
  g.cgb = CG < 32 ? CG : 32;
  g.chunks = (CG + g.cgb - 1) / g.cgb;
  const int vpb = kMfBlock / g.cgb;
  long long want = (V + static_cast<long long>(vpb) * voxels_per_thread - 1) / (static_cast<long long>(vpb) * voxels_per_thread);
  long long cap = static_cast<long long>(sms) * 8 / (static_cast<long long>(g.chunks) * zmul);
  if (cap < 1) cap = 1;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  g.gx = static_cast<int>(want);
  return g;
}

// ------------------------------------------------------------------------------------------
// depthwise 3x3x3 convolution (padding 1, no bias): y[n,v,c] = sum_t w[c][t] * a[n, v + t - 1, c]
// flip = 1 computes the data gradient (the same sum with the taps mirrored).
// ------------------------------------------------------------------------------------------
constexpr int kDwRun = 4;

template <typename T>
__global__ void dwconv3_kernel(const T* __restrict__ a, long long ap, const float* __restrict__ w, T* __restrict__ y, long long yp,
                               int D, int H, int W, int C, int cgb, int flip) {
  extern __shared__ float sm_w[];   // [27][cgb * 8]
  const int CG = C / 8;
  const int n = blockIdx.y;
  const int cg0 = static_cast<int>(blockIdx.z) * cgb;
  const int ncg = (CG - cg0) < cgb ? (CG - cg0) : cgb;
  const int row_w = cgb * 8;
  for (int i = threadIdx.x; i < 27 * ncg * 8; i += blockDim.x) {
    const int t = i / (ncg * 8), cl = i - t * (ncg * 8);
    sm_w[t * row_w + cl] = w[static_cast<long long>(cg0 * 8 + cl) * 27 + (flip ? 26 - t : t)];
  }
  __syncthreads();
  MfMap m = mf_map(CG, cgb);
  if (!m.active) return;
  const int cl0 = (m.cg - cg0) * 8;
  const long long V = static_cast<long long>(D) * H * W;
  // A thread produces a run of kDwRun consecutive x positions of one row: the 3 x 3 x (kDwRun + 2) input columns it needs are
  // loaded once (13.5 loads per output voxel instead of 27) and every weight slice is read from shared memory once per run.
  // A block walks a CONTIGUOUS range of runs (not a grid stride): the rows y - 1, y of the three input planes it needs for row
  // y were loaded for the previous rows and are still in L1 — with a grid stride every neighbour row came from L2
  // (256 channels at 64^3: 273 us -> see profiles/r02_medformer_trace.txt).
  const int rpr = (W + kDwRun - 1) / kDwRun;
  const long long runs = static_cast<long long>(D) * H * rpr;
  const int vpb = blockDim.x / cgb;
  const long long per = ((runs + gridDim.x - 1) / gridDim.x + vpb - 1) / vpb * vpb;
  const long long r_lo = static_cast<long long>(blockIdx.x) * per;
  const long long r_hi = r_lo + per < runs ? r_lo + per : runs;
  for (long long r = r_lo + threadIdx.x / cgb; r < r_hi; r += vpb) {
    const unsigned ru = static_cast<unsigned>(r);
    const unsigned row = ru / static_cast<unsigned>(rpr);
    const int x0 = static_cast<int>(ru - row * rpr) * kDwRun;
    const int z = static_cast<int>(row / static_cast<unsigned>(H));
    const int yy = static_cast<int>(row - static_cast<unsigned>(z) * H);
    float2 acc[kDwRun][4];           // packed fp32 pairs: the 96 FMAs per (dz, dy) issue as 48 FFMA2
#pragma unroll
    for (int i = 0; i < kDwRun; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
    for (int dz = 0; dz < 3; ++dz) {
      const int zz = z + dz - 1;
      if (zz < 0 || zz >= D) continue;
      for (int dy = 0; dy < 3; ++dy) {
        const int y2 = yy + dy - 1;
        if (y2 < 0 || y2 >= H) continue;
        const long long rowbase = ((static_cast<long long>(n) * D + zz) * H + y2) * W;
        float2 col[kDwRun + 2][4];
#pragma unroll
        for (int i = 0; i < kDwRun + 2; ++i) {
          const int x2 = x0 + i - 1;
          float f[8];
          if (x2 >= 0 && x2 < W) {
            Vec8<T>::load(a + (rowbase + x2) * ap + m.cg * 8, f);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = 0.f;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) col[i][j] = make_float2(f[2 * j], f[2 * j + 1]);
        }
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const float4* wt = reinterpret_cast<const float4*>(sm_w + ((dz * 3 + dy) * 3 + dx) * row_w + cl0);
          const float4 wa = wt[0], wb = wt[1];
          const float2 w2[4] = {make_float2(wa.x, wa.y), make_float2(wa.z, wa.w), make_float2(wb.x, wb.y), make_float2(wb.z, wb.w)};
#pragma unroll
          for (int i = 0; i < kDwRun; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = __ffma2_rn(col[i + dx][j], w2[j], acc[i][j]);
        }
      }
    }
    const long long obase = ((static_cast<long long>(n) * D + z) * H + yy) * W;
#pragma unroll
    for (int i = 0; i < kDwRun; ++i) {
      const float o[8] = {acc[i][0].x, acc[i][0].y, acc[i][1].x, acc[i][1].y, acc[i][2].x, acc[i][2].y, acc[i][3].x, acc[i][3].y};
      if (x0 + i < W) Vec8<T>::store(y + (obase + x0 + i) * yp + m.cg * 8, o);
    }
  }
}

// dw[c][t] += sum_{n,v} dy[n,v,c] * a[n, v + t - 1, c]; blockIdx.z = chunk * 3 + kd (nine taps per block: 72 accumulators)
template <typename T>
__global__ void dwconv3_wgrad_kernel(const T* __restrict__ a, long long ap, const T* __restrict__ dy, long long dyp,
                                     float* __restrict__ dw, int D, int H, int W, int C, int cgb) {
  extern __shared__ float sm_acc[];   // [9][cgb * 8]
  const int CG = C / 8;
  const int n = blockIdx.y;
  // blockIdx.x = 3 * (voxel range) + kd: the three blocks that read the same voxel range are scheduled together, so two of
  // the three passes over dy / a hit L2 (with kd on blockIdx.z they ran as three separate sweeps: 854 MB of DRAM traffic for
  // 268 MB of tensors at 256 channels x 64^3, profiles/r02_medformer_kernels_ncu_summary.txt)
  const int chunk = static_cast<int>(blockIdx.z), dz = static_cast<int>(blockIdx.x) % 3;
  const int bx = static_cast<int>(blockIdx.x) / 3, nbx = static_cast<int>(gridDim.x) / 3;
  const int cg0 = chunk * cgb;
  const int row = cgb * 8;
  for (int i = threadIdx.x; i < 9 * row; i += blockDim.x) sm_acc[i] = 0.f;
  __syncthreads();
  const int vpb = blockDim.x / cgb;
  const int cgl = static_cast<int>(threadIdx.x) % cgb;
  const int cg = cg0 + cgl;
  const bool active = static_cast<int>(threadIdx.x) < vpb * cgb && cg < CG;
  const long long V = static_cast<long long>(D) * H * W;
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
  if (active) {
    // contiguous voxel range per block (see dwconv3_kernel): the three input rows of plane z + dz - 1 stay in L1 from row to row
    const long long per = ((V + nbx - 1) / nbx + vpb - 1) / vpb * vpb;
    const long long v_lo = static_cast<long long>(bx) * per;
    const long long v_hi = v_lo + per < V ? v_lo + per : V;
    for (long long v = v_lo + threadIdx.x / cgb; v < v_hi; v += vpb) {
      const unsigned vu = static_cast<unsigned>(v);
      const unsigned r = vu / static_cast<unsigned>(W);
      const int x = static_cast<int>(vu - r * W);
      const int z = static_cast<int>(r / static_cast<unsigned>(H));
      const int yy = static_cast<int>(r - static_cast<unsigned>(z) * H);
      const int zz = z + dz - 1;
      if (zz < 0 || zz >= D) continue;
      float g[8];
      Vec8<T>::load(dy + (static_cast<long long>(n) * V + v) * dyp + cg * 8, g);
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int y2 = yy + t / 3 - 1, x2 = x + t % 3 - 1;
        if (y2 < 0 || y2 >= H || x2 < 0 || x2 >= W) continue;
        const long long vin = ((static_cast<long long>(n) * D + zz) * H + y2) * W + x2;
        float f[8];
        Vec8<T>::load(a + vin * ap + cg * 8, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[t][j] = fmaf(g[j], f[j], acc[t][j]);
      }
    }
  }
  // block reduction into sm_acc.  Few voxel slots per block: the threads that share a channel group take turns (conflict
  // free, no atomics; every thread of the block passes the same barriers); many slots (small C): shared-memory atomics.
  if (vpb <= 16) {
    for (int k = 0; k < vpb; ++k) {
      if (active && static_cast<int>(threadIdx.x) / cgb == k) {
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
          for (int j = 0; j < 8; ++j) sm_acc[t * row + cgl * 8 + j] += acc[t][j];
      }
      __syncthreads();
    }
  } else {
    if (active) {
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(&sm_acc[t * row + cgl * 8 + j], acc[t][j]);
    }
    __syncthreads();
  }
  const int ncg = (CG - cg0) < cgb ? (CG - cg0) : cgb;
  for (int i = threadIdx.x; i < 9 * ncg * 8; i += blockDim.x) {
    const int t = i / (ncg * 8), cl = i - t * (ncg * 8);
    atomicAdd(&dw[static_cast<long long>(cg0 * 8 + cl) * 27 + dz * 9 + t], sm_acc[t * row + cl]);
  }
}

// ------------------------------------------------------------------------------------------
// per-(sample, channel) scale (SEBlock) and the per-(sample, channel) dot product its backward needs
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void scale_channels_kernel(const T* __restrict__ x, long long xp, const float* __restrict__ s, T* __restrict__ y,
                                      long long yp, int C, long long V, int cgb) {
  const int CG = C / 8;
  const int n = blockIdx.y;
  MfMap m = mf_map(CG, cgb);
  if (!m.active) return;
  float sc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sc[j] = s[static_cast<long long>(n) * C + m.cg * 8 + j];
  for (long long v = m.v0; v < V; v += m.vstride) {
    float f[8];
    Vec8<T>::load(x + (static_cast<long long>(n) * V + v) * xp + m.cg * 8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] *= sc[j];
    Vec8<T>::store(y + (static_cast<long long>(n) * V + v) * yp + m.cg * 8, f);
  }
}

template <typename T>
__global__ void channel_dot_kernel(const T* __restrict__ a, long long ap, const T* __restrict__ b, long long bp,
                                   float* __restrict__ out, int C, long long V, int cgb) {
  extern __shared__ float sm_acc[];   // [cgb * 8]
  const int CG = C / 8;
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < cgb * 8; i += blockDim.x) sm_acc[i] = 0.f;
  __syncthreads();
  MfMap m = mf_map(CG, cgb);
  if (m.active) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (long long v = m.v0; v < V; v += m.vstride) {
      float f[8], g[8];
      Vec8<T>::load(a + (static_cast<long long>(n) * V + v) * ap + m.cg * 8, f);
      Vec8<T>::load(b + (static_cast<long long>(n) * V + v) * bp + m.cg * 8, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(f[j], g[j], acc[j]);
    }
    const int cgl = static_cast<int>(threadIdx.x) % cgb;
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&sm_acc[cgl * 8 + j], acc[j]);
  }
  __syncthreads();
  const int cg0 = static_cast<int>(blockIdx.z) * cgb;
  const int ncg = (CG - cg0) < cgb ? (CG - cg0) : cgb;
  for (int i = threadIdx.x; i < ncg * 8; i += blockDim.x) atomicAdd(&out[static_cast<long long>(n) * C + cg0 * 8 + i], sm_acc[i]);
}

// ------------------------------------------------------------------------------------------
// column softmax over the voxels: running (max, sum of exp) per column, per block, then merged
// ------------------------------------------------------------------------------------------
RSB_DEVICE void online_merge(float& m, float& s, float m2, float s2) {
  const float mm = fmaxf(m, m2);
  if (mm == -INFINITY) { m = mm; s = 0.f; return; }
  s = s * __expf(m - mm) + s2 * __expf(m2 - mm);
  m = mm;
}

// partial[(row * nblk + blk) * 64 + {k, 32 + k}] = (max, sum) of block blk; sm = [warps][64] scratch
RSB_DEVICE void colstats_block_flush(float* sm, float m, float s, float* partial_row) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  sm[warp * 64 + lane] = m;
  sm[warp * 64 + 32 + lane] = s;
  __syncthreads();
  if (warp == 0) {
    float mm = sm[lane], ss = sm[32 + lane];
    for (int w2 = 1; w2 < nwarps; ++w2) online_merge(mm, ss, sm[w2 * 64 + lane], sm[w2 * 64 + 32 + lane]);
    partial_row[lane] = mm;
    partial_row[32 + lane] = ss;
  }
}

template <typename T>
__global__ void col_softmax_partial_kernel(const T* __restrict__ x, long long xp, int K, long long V, float* __restrict__ partial) {
  __shared__ float sm[8 * 64];
  const int n = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  float m = -INFINITY, s = 0.f;
  if (lane < K) {
    for (long long v = static_cast<long long>(blockIdx.x) * nwarps + warp; v < V; v += static_cast<long long>(gridDim.x) * nwarps) {
      const float a = mf_ld(x + ((static_cast<long long>(n) * V + v) * xp + lane));
      online_merge(m, s, a, 1.f);
    }
  }
  colstats_block_flush(sm, m, s, partial + (static_cast<long long>(n) * gridDim.x + blockIdx.x) * 64);
}

// ms[row * 64 + {k, 32 + k}] = merged (max, sum) over the nblk partials of a row; one block of 8 warps per row: warp w merges
// the partials w, w + 8, ..., warp 0 merges the eight results (a single warp walking 256 partials took 64 us)
__global__ void colstats_merge_kernel(const float* __restrict__ partial, int nblk, float* __restrict__ ms) {
  __shared__ float sm[8 * 64];
  const int row = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  float m = -INFINITY, s = 0.f;
  for (int b = warp; b < nblk; b += nwarps) {
    const float* p = partial + (static_cast<long long>(row) * nblk + b) * 64;
    online_merge(m, s, p[lane], p[32 + lane]);
  }
  colstats_block_flush(sm, m, s, ms + static_cast<long long>(row) * 64);
}

// ------------------------------------------------------------------------------------------
// SemanticMapGeneration: smap[n,c,k] = sum_v feat[n,v,c] * softmax_v(logit[n,:,k])[v]
// ------------------------------------------------------------------------------------------
constexpr int kPoolVB = 32;   // voxels staged per round

template <typename T>
__global__ void softmax_pool_fwd_kernel(const T* __restrict__ feat, long long fp, const T* __restrict__ logit, long long lp,
                                        const float* __restrict__ ms, float* __restrict__ smap, int C, int K, long long V) {
  __shared__ float p_s[kPoolVB][32];
  const int n = blockIdx.y;
  const float* msn = ms + static_cast<long long>(n) * 64;
  float acc[2][27];
#pragma unroll
  for (int o = 0; o < 2; ++o)
#pragma unroll
    for (int k = 0; k < 27; ++k) acc[o][k] = 0.f;
  // a block owns the voxel range [lo, hi)
  const long long per = (V + gridDim.x - 1) / gridDim.x;
  const long long lo = static_cast<long long>(blockIdx.x) * per;
  const long long hi = lo + per < V ? lo + per : V;
  for (long long base = lo; base < hi; base += kPoolVB) {
    __syncthreads();
    for (int i = threadIdx.x; i < kPoolVB * 32; i += blockDim.x) {
      const int vv = i >> 5, k = i & 31;
      float p = 0.f;
      if (k < K && base + vv < hi)
        p = __expf(mf_ld(logit + ((static_cast<long long>(n) * V + base + vv) * lp + k)) - msn[k]) / msn[32 + k];
      p_s[vv][k] = p;
    }
    __syncthreads();
    const int nv = (hi - base) < kPoolVB ? static_cast<int>(hi - base) : kPoolVB;
#pragma unroll
    for (int o = 0; o < 2; ++o) {
      const int c = static_cast<int>(threadIdx.x) + o * blockDim.x;
      if (c >= C) continue;
      for (int vv = 0; vv < nv; ++vv) {
        const float f = mf_ld(feat + ((static_cast<long long>(n) * V + base + vv) * fp + c));
#pragma unroll
        for (int k = 0; k < 27; ++k) acc[o][k] = fmaf(f, p_s[vv][k], acc[o][k]);
      }
    }
  }
#pragma unroll
  for (int o = 0; o < 2; ++o) {
    const int c = static_cast<int>(threadIdx.x) + o * blockDim.x;
    if (c >= C) continue;
    for (int k = 0; k < K; ++k) atomicAdd(&smap[(static_cast<long long>(n) * C + c) * K + k], acc[o][k]);
  }
}

// d_feat[v,c] = sum_k p[v,k] dS[c,k];  d_logit[v,k] = p[v,k] (sum_c feat[v,c] dS[c,k] - t[k]),  t[k] = sum_c dS[c,k] S[c,k]
template <typename T>
__global__ void softmax_pool_bwd_kernel(const T* __restrict__ feat, long long fp, const T* __restrict__ logit, long long lp,
                                        const float* __restrict__ ms, const float* __restrict__ dS, const float* __restrict__ tk,
                                        T* __restrict__ dfeat, long long dfp, T* __restrict__ dlogit, long long dlp, int C, int K,
                                        int Kp, long long V) {
  extern __shared__ float dS_s[];   // [C][29]
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < C * K; i += blockDim.x) {
    const int c = i / K, k = i - c * K;
    dS_s[c * 29 + k] = dS[(static_cast<long long>(n) * C + c) * K + k];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const float* msn = ms + static_cast<long long>(n) * 64;
  const float mk = lane < K ? msn[lane] : 0.f, sk = lane < K ? msn[32 + lane] : 1.f;
  const float t = lane < K ? tk[static_cast<long long>(n) * K + lane] : 0.f;
  for (long long v = static_cast<long long>(blockIdx.x) * nwarps + warp; v < V; v += static_cast<long long>(gridDim.x) * nwarps) {
    const long long row = static_cast<long long>(n) * V + v;
    float p = 0.f;
    if (lane < K) p = __expf(mf_ld(logit + (row * lp + lane)) - mk) / sk;
    // d_logit: lane = k
    float dp = 0.f;
    if (lane < K)
      for (int c = 0; c < C; ++c) dp = fmaf(mf_ld(feat + (row * fp + c)), dS_s[c * 29 + lane], dp);
    if (lane < Kp) mf_st(dlogit + (row * dlp + lane), lane < K ? p * (dp - t) : 0.f);
    // d_feat: lane strides the channels
    for (int c0 = 0; c0 < C; c0 += 32) {
      const int c = c0 + lane;
      float a = 0.f;
      for (int k = 0; k < K; ++k) {
        const float pk = __shfl_sync(0xffffffffu, p, k);
        if (c < C) a = fmaf(pk, dS_s[c * 29 + k], a);
      }
      if (c < C) mf_st(dfeat + (row * dfp + c), a);
    }
  }
}

// ------------------------------------------------------------------------------------------
// BidirectionAttention (medformer_utils.py:67-103).  Per (sample, head): A[i,j] = scale * q_i . mq_j over the voxels i and the
// J <= 32 map tokens j; P1 = softmax_j(A), P2 = softmax_i(A);
//   feat_out_i = sum_j P1[i,j] mv_j          map_out_j = sum_i P2[i,j] fv_i
// Channel c of a voxel-side tensor belongs to (dim d = c / heads, head h = c % heads) ('b (dim_head heads) d h w').
// One warp per voxel: lane j owns the token, lane d owns the head dimension (dim_head <= 32).  The logits are recomputed in
// every pass (dim_head FMAs per lane) instead of being stored.  Map-side tensors are fp32 [N, heads, J, dh].
// ------------------------------------------------------------------------------------------
constexpr int kAttPad = 33;

RSB_DEVICE float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// A[i, lane] for the voxel whose q values sit in the lanes (lane d holds q_d)
RSB_DEVICE float att_logit(float qd, const float* mq_s, int dh, int lane, float scale) {
  float a = 0.f;
  for (int d = 0; d < dh; ++d) a = fmaf(__shfl_sync(0xffffffffu, qd, d), mq_s[d * kAttPad + lane], a);
  return a * scale;
}

RSB_DEVICE void att_stage(float* dst, const float* src, int J, int dh) {   // src [J][dh] -> dst[d * 33 + j], zero padded
  for (int i = threadIdx.x; i < 32 * kAttPad; i += blockDim.x) dst[i] = 0.f;
  __syncthreads();
  for (int i = threadIdx.x; i < J * dh; i += blockDim.x) {
    const int j = i / dh, d = i - j * dh;
    dst[d * kAttPad + j] = src[i];
  }
  __syncthreads();
}

template <typename T>
__global__ void attn_colstats_kernel(const T* __restrict__ q, long long qp, const float* __restrict__ mq, float* __restrict__ partial,
                                     int heads, int dh, int J, long long V, float scale) {
  __shared__ float mq_s[32 * kAttPad];
  __shared__ float sm[8 * 64];
  const int nh = blockIdx.y, n = nh / heads, h = nh - n * heads;
  att_stage(mq_s, mq + static_cast<long long>(nh) * J * dh, J, dh);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  float m = -INFINITY, s = 0.f;
  for (long long v = static_cast<long long>(blockIdx.x) * nwarps + warp; v < V; v += static_cast<long long>(gridDim.x) * nwarps) {
    const long long row = static_cast<long long>(n) * V + v;
    const float qd = lane < dh ? mf_ld(q + (row * qp + lane * heads + h)) : 0.f;
    const float a = att_logit(qd, mq_s, dh, lane, scale);
    if (lane < J) online_merge(m, s, a, 1.f);
  }
  colstats_block_flush(sm, m, s, partial + (static_cast<long long>(nh) * gridDim.x + blockIdx.x) * 64);
}

template <typename T>
__global__ void attn_fwd_kernel(const T* __restrict__ q, const T* __restrict__ fv, long long qp, const float* __restrict__ mq,
                                const float* __restrict__ mv, const float* __restrict__ ms, T* __restrict__ fo, long long fop,
                                float* __restrict__ mo, int heads, int dh, int J, long long V, float scale) {
  __shared__ float mq_s[32 * kAttPad];
  __shared__ float mv_s[32 * kAttPad];
  __shared__ float mo_s[32 * kAttPad];
  const int nh = blockIdx.y, n = nh / heads, h = nh - n * heads;
  att_stage(mq_s, mq + static_cast<long long>(nh) * J * dh, J, dh);
  att_stage(mv_s, mv + static_cast<long long>(nh) * J * dh, J, dh);
  for (int i = threadIdx.x; i < 32 * kAttPad; i += blockDim.x) mo_s[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const float cm = ms[static_cast<long long>(nh) * 64 + lane], cs = ms[static_cast<long long>(nh) * 64 + 32 + lane];
  float macc[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) macc[d] = 0.f;
  for (long long v = static_cast<long long>(blockIdx.x) * nwarps + warp; v < V; v += static_cast<long long>(gridDim.x) * nwarps) {
    const long long row = static_cast<long long>(n) * V + v;
    const float qd = lane < dh ? mf_ld(q + (row * qp + lane * heads + h)) : 0.f;
    const float vd = lane < dh ? mf_ld(fv + (row * qp + lane * heads + h)) : 0.f;
    const float a = att_logit(qd, mq_s, dh, lane, scale);
    // P1: softmax over the tokens (lanes)
    const float rmax = warp_max(lane < J ? a : -INFINITY);
    const float e = lane < J ? __expf(a - rmax) : 0.f;
    const float p1 = e / warp_sum(e);
    float o = 0.f;
    for (int j = 0; j < J; ++j) o = fmaf(__shfl_sync(0xffffffffu, p1, j), mv_s[lane * kAttPad + j], o);   // lane = d
    if (lane < dh) mf_st(fo + (row * fop + lane * heads + h), o);
    // P2: softmax over the voxels (column statistics from the first pass)
    const float p2 = lane < J ? __expf(a - cm) / cs : 0.f;
#pragma unroll
    for (int d = 0; d < 32; ++d) {
      const float vv = __shfl_sync(0xffffffffu, vd, d);
      if (d < dh) macc[d] = fmaf(p2, vv, macc[d]);
    }
  }
  if (lane < J) {
#pragma unroll
    for (int d = 0; d < 32; ++d)
      if (d < dh) atomicAdd(&mo_s[d * kAttPad + lane], macc[d]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < J * dh; i += blockDim.x) {
    const int j = i / dh, d = i - j * dh;
    atomicAdd(&mo[static_cast<long long>(nh) * J * dh + i], mo_s[d * kAttPad + j]);
  }
}

// backward: dfo [N,V,C] (T), dmo [N,heads,J,dh], tj[nh, j] = dmo_j . map_out_j  ->  dq, dfv (T, pitch dqp), dmq, dmv (fp32, zeroed)
template <typename T>
__global__ void attn_bwd_kernel(const T* __restrict__ q, const T* __restrict__ fv, long long qp, const float* __restrict__ mq,
                                const float* __restrict__ mv, const float* __restrict__ ms, const T* __restrict__ dfo, long long dfop,
                                const float* __restrict__ dmo, const float* __restrict__ tj, T* __restrict__ dq, T* __restrict__ dfv,
                                long long dqp, float* __restrict__ dmq, float* __restrict__ dmv, int heads, int dh, int J, long long V,
                                float scale) {
  __shared__ float mq_s[32 * kAttPad];
  __shared__ float mv_s[32 * kAttPad];
  __shared__ float dmo_s[32 * kAttPad];
  __shared__ float acc_q[32 * kAttPad];
  __shared__ float acc_v[32 * kAttPad];
  const int nh = blockIdx.y, n = nh / heads, h = nh - n * heads;
  att_stage(mq_s, mq + static_cast<long long>(nh) * J * dh, J, dh);
  att_stage(mv_s, mv + static_cast<long long>(nh) * J * dh, J, dh);
  att_stage(dmo_s, dmo + static_cast<long long>(nh) * J * dh, J, dh);
  for (int i = threadIdx.x; i < 32 * kAttPad; i += blockDim.x) { acc_q[i] = 0.f; acc_v[i] = 0.f; }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const float cm = ms[static_cast<long long>(nh) * 64 + lane], cs = ms[static_cast<long long>(nh) * 64 + 32 + lane];
  const float t2 = lane < J ? tj[static_cast<long long>(nh) * J + lane] : 0.f;
  float gq[32], gv[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) { gq[d] = 0.f; gv[d] = 0.f; }
  for (long long v = static_cast<long long>(blockIdx.x) * nwarps + warp; v < V; v += static_cast<long long>(gridDim.x) * nwarps) {
    const long long row = static_cast<long long>(n) * V + v;
    const float qd = lane < dh ? mf_ld(q + (row * qp + lane * heads + h)) : 0.f;
    const float vd = lane < dh ? mf_ld(fv + (row * qp + lane * heads + h)) : 0.f;
    const float gd = lane < dh ? mf_ld(dfo + (row * dfop + lane * heads + h)) : 0.f;
    const float a = att_logit(qd, mq_s, dh, lane, scale);
    const float rmax = warp_max(lane < J ? a : -INFINITY);
    const float e = lane < J ? __expf(a - rmax) : 0.f;
    const float p1 = e / warp_sum(e);
    const float p2 = lane < J ? __expf(a - cm) / cs : 0.f;
    // lane = j: dP1_j = dfo . mv_j ; dP2_j = dmo_j . fv
    float dp1 = 0.f, dp2 = 0.f;
    for (int d = 0; d < dh; ++d) {
      dp1 = fmaf(__shfl_sync(0xffffffffu, gd, d), mv_s[d * kAttPad + lane], dp1);
      dp2 = fmaf(__shfl_sync(0xffffffffu, vd, d), dmo_s[d * kAttPad + lane], dp2);
    }
    const float r1 = warp_sum(p1 * dp1);
    const float dA = p1 * (dp1 - r1) + p2 * (dp2 - t2);     // zero in the lanes >= J (p1 = p2 = 0)
    // lane = d: dq_d = scale * sum_j dA_j mq[j][d] ; dfv_d = sum_j P2_j dmo[j][d]
    float oq = 0.f, ov = 0.f;
    for (int j = 0; j < J; ++j) {
      oq = fmaf(__shfl_sync(0xffffffffu, dA, j), mq_s[lane * kAttPad + j], oq);
      ov = fmaf(__shfl_sync(0xffffffffu, p2, j), dmo_s[lane * kAttPad + j], ov);
    }
    if (lane < dh) {
      mf_st(dq + (row * dqp + lane * heads + h), oq * scale);
      mf_st(dfv + (row * dqp + lane * heads + h), ov);
    }
    // lane = j rows of the map-side gradients
#pragma unroll
    for (int d = 0; d < 32; ++d) {
      const float qq = __shfl_sync(0xffffffffu, qd, d), gg = __shfl_sync(0xffffffffu, gd, d);
      if (d < dh) { gq[d] = fmaf(dA, qq, gq[d]); gv[d] = fmaf(p1, gg, gv[d]); }
    }
  }
  if (lane < J) {
#pragma unroll
    for (int d = 0; d < 32; ++d)
      if (d < dh) { atomicAdd(&acc_q[d * kAttPad + lane], gq[d] * scale); atomicAdd(&acc_v[d * kAttPad + lane], gv[d]); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < J * dh; i += blockDim.x) {
    const int j = i / dh, d = i - j * dh;
    atomicAdd(&dmq[static_cast<long long>(nh) * J * dh + i], acc_q[d * kAttPad + j]);
    atomicAdd(&dmv[static_cast<long long>(nh) * J * dh + i], acc_v[d * kAttPad + j]);
  }
}

}  // namespace rsb

using namespace rsb;

#define MF_BY_DTYPE(dtype, CALL_BF16, CALL_F32)            \
  if ((dtype) == RSB_BF16) { CALL_BF16; }                  \
  else if ((dtype) == RSB_F32) { CALL_F32; }               \
  else { set_last_error("bad dtype %d", (dtype)); return -1; }

#define MF_COMMON(C_, N_)                                                                                   \
  RSB_REQUIRE((C_) > 0 && (C_) % 8 == 0, "channel count must be a positive multiple of 8 (got %d)", (C_)); \
  RSB_REQUIRE((N_) > 0 && (N_) <= 65535, "bad batch size %d", (N_));                                       \
  const int sms = rsb_num_sms();                                                                            \
  RSB_REQUIRE(sms > 0, "no CUDA device");                                                                   \
  cudaStream_t st = static_cast<cudaStream_t>(stream);

extern "C" int rsb_dwconv3_forward(const void* a, int a_pitch, const float* w, void* y, int y_pitch, int dtype, int flip, int N, int D,
                                   int H, int W, int C, void* stream) {
  RSB_REQUIRE(a && w && y, "dwconv3: null pointer");
  RSB_REQUIRE(D > 0 && H > 0 && W > 0, "dwconv3: bad geometry");
  MF_COMMON(C, N)
  const long long V = static_cast<long long>(D) * H * W;
  RSB_REQUIRE(V < (1LL << 31), "dwconv3: volume too large");
  const MfGrid g = mf_grid(C, static_cast<long long>(D) * H * ((W + kDwRun - 1) / kDwRun), sms);
  dim3 grid(g.gx, N, g.chunks);
  const size_t sm = sizeof(float) * 27 * g.cgb * 8;
  MF_BY_DTYPE(dtype,
              (dwconv3_kernel<__nv_bfloat16><<<grid, kMfBlock, sm, st>>>((const __nv_bfloat16*)a, a_pitch, w, (__nv_bfloat16*)y, y_pitch, D, H, W, C, g.cgb, flip)),
              (dwconv3_kernel<float><<<grid, kMfBlock, sm, st>>>((const float*)a, a_pitch, w, (float*)y, y_pitch, D, H, W, C, g.cgb, flip)))
  return check_launch("dwconv3");
}

extern "C" int rsb_dwconv3_wgrad(const void* a, int a_pitch, const void* dy, int dy_pitch, int dtype, float* dw, int N, int D, int H,
                                 int W, int C, void* stream) {
  RSB_REQUIRE(a && dy && dw, "dwconv3_wgrad: null pointer");
  RSB_REQUIRE(D > 0 && H > 0 && W > 0, "dwconv3_wgrad: bad geometry");
  MF_COMMON(C, N)
  const long long V = static_cast<long long>(D) * H * W;
  RSB_REQUIRE(V < (1LL << 31), "dwconv3_wgrad: volume too large");
  cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * 27 * C, st);
  RSB_REQUIRE(e == cudaSuccess, "dwconv3_wgrad: memset failed: %s", cudaGetErrorString(e));
  const MfGrid g = mf_grid(C, V, sms, 3, 16);
  dim3 grid(g.gx * 3, N, g.chunks);
  const size_t sm = sizeof(float) * 9 * g.cgb * 8;
  MF_BY_DTYPE(dtype,
              (dwconv3_wgrad_kernel<__nv_bfloat16><<<grid, kMfBlock, sm, st>>>((const __nv_bfloat16*)a, a_pitch, (const __nv_bfloat16*)dy, dy_pitch, dw, D, H, W, C, g.cgb)),
              (dwconv3_wgrad_kernel<float><<<grid, kMfBlock, sm, st>>>((const float*)a, a_pitch, (const float*)dy, dy_pitch, dw, D, H, W, C, g.cgb)))
  return check_launch("dwconv3_wgrad");
}

extern "C" int rsb_scale_channels(const void* x, int x_pitch, const float* s, void* y, int y_pitch, int dtype, int N, long long V, int C,
                                  void* stream) {
  RSB_REQUIRE(x && s && y && V > 0, "scale_channels: bad arguments");
  MF_COMMON(C, N)
  const MfGrid g = mf_grid(C, V, sms);
  dim3 grid(g.gx, N, g.chunks);
  MF_BY_DTYPE(dtype,
              (scale_channels_kernel<__nv_bfloat16><<<grid, kMfBlock, 0, st>>>((const __nv_bfloat16*)x, x_pitch, s, (__nv_bfloat16*)y, y_pitch, C, V, g.cgb)),
              (scale_channels_kernel<float><<<grid, kMfBlock, 0, st>>>((const float*)x, x_pitch, s, (float*)y, y_pitch, C, V, g.cgb)))
  return check_launch("scale_channels");
}

extern "C" int rsb_channel_dot(const void* a, int a_pitch, const void* b, int b_pitch, int dtype, float* out, int N, long long V, int C,
                               void* stream) {
  RSB_REQUIRE(a && b && out && V > 0, "channel_dot: bad arguments");
  MF_COMMON(C, N)
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * static_cast<size_t>(N) * C, st);
  RSB_REQUIRE(e == cudaSuccess, "channel_dot: memset failed: %s", cudaGetErrorString(e));
  const MfGrid g = mf_grid(C, V, sms, 1, 8);
  dim3 grid(g.gx, N, g.chunks);
  const size_t sm = sizeof(float) * g.cgb * 8;
  MF_BY_DTYPE(dtype,
              (channel_dot_kernel<__nv_bfloat16><<<grid, kMfBlock, sm, st>>>((const __nv_bfloat16*)a, a_pitch, (const __nv_bfloat16*)b, b_pitch, out, C, V, g.cgb)),
              (channel_dot_kernel<float><<<grid, kMfBlock, sm, st>>>((const float*)a, a_pitch, (const float*)b, b_pitch, out, C, V, g.cgb)))
  return check_launch("channel_dot");
}

static inline int mf_voxel_blocks(long long V, int sms, int rows, int per_sm = 4) {
  long long want = (V + 7) / 8;                    // one warp per voxel: a few voxels per block round
  long long cap = static_cast<long long>(sms) * per_sm / (rows > 0 ? rows : 1);
  if (cap < 1) cap = 1;
  if (cap > 256) cap = 256;
  if (want > cap) want = cap;
  return static_cast<int>(want < 1 ? 1 : want);
}

extern "C" size_t rsb_colstats_workspace_floats(int rows) { return static_cast<size_t>(rows) * 256 * 64; }

/* SemanticMapGeneration (medformer_utils.py:222-235) */
extern "C" int rsb_softmax_pool_forward(const void* feat, int feat_pitch, const void* logit, int logit_pitch, int dtype, float* ms,
                                        float* workspace, float* smap, int N, long long V, int C, int K, void* stream) {
  RSB_REQUIRE(feat && logit && ms && workspace && smap, "softmax_pool: null pointer");
  RSB_REQUIRE(K > 0 && K <= 27 && C > 0 && C <= 512 && V > 0, "softmax_pool: needs K <= 27 and C <= 512 (K=%d C=%d)", K, C);
  RSB_REQUIRE(N > 0 && N <= 65535, "softmax_pool: bad batch size");
  const int sms = rsb_num_sms();
  RSB_REQUIRE(sms > 0, "no CUDA device");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nblk = mf_voxel_blocks(V, sms, N);
  dim3 grid(nblk, N);
  MF_BY_DTYPE(dtype,
              (col_softmax_partial_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)logit, logit_pitch, K, V, workspace)),
              (col_softmax_partial_kernel<float><<<grid, 256, 0, st>>>((const float*)logit, logit_pitch, K, V, workspace)))
  colstats_merge_kernel<<<N, 256, 0, st>>>(workspace, nblk, ms);
  cudaError_t e = cudaMemsetAsync(smap, 0, sizeof(float) * static_cast<size_t>(N) * C * K, st);
  RSB_REQUIRE(e == cudaSuccess, "softmax_pool: memset failed: %s", cudaGetErrorString(e));
  MF_BY_DTYPE(dtype,
              (softmax_pool_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)feat, feat_pitch, (const __nv_bfloat16*)logit, logit_pitch, ms, smap, C, K, V)),
              (softmax_pool_fwd_kernel<float><<<grid, 256, 0, st>>>((const float*)feat, feat_pitch, (const float*)logit, logit_pitch, ms, smap, C, K, V)))
  return check_launch("softmax_pool_forward");
}

extern "C" int rsb_softmax_pool_backward(const void* feat, int feat_pitch, const void* logit, int logit_pitch, int dtype, const float* ms,
                                         const float* dS, const float* tk, void* dfeat, int dfeat_pitch, void* dlogit, int dlogit_pitch,
                                         int N, long long V, int C, int K, int Kp, void* stream) {
  RSB_REQUIRE(feat && logit && ms && dS && tk && dfeat && dlogit, "softmax_pool_backward: null pointer");
  RSB_REQUIRE(K > 0 && K <= 27 && Kp >= K && Kp <= 32 && C > 0 && C <= 400 && V > 0, "softmax_pool_backward: needs K <= 27 and C <= 400");
  RSB_REQUIRE(N > 0 && N <= 65535, "softmax_pool_backward: bad batch size");
  const int sms = rsb_num_sms();
  RSB_REQUIRE(sms > 0, "no CUDA device");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid(mf_voxel_blocks(V, sms, N), N);
  const size_t sm = sizeof(float) * C * 29;
  MF_BY_DTYPE(dtype,
              (softmax_pool_bwd_kernel<__nv_bfloat16><<<grid, 256, sm, st>>>((const __nv_bfloat16*)feat, feat_pitch, (const __nv_bfloat16*)logit, logit_pitch, ms, dS, tk, (__nv_bfloat16*)dfeat, dfeat_pitch, (__nv_bfloat16*)dlogit, dlogit_pitch, C, K, Kp, V)),
              (softmax_pool_bwd_kernel<float><<<grid, 256, sm, st>>>((const float*)feat, feat_pitch, (const float*)logit, logit_pitch, ms, dS, tk, (float*)dfeat, dfeat_pitch, (float*)dlogit, dlogit_pitch, C, K, Kp, V)))
  return check_launch("softmax_pool_backward");
}

/* BidirectionAttention (medformer_utils.py:67-103) */
extern "C" int rsb_biattention_forward(const void* q, const void* fv, int qv_pitch, int dtype, const float* mq, const float* mv, float* ms,
                                       float* workspace, void* feat_out, int feat_out_pitch, float* map_out, int N, long long V,
                                       int heads, int dim_head, int J, float scale, void* stream) {
  RSB_REQUIRE(q && fv && mq && mv && ms && workspace && feat_out && map_out, "biattention: null pointer");
  RSB_REQUIRE(heads > 0 && dim_head > 0 && dim_head <= 32 && J > 0 && J <= 32 && V > 0, "biattention: needs dim_head <= 32 and <= 32 map tokens (dim_head=%d J=%d)", dim_head, J);
  RSB_REQUIRE(N > 0 && static_cast<long long>(N) * heads <= 65535, "biattention: bad batch size");
  const int sms = rsb_num_sms();
  RSB_REQUIRE(sms > 0, "no CUDA device");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int rows = N * heads;
  const int nblk = mf_voxel_blocks(V, sms, rows, 10);
  dim3 grid(nblk, rows);
  MF_BY_DTYPE(dtype,
              (attn_colstats_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>((const __nv_bfloat16*)q, qv_pitch, mq, workspace, heads, dim_head, J, V, scale)),
              (attn_colstats_kernel<float><<<grid, 128, 0, st>>>((const float*)q, qv_pitch, mq, workspace, heads, dim_head, J, V, scale)))
  colstats_merge_kernel<<<rows, 256, 0, st>>>(workspace, nblk, ms);
  cudaError_t e = cudaMemsetAsync(map_out, 0, sizeof(float) * static_cast<size_t>(rows) * J * dim_head, st);
  RSB_REQUIRE(e == cudaSuccess, "biattention: memset failed: %s", cudaGetErrorString(e));
  MF_BY_DTYPE(dtype,
              (attn_fwd_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)fv, qv_pitch, mq, mv, ms, (__nv_bfloat16*)feat_out, feat_out_pitch, map_out, heads, dim_head, J, V, scale)),
              (attn_fwd_kernel<float><<<grid, 128, 0, st>>>((const float*)q, (const float*)fv, qv_pitch, mq, mv, ms, (float*)feat_out, feat_out_pitch, map_out, heads, dim_head, J, V, scale)))
  return check_launch("biattention_forward");
}

extern "C" int rsb_biattention_backward(const void* q, const void* fv, int qv_pitch, int dtype, const float* mq, const float* mv,
                                        const float* ms, const void* dfeat_out, int dfeat_out_pitch, const float* dmap_out,
                                        const float* tj, void* dq, void* dfv, int dqv_pitch, float* dmq, float* dmv, int N, long long V,
                                        int heads, int dim_head, int J, float scale, void* stream) {
  RSB_REQUIRE(q && fv && mq && mv && ms && dfeat_out && dmap_out && tj && dq && dfv && dmq && dmv, "biattention_backward: null pointer");
  RSB_REQUIRE(heads > 0 && dim_head > 0 && dim_head <= 32 && J > 0 && J <= 32 && V > 0, "biattention_backward: needs dim_head <= 32 and <= 32 map tokens");
  RSB_REQUIRE(N > 0 && static_cast<long long>(N) * heads <= 65535, "biattention_backward: bad batch size");
  const int sms = rsb_num_sms();
  RSB_REQUIRE(sms > 0, "no CUDA device");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int rows = N * heads;
  const size_t mbytes = sizeof(float) * static_cast<size_t>(rows) * J * dim_head;
  cudaError_t e = cudaMemsetAsync(dmq, 0, mbytes, st);
  RSB_REQUIRE(e == cudaSuccess, "biattention_backward: memset failed: %s", cudaGetErrorString(e));
  e = cudaMemsetAsync(dmv, 0, mbytes, st);
  RSB_REQUIRE(e == cudaSuccess, "biattention_backward: memset failed: %s", cudaGetErrorString(e));
  dim3 grid(mf_voxel_blocks(V, sms, rows, 6), rows);   // 132 registers: three 128-thread blocks per SM
  MF_BY_DTYPE(dtype,
              (attn_bwd_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)fv, qv_pitch, mq, mv, ms, (const __nv_bfloat16*)dfeat_out, dfeat_out_pitch, dmap_out, tj, (__nv_bfloat16*)dq, (__nv_bfloat16*)dfv, dqv_pitch, dmq, dmv, heads, dim_head, J, V, scale)),
              (attn_bwd_kernel<float><<<grid, 128, 0, st>>>((const float*)q, (const float*)fv, qv_pitch, mq, mv, ms, (const float*)dfeat_out, dfeat_out_pitch, dmap_out, tj, (float*)dq, (float*)dfv, dqv_pitch, dmq, dmv, heads, dim_head, J, V, scale)))
  return check_launch("biattention_backward");
}
