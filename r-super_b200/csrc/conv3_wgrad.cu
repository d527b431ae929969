// conv3_wgrad.cu — weight gradient of the 3x3x3 convolution on tcgen05.
//
//   dW[co][ci][kd,kh,kw] = sum_{n,z,y,x} dy[n,z,y,x][co] * a[n,z+kd-1,y+kh-1,x+kw-1][ci]
//   a = act(instnorm(x)) is recomputed in the staging exactly as the forward kernel does
//   (autograd of nn.Conv3d in ConvNormAct, rsuper_train/model/dim3/conv_layers.py:29-49).
//
// GEMM view (K = voxels): both operands are read MN-major straight out of the same
// [8-channel group][voxel][8 ch] shared-memory planes the forward kernel uses — in that layout a
// core matrix is "8 voxels x 8 channels", so the channel-contiguous NDHWC rows need no transpose.
//   B (N side) = one haloed a-plane (18 x 10 voxels x IT channels); the in-plane taps (kh,kw) are
//                descriptor start offsets into it.
//   A (M side) = dy planes (16 x 8 voxels x OT channels).  M = 128 rows = PM = 128/OT consecutive
//                dy z-planes stacked, so that one MMA covers several kd taps at once
//                (OT = 32: planes z-1..z+2 -> kd = 2,1,0 + one ignored block).  The dy planes live in
//                a 4-deep ring that is written twice ("mirrored", 8 physical slots) so every window of
//                PM consecutive planes is contiguous whatever the ring phase.
//   accumulators: (kd-window, in-plane tap) x IT fp32 columns in TMEM, live for the whole kernel;
//                each persistent CTA sums over its share of the volume and dumps one partial; a
//                second tiny kernel reduces the partials deterministically into fp32 OIDHW.
// When (Cout, Cin, 27 taps) does not fit 512 TMEM columns the problem is split into groups
// (o-tile, i-tile, tap subset); CTAs are dealt round-robin to groups.
#include "rsb_common.cuh"

#include "../../include/rsuper_b200.h"

namespace rsb {

constexpr int kWgThreads = 320;
constexpr int kWgProducerWarp0 = 2;
constexpr int kWgNumProducerThreads = 256;
constexpr int kWgTileY = 16, kWgTileX = 8;
constexpr int kWgPlaneVox = 180;            // haloed a-plane
constexpr int kWgAChunkBytes = kWgPlaneVox * 16;  // 2880
constexpr int kWgDyChunkBytes = 128 * 16;         // 2048
constexpr int kWgCtrlBytes = 4096;                // barriers + per-sample norm table

struct WgradDev {
  int N, D, H, W, Cin, Cout;
  const void* x;
  long long x_pitch;
  const float* in_stats;
  float eps, slope, inv_count;
  const void* dy;
  long long dy_pitch;
  float* ws;
  // tiling
  int OT, IT, PM, KS, TS;  // o-tile, i-tile, planes per window, windows, in-plane taps per group
  int CGo, CGi;            // OT/8, IT/8
  int n_otiles, n_itiles, n_tapsets, n_groups;
  int ranks;               // CTAs per group
  int tiles_y, tiles_x, zchunks, zlen, n_units;
  int mirrored;            // dy ring written twice
  int dy_slot_bytes;       // CGo * 2048
  int piece_floats;        // 3 * TS * OT * IT
  int cgi_shift, cgo_shift;  // log2(CGi) / log2(CGo) when a power of two, else -1
  uint32_t idesc;
  long long* dbg;
};

struct __align__(16) WgradSmem {
  uint64_t full[2], empty[2], done;
  uint32_t tmem_base;
  uint32_t pad_[1];
  float2 norm[256];  // (scale, shift) of the CTA's IT input channels for the current sample
};

static_assert(sizeof(WgradSmem) <= kWgCtrlBytes, "control block overflows its smem region");

struct WgUnit {
  int n, y0, x0, zs, ze;
};
RSB_DEVICE WgUnit wg_decode_unit(const WgradDev& a, int u) {
  WgUnit r;
  int t = u;
  const int zc = t % a.zchunks; t /= a.zchunks;
  const int xt = t % a.tiles_x; t /= a.tiles_x;
  const int yt = t % a.tiles_y; t /= a.tiles_y;
  r.n = t;
  r.y0 = yt * kWgTileY;
  r.x0 = xt * kWgTileX;
  r.zs = zc * a.zlen;
  r.ze = min(a.D, r.zs + a.zlen);
  return r;
}
RSB_DEVICE int mod4(int v) { return ((v % 4) + 4) % 4; }

template <typename T>
__global__ void __launch_bounds__(kWgThreads, 1) conv3_wgrad_kernel(const WgradDev a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  WgradSmem& sm = *reinterpret_cast<WgradSmem*>(smem_raw);
  uint8_t* dy_buf = smem_raw + kWgCtrlBytes;
  const int dy_slots = a.mirrored ? 8 : 4;
  uint8_t* a_buf = dy_buf + dy_slots * a.dy_slot_bytes;
  const int a_slot_bytes = a.CGi * kWgAChunkBytes;
  const uint32_t dy_base = smem_u32(dy_buf);
  const uint32_t a_base = smem_u32(a_buf);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = blockIdx.x % a.n_groups;
  const int rank = blockIdx.x / a.n_groups;
  // group -> (o-tile, i-tile, tap set)
  const int tapset = group % a.n_tapsets;
  const int itile = (group / a.n_tapsets) % a.n_itiles;
  const int otile = group / (a.n_tapsets * a.n_itiles);
  const int o0 = otile * a.OT, i0 = itile * a.IT;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&sm.full[i]), kWgNumProducerThreads);
      mbar_init(smem_u32(&sm.empty[i]), 1);
    }
    mbar_init(smem_u32(&sm.done), 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&sm.tmem_base), 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = sm.tmem_base;

  if (warp == 0) {
    // =========================== MMA issuer ===========================
    // warp-converged waits; one elected lane issues (see conv3_igemm.cu)
    {
      const uint32_t a_hi = ((static_cast<uint32_t>(kWgDyChunkBytes) >> 4) & 0x3FFFu) | (1u << 14);  // dy: SBO = next 8 couts
      const uint32_t a_lbo = ((128u >> 4) & 0x3FFFu) << 16;                                           // next 8 voxels (y row)
      const uint32_t b_hi = ((static_cast<uint32_t>(kWgAChunkBytes) >> 4) & 0x3FFFu) | (1u << 14);   // a: SBO = next 8 cins
      const uint32_t b_lbo = ((160u >> 4) & 0x3FFFu) << 16;                                           // next y row of the halo plane
      uint32_t t = 0;
      long long tw_full = 0;
      const long long t_begin = clock64();
      for (int u = rank; u < a.n_units; u += a.ranks) {
        const WgUnit un = wg_decode_unit(a, u);
        for (int zb = un.zs; zb < un.ze; ++zb) {
          const uint32_t s = t & 1u;
          const long long tq = clock64();
          mbar_wait(smem_u32(&sm.full[s]), (t >> 1) & 1u);
          tw_full += clock64() - tq;
          tc_fence_after_sync();
          const uint32_t b_slot_lo = b_lbo | ((a_base + s * a_slot_bytes) >> 4);
          const uint32_t first = t != 0 ? 1u : 0u;
          if (elect_one()) {
          for (int w = 0; w < a.KS; ++w) {
            const int win = mod4(zb - 1 + w * a.PM);
            const uint32_t a_win_lo = a_lbo | ((dy_base + win * a.dy_slot_bytes) >> 4);
            for (int ti = 0; ti < a.TS; ++ti) {
              const int tap = tapset * a.TS + ti;  // in-plane tap index kh*3+kw
              const int kh = tap / 3, kw = tap % 3;
              const uint32_t d_col = tmem_base + (w * a.TS + ti) * a.IT;
              const uint32_t b_tap_lo = b_slot_lo + kh * 10 + kw;
              // 8 K-steps of 16 voxels (2 y-rows): dy advances 256 B, the haloed a-plane 2 * 160 B
#pragma unroll
              for (int ks = 0; ks < 8; ++ks) {
                const uint64_t adesc = (static_cast<uint64_t>(a_hi) << 32) | (a_win_lo + ks * 16);
                const uint64_t bdesc = (static_cast<uint64_t>(b_hi) << 32) | (b_tap_lo + ks * 20);
                umma_bf16_ss(d_col, adesc, bdesc, a.idesc, ks == 0 ? first : 1u);
              }
            }
          }
          umma_commit(smem_u32(&sm.empty[s]));
          }
          __syncwarp();
          ++t;
        }
      }
      if (elect_one()) umma_commit(smem_u32(&sm.done));
      __syncwarp();
      if (a.dbg != nullptr && lane == 0) {
        long long* d = a.dbg + static_cast<size_t>(blockIdx.x) * 16;
        d[0] = clock64() - t_begin; d[1] = tw_full; d[2] = t;
      }
    }
  } else if (warp >= kWgProducerWarp0) {
    // =========================== producers ===========================
    const int pt = threadIdx.x - kWgProducerWarp0 * 32;  // 0..255
    const T* __restrict__ xg = reinterpret_cast<const T*>(a.x);
    const T* __restrict__ dyg = reinterpret_cast<const T*>(a.dy);
    const bool has_norm = a.in_stats != nullptr;
    const long long xplane = static_cast<long long>(a.H) * a.W * a.x_pitch;
    const long long dplane = static_cast<long long>(a.H) * a.W * a.dy_pitch;
    constexpr int U = 4;  // independent loads in flight per thread
    uint32_t t = 0;
    int cur_n = -1;
    long long tw_empty = 0;
    const long long tp_begin = clock64();
    for (int u = rank; u < a.n_units; u += a.ranks) {
      const WgUnit un = wg_decode_unit(a, u);
      if (un.n != cur_n) {
        // per-sample (scale, shift) of this CTA's input channels: a = act(x * scale + shift)
        named_bar_sync(2, kWgNumProducerThreads);
        for (int c = pt; c < a.IT; c += kWgNumProducerThreads) {
          float2 ns = make_float2(1.f, 0.f);
          if (has_norm && i0 + c < a.Cin) {
            const float* st = a.in_stats + (static_cast<size_t>(un.n) * a.x_pitch + i0 + c) * 2;
            float mean, rstd;
            stats_to_mean_rstd(st[0], st[1], a.inv_count, a.eps, mean, rstd);
            ns = make_float2(rstd, -mean * rstd);
          }
          sm.norm[c] = ns;
        }
        named_bar_sync(2, kWgNumProducerThreads);
        cur_n = un.n;
      }
      for (int zb = un.zs; zb < un.ze; ++zb) {
        const uint32_t s = t & 1u;
        const long long tq = clock64();
        mbar_wait(smem_u32(&sm.empty[s]), ((t >> 1) & 1u) ^ 1u);
        tw_empty += clock64() - tq;
        const bool col_start = (zb == un.zs);
        if (col_start && t >= 1) {
          // a new column rewrites every dy slot: the previous step must have drained too
          mbar_wait(smem_u32(&sm.empty[s ^ 1u]), ((t - 1) >> 1) & 1u);
        }
        // ---- haloed a-plane zb -> a slot s ----
        {
          uint8_t* dst = a_buf + s * a_slot_bytes;
          const int ops = 23 * a.CGi * 8;
          const long long zbase = (static_cast<long long>(un.n) * a.D + zb) * xplane;
          for (int ib = pt; ib < ops; ib += kWgNumProducerThreads * U) {
            Raw8<T> raw[U];
            int off[U], cjs[U];
            bool inb[U];
#pragma unroll
            for (int q = 0; q < U; ++q) {
              const int i = ib + q * kWgNumProducerThreads;
              off[q] = -1;
              inb[q] = false;
              cjs[q] = 0;
              if (i < ops) {
                const int vi = i & 7;
                const int w8 = i >> 3;
                const int cj = a.cgi_shift >= 0 ? (w8 & (a.CGi - 1)) : (w8 % a.CGi);
                const int g8 = a.cgi_shift >= 0 ? (w8 >> a.cgi_shift) : (w8 / a.CGi);
                const int vox = g8 * 8 + vi;
                if (vox < kWgPlaneVox) {
                  const int yy = vox / 10, xx = vox - yy * 10;
                  const int y = un.y0 - 1 + yy, xq = un.x0 - 1 + xx;
                  const int ch = i0 + cj * 8;
                  off[q] = (cj * kWgPlaneVox + vox) * 16;
                  cjs[q] = cj;
                  inb[q] = ch < a.Cin && y >= 0 && y < a.H && xq >= 0 && xq < a.W;
                  if (inb[q]) raw[q].load(xg + zbase + (static_cast<long long>(y) * a.W + xq) * a.x_pitch + ch);
                }
              }
            }
#pragma unroll
            for (int q = 0; q < U; ++q) {
              if (off[q] >= 0) {
                uint4 o = make_uint4(0u, 0u, 0u, 0u);
                if (inb[q]) {
                  float f[8];
                  raw[q].to_float(f);
                  if (has_norm) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                      const float2 ns = sm.norm[cjs[q] * 8 + j];
                      const float h = fmaf(f[j], ns.x, ns.y);
                      f[j] = h > 0.f ? h : h * a.slope;
                    }
                  }
                  o.x = pack_bf16x2(f[0], f[1]);
                  o.y = pack_bf16x2(f[2], f[3]);
                  o.z = pack_bf16x2(f[4], f[5]);
                  o.w = pack_bf16x2(f[6], f[7]);
                }
                *reinterpret_cast<uint4*>(dst + off[q]) = o;
              }
            }
          }
        }
        // ---- dy planes: zb+1 always, zb-1 and zb at a column start ----
        for (int rel = col_start ? -1 : 1; rel <= 1; ++rel) {
          const int z = zb + rel;
          const int slot = mod4(z);
          uint8_t* dst0 = dy_buf + slot * a.dy_slot_bytes;
          const bool zin = z >= 0 && z < a.D;
          const int ops = 16 * a.CGo * 8;
          const long long zbase = (static_cast<long long>(un.n) * a.D + z) * dplane;
          for (int ib = pt; ib < ops; ib += kWgNumProducerThreads * U) {
            Raw8<T> raw[U];
            int off[U];
            bool inb[U];
#pragma unroll
            for (int q = 0; q < U; ++q) {
              const int i = ib + q * kWgNumProducerThreads;
              off[q] = -1;
              inb[q] = false;
              if (i < ops) {
                const int vi = i & 7;
                const int w8 = i >> 3;
                const int cj = a.cgo_shift >= 0 ? (w8 & (a.CGo - 1)) : (w8 % a.CGo);
                const int g8 = a.cgo_shift >= 0 ? (w8 >> a.cgo_shift) : (w8 / a.CGo);
                const int y = un.y0 + g8, xq = un.x0 + vi;
                const int ch = o0 + cj * 8;
                off[q] = (cj * 128 + g8 * 8 + vi) * 16;
                inb[q] = zin && ch < a.Cout && y < a.H && xq < a.W;
                if (inb[q]) raw[q].load(dyg + zbase + (static_cast<long long>(y) * a.W + xq) * a.dy_pitch + ch);
              }
            }
#pragma unroll
            for (int q = 0; q < U; ++q) {
              if (off[q] >= 0) {
                uint4 o = make_uint4(0u, 0u, 0u, 0u);
                if (inb[q]) {
                  float f[8];
                  raw[q].to_float(f);
                  o.x = pack_bf16x2(f[0], f[1]);
                  o.y = pack_bf16x2(f[2], f[3]);
                  o.z = pack_bf16x2(f[4], f[5]);
                  o.w = pack_bf16x2(f[6], f[7]);
                }
                *reinterpret_cast<uint4*>(dst0 + off[q]) = o;
                if (a.mirrored) *reinterpret_cast<uint4*>(dst0 + 4 * a.dy_slot_bytes + off[q]) = o;
              }
            }
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(smem_u32(&sm.full[s]));
        ++t;
      }
    }
    if (a.dbg != nullptr && pt == 0) {
      long long* d = a.dbg + static_cast<size_t>(blockIdx.x) * 16;
      d[3] = clock64() - tp_begin; d[4] = tw_empty;
    }
    // =========================== epilogue (warps 2..5) ===========================
    if (warp < kWgProducerWarp0 + 4) {
      mbar_wait(smem_u32(&sm.done), 0);
      tc_fence_after_sync();
      __syncwarp();
      const int ew = warp & 3;
      const int row = ew * 32 + lane;
      const int j = row / a.OT, co = row % a.OT;
      float* piece = a.ws + static_cast<size_t>(blockIdx.x) * a.piece_floats;
      for (int w = 0; w < a.KS; ++w) {
        const int off = w * a.PM + j;  // dy plane offset relative to zb-1  ->  kd = 2 - off
        for (int ti = 0; ti < a.TS; ++ti) {
          for (int cc = 0; cc < a.IT; cc += 16) {
            uint32_t r[16];
            tmem_ld16(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + (w * a.TS + ti) * a.IT + cc, r);
            tmem_ld_wait();
            if (off <= 2) {
              float* dst = piece + ((static_cast<size_t>(off) * a.TS + ti) * a.OT + co) * a.IT + cc;
#pragma unroll
              for (int q = 0; q < 16; q += 4)
                *reinterpret_cast<float4*>(dst + q) =
                    make_float4(__uint_as_float(r[q]), __uint_as_float(r[q + 1]), __uint_as_float(r[q + 2]),
                                __uint_as_float(r[q + 3]));
            }
          }
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

// Deterministic reduction of the per-CTA partials into fp32 OIDHW.
__global__ void wgrad_reduce_kernel(const WgradDev a, float* __restrict__ dw, int accumulate) {
  const long long total = static_cast<long long>(a.n_groups) * a.piece_floats;
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int group = static_cast<int>(e / a.piece_floats);
    int r = static_cast<int>(e % a.piece_floats);
    const int ci = r % a.IT; r /= a.IT;
    const int co = r % a.OT; r /= a.OT;
    const int ti = r % a.TS; r /= a.TS;
    const int off = r;  // 0..2
    const int tapset = group % a.n_tapsets;
    const int itile = (group / a.n_tapsets) % a.n_itiles;
    const int otile = group / (a.n_tapsets * a.n_itiles);
    const int o = otile * a.OT + co, i = itile * a.IT + ci;
    if (o >= a.Cout || i >= a.Cin) continue;
    const int tap = (2 - off) * 9 + tapset * a.TS + ti;
    float acc = 0.f;
    const size_t pe = e % a.piece_floats;
    for (int rk = 0; rk < a.ranks; ++rk)
      acc += a.ws[(static_cast<size_t>(rk) * a.n_groups + group) * a.piece_floats + pe];
    float* dst = dw + (static_cast<size_t>(o) * a.Cin + i) * 27 + tap;
    *dst = accumulate ? (*dst + acc) : acc;
  }
}

static inline int round_up_i(int v, int m) { return (v + m - 1) / m * m; }

// Tiling plan shared by the workspace query and the launcher.
static int wgrad_plan(int Cout, int Cin, int N, int D, int H, int W, int max_ctas, WgradDev& d) {
  d.OT = Cout <= 32 ? 32 : (Cout <= 64 ? 64 : 128);
  d.PM = 128 / d.OT;
  d.KS = (3 + d.PM - 1) / d.PM;
  d.n_otiles = (Cout + d.OT - 1) / d.OT;
  const int ci_pad = round_up_i(Cin, 16);
  // prefer a wide i-tile (tensor-pipe efficiency), then as many in-plane taps per group as fit TMEM
  int bestIT = 16, bestTS = 1;
  long long best_score = -1;
  const int dy_bytes = (d.PM > 1 ? 8 : 4) * (d.OT / 8) * kWgDyChunkBytes;
  for (int it = 16; it <= 256 && it <= ci_pad; it += 16) {
    if (ci_pad % it) continue;
    const int a_bytes = 2 * (it / 8) * kWgAChunkBytes;
    if (kWgCtrlBytes + dy_bytes + a_bytes > 227 * 1024) continue;
    for (int ts : {9, 3, 1}) {
      if (d.KS * ts * it > 512) continue;
      const long long score = static_cast<long long>(it < 128 ? it : 128) * 16 + ts;  // IT first, then TS
      if (score > best_score) { best_score = score; bestIT = it; bestTS = ts; }
      break;
    }
  }
  if (best_score < 0) return -1;
  d.IT = bestIT;
  d.TS = bestTS;
  d.CGo = d.OT / 8;
  d.CGi = d.IT / 8;
  d.n_itiles = ci_pad / d.IT;
  d.n_tapsets = 9 / d.TS;
  d.n_groups = d.n_otiles * d.n_itiles * d.n_tapsets;
  d.mirrored = d.PM > 1 ? 1 : 0;
  d.dy_slot_bytes = d.CGo * kWgDyChunkBytes;
  d.piece_floats = 3 * d.TS * d.OT * d.IT;
  d.tiles_y = (H + kWgTileY - 1) / kWgTileY;
  d.tiles_x = (W + kWgTileX - 1) / kWgTileX;
  int ranks = max_ctas / d.n_groups;
  if (ranks < 1) ranks = 1;
  // z chunking: enough units for every rank of a group, but keep chunks long (each pays 2 extra dy planes)
  const long long cols = static_cast<long long>(N) * d.tiles_y * d.tiles_x;
  int zlen = D;
  while (zlen > 8 && cols * ((D + zlen - 1) / zlen) < 4LL * ranks) zlen = (zlen + 1) / 2;
  d.zlen = zlen;
  d.zchunks = (D + zlen - 1) / zlen;
  const long long units = cols * d.zchunks;
  if (units >= (1LL << 31)) return -1;
  d.n_units = static_cast<int>(units);
  if (ranks > d.n_units) ranks = d.n_units;
  d.ranks = ranks;
  d.idesc = make_idesc_bf16(128, d.IT, 1, 1);
  auto log2_or_neg = [](int v) { int l = 0; while ((1 << l) < v) ++l; return (1 << l) == v ? l : -1; };
  d.cgi_shift = log2_or_neg(d.CGi);
  d.cgo_shift = log2_or_neg(d.CGo);
  return 0;
}

}  // namespace rsb

using namespace rsb;

static long long* g_wgrad_timing_buffer = nullptr;
// profiling aid (see rsb_debug_set_timing_buffer): [0] MMA warp total cycles, [1] its wait on `full`,
// [2] steps, [3] producer total, [4] producer wait on `empty`
extern "C" int rsb_debug_set_wgrad_timing_buffer(void* device_ptr) {
  g_wgrad_timing_buffer = reinterpret_cast<long long*>(device_ptr);
  return 0;
}

extern "C" size_t rsb_conv3_wgrad_workspace_bytes(int Cout, int Cin, int max_ctas) {
  if (max_ctas <= 0) max_ctas = rsb_num_sms();
  if (max_ctas <= 0) max_ctas = 148;
  WgradDev d{};
  // geometry does not change the piece size; ranks * groups <= max(max_ctas, groups)
  if (wgrad_plan(Cout, Cin, 1, 64, 64, 64, max_ctas, d) != 0) return 0;
  const size_t ctas = static_cast<size_t>(max_ctas > d.n_groups ? max_ctas : d.n_groups);
  return ctas * d.piece_floats * sizeof(float);
}

extern "C" int rsb_conv3_wgrad(const RsbConv3WgradArgs* p, void* stream) {
  RSB_REQUIRE(p != nullptr, "wgrad: null args");
  RSB_REQUIRE(p->x && p->dy && p->dw_oidhw && p->workspace, "wgrad: null pointer");
  RSB_REQUIRE(p->N > 0 && p->D > 0 && p->H > 0 && p->W > 0, "wgrad: bad geometry");
  RSB_REQUIRE(p->Cin > 0 && p->Cin % 8 == 0 && p->Cout > 0 && p->Cout % 8 == 0,
              "wgrad: channel counts must be positive multiples of 8 (Cin=%d Cout=%d)", p->Cin, p->Cout);
  RSB_REQUIRE(p->x_pitch >= p->Cin && p->x_pitch % 8 == 0 && p->dy_pitch >= p->Cout && p->dy_pitch % 8 == 0,
              "wgrad: bad pitch");
  RSB_REQUIRE(p->dtype == RSB_BF16 || p->dtype == RSB_F32, "wgrad: bad dtype %d", p->dtype);
  int sms = p->max_ctas > 0 ? p->max_ctas : rsb_num_sms();
  RSB_REQUIRE(sms > 0, "wgrad: could not query the SM count");

  WgradDev d{};
  d.N = p->N; d.D = p->D; d.H = p->H; d.W = p->W; d.Cin = p->Cin; d.Cout = p->Cout;
  d.x = p->x; d.x_pitch = p->x_pitch; d.in_stats = p->in_stats; d.eps = p->eps; d.slope = p->slope;
  d.inv_count = 1.0f / (static_cast<float>(p->D) * p->H * p->W);
  d.dy = p->dy; d.dy_pitch = p->dy_pitch;
  d.ws = reinterpret_cast<float*>(p->workspace);
  d.dbg = g_wgrad_timing_buffer;
  RSB_REQUIRE(wgrad_plan(p->Cout, p->Cin, p->N, p->D, p->H, p->W, sms, d) == 0, "wgrad: no feasible tiling");
  const int grid = d.ranks * d.n_groups;
  const size_t need = static_cast<size_t>(grid) * d.piece_floats * sizeof(float);
  RSB_REQUIRE(p->workspace_bytes >= need, "wgrad: workspace too small (%zu < %zu)", p->workspace_bytes, need);

  size_t smem = kWgCtrlBytes + static_cast<size_t>(d.mirrored ? 8 : 4) * d.dy_slot_bytes + 2 * static_cast<size_t>(d.CGi) * kWgAChunkBytes;
  if (smem < 120 * 1024) smem = 120 * 1024;  // 1 CTA / SM (each CTA owns all 512 TMEM columns)
  RSB_REQUIRE(smem <= 227 * 1024, "wgrad: shared memory budget exceeded (%zu)", smem);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  if (p->dtype == RSB_BF16) {
    e = cudaFuncSetAttribute(conv3_wgrad_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    RSB_REQUIRE(e == cudaSuccess, "wgrad: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    conv3_wgrad_kernel<__nv_bfloat16><<<grid, kWgThreads, smem, st>>>(d);
  } else {
    e = cudaFuncSetAttribute(conv3_wgrad_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    RSB_REQUIRE(e == cudaSuccess, "wgrad: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    conv3_wgrad_kernel<float><<<grid, kWgThreads, smem, st>>>(d);
  }
  int rc = check_launch("conv3_wgrad_kernel");
  if (rc) return rc;
  const long long total = static_cast<long long>(d.n_groups) * d.piece_floats;
  int rblocks = static_cast<int>((total + 255) / 256);
  if (rblocks > sms * 8) rblocks = sms * 8;
  wgrad_reduce_kernel<<<rblocks, 256, 0, st>>>(d, p->dw_oidhw, p->accumulate);
  return check_launch("wgrad_reduce_kernel");
}
