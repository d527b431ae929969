// conv3_igemm.cu — 3x3x3 convolution (stride 1, pad 1, no bias) as an implicit GEMM on tcgen05.
//
// Replaces nn.Conv3d inside ConvNormAct (rsuper_train/model/dim3/conv_layers.py:29-38) together
// with the pre-activation InstanceNorm3d(eps=1e-4, affine=False) + ReLU in front of it
// (conv_layers.py:39-49), the BasicBlock residual add (conv_layers.py:92) and the statistics the
// NEXT InstanceNorm needs.  With flipped/transposed weights the same kernel is the data gradient,
// whose epilogue applies act'(xhat) and accumulates the two InstanceNorm-backward reductions.
//
// Design (B200 / sm_100a, one persistent CTA per SM, 448 threads = MMA + loader + 4 epilogue + 8 producer warps):
//   GEMM view   M = 128 output voxels (16 y x 8 x of one z-plane), N = Cout tile, K = 27 * Cin.
//   work item   (n, z-block of PZ planes, y-tile, x-tile, N-tile): PZ accumulators of 128 x NT fp32
//               live in TMEM (double-buffered when 2*PZ*NT <= 512 columns).
//   A operand   "im2col in registers": 8 producer warps read the raw NDHWC halo box
//               (PZ+2) x 18 x 10 voxels x 32 channels, apply (x-mean)*rstd and (Leaky)ReLU in
//               registers, force the zero padding, and write bf16 into shared memory in the
//               SWIZZLE_NONE K-major core-matrix layout [plane][k/8][voxel][8 ch].  In that layout
//               every one of the 27 filter taps is just a different descriptor START ADDRESS into
//               the same halo box (rows 16 B apart, 8-row groups 160 B apart = next y, k-groups one
//               channel-plane apart), so the box is staged once and read 27 times by the tensor
//               pipe; nothing is re-fetched per tap.
//   B operand   weights pre-packed (rsb_conv3_pack_weights) into the matching core-matrix image so
//               that one (chunk, tap, N-tile) slice is a single contiguous bulk copy
//               (cp.async.bulk -> UBLKCP) completing on an mbarrier; 8-stage ring.
//   MMA         one thread issues tcgen05.mma.cta_group::1.kind::f16 (bf16 x bf16 -> fp32),
//               tcgen05.commit releases smem stages / publishes accumulators.
//   epilogue    4 warps: tcgen05.ld 32x32b.x16, + residual, InstanceNorm (sum, sumsq) via a
//               16-value butterfly over the warp, stores NDHWC rows with a channel pitch.
#include "rsb_common.cuh"

#include "../../include/rsuper_b200.h"

namespace rsb {

constexpr int kTileY = 16;
constexpr int kTileX = 8;
constexpr int kHaloY = kTileY + 2;
constexpr int kHaloX = kTileX + 2;
constexpr int kPlaneVox = kHaloY * kHaloX;        // 180
constexpr int kChunk = 32;                        // channels per K chunk
constexpr int kChunkPlaneBytes = kPlaneVox * 16;  // one 8-channel plane of the halo box: 2880
constexpr int kZPlaneBytes = 4 * kChunkPlaneBytes;  // 11520
constexpr int kNumBStages = 8;
constexpr int kThreads = 448;
constexpr int kProducerWarp0 = 6;
constexpr int kNumProducerWarps = 8;
constexpr int kEpiWarp0 = 2;
constexpr int kMaxNT = 256;

struct Conv3Dev {
  int N, D, H, W, Cin, Cout;
  const void* x;
  long long x_pitch;
  const float* in_stats;
  float eps, slope, inv_count;
  const uint8_t* w_packed;
  void* y;
  long long y_pitch;
  const void* res;
  long long res_pitch;
  float* out_stats;
  const void* mask_x;
  long long mask_x_pitch;
  const float* mask_stats;
  float* bwd_sums;
  // derived
  int NT, ntiles, nchunks, last_ksteps, cout_groups;  // cout_groups = CoutPad / 8
  int tiles_x, tiles_y, zblocks;
  int num_items;
  int acc_stages;
  uint32_t idesc;
  long long* dbg;  // optional per-CTA wait-cycle counters (rsb_debug_set_timing_buffer)
};

struct __align__(16) Conv3Smem {
  // barriers
  uint64_t a_full[2], a_empty[2];
  uint64_t b_full[kNumBStages], b_empty[kNumBStages];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
  uint32_t pad_[3];
  float stat[4][kMaxNT][2];  // per epilogue warp partial (sum, sumsq) / (S1, S2)
  float mstat[kMaxNT][2];    // (mean, rstd) of the masking tensor for the current item
};

template <int PZ>
constexpr int a_unit_bytes() {
  return (PZ + 2) * kZPlaneBytes;
}

struct ItemCoord {
  int n, z0, y0, x0, n0;
};

RSB_DEVICE ItemCoord decode_item(const Conv3Dev& a, int item, int PZ) {
  ItemCoord c;
  int t = item;
  int nt = t % a.ntiles; t /= a.ntiles;
  int xt = t % a.tiles_x; t /= a.tiles_x;
  int yt = t % a.tiles_y; t /= a.tiles_y;
  int zb = t % a.zblocks; t /= a.zblocks;
  c.n = t;
  c.z0 = zb * PZ;
  c.y0 = yt * kTileY;
  c.x0 = xt * kTileX;
  c.n0 = nt * a.NT;
  return c;
}

// Sum 16 per-lane values across the 32 lanes of a warp.  On return, lanes with bit0 == 0 hold the
// total for column butterfly_col(lane) (lanes with bit0 == 1 hold a duplicate).
RSB_DEVICE float butterfly16(float (&v)[16], int lane) {
  float b8[8];
  {
    const bool up = lane & 16;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float send = up ? v[j] : v[j + 8];
      float keep = up ? v[j + 8] : v[j];
      b8[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  float b4[4];
  {
    const bool up = lane & 8;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float send = up ? b8[j] : b8[j + 4];
      float keep = up ? b8[j + 4] : b8[j];
      b4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  float b2[2];
  {
    const bool up = lane & 4;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float send = up ? b4[j] : b4[j + 2];
      float keep = up ? b4[j + 2] : b4[j];
      b2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  float b1;
  {
    const bool up = lane & 2;
    float send = up ? b2[0] : b2[1];
    float keep = up ? b2[1] : b2[0];
    b1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  b1 += __shfl_xor_sync(0xffffffffu, b1, 1);
  return b1;
}
RSB_DEVICE int butterfly_col(int lane) {
  return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}

template <typename T, int PZ>
__global__ void __launch_bounds__(kThreads, 1) conv3_igemm_kernel(const Conv3Dev a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Conv3Smem& sm = *reinterpret_cast<Conv3Smem*>(smem_raw);
  constexpr int kCtrlBytes = (sizeof(Conv3Smem) + 1023) / 1024 * 1024;
  uint8_t* a_buf = smem_raw + kCtrlBytes;                 // 2 units
  uint8_t* b_buf = a_buf + 2 * a_unit_bytes<PZ>();        // kNumBStages stages of NT*64 bytes
  const uint32_t a_base = smem_u32(a_buf);
  const uint32_t b_base = smem_u32(b_buf);
  const uint32_t b_stage_bytes = static_cast<uint32_t>(a.NT) * 64u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---------------- one-time setup ----------------
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&sm.a_full[i]), kNumProducerWarps * 32);
      mbar_init(smem_u32(&sm.a_empty[i]), 1);
      mbar_init(smem_u32(&sm.acc_full[i]), 1);
      mbar_init(smem_u32(&sm.acc_empty[i]), 128);
    }
    for (int i = 0; i < kNumBStages; ++i) {
      mbar_init(smem_u32(&sm.b_full[i]), 1);
      mbar_init(smem_u32(&sm.b_empty[i]), 1);
    }
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < 4 * kMaxNT * 2; i += kThreads) (&sm.stat[0][0][0])[i] = 0.f;
  if (warp == 1) {
    tmem_alloc(smem_u32(&sm.tmem_base), 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = sm.tmem_base;

  const int acc_cols = PZ * a.NT;

  if (warp == 0) {
    // =========================== MMA issuer ===========================
    // The whole warp walks the pipeline (waits are warp-converged); one elected lane issues.  Using
    // elect.sync (instead of `lane == 0`) lets ptxas emit the uniform-datapath UTCHMMA without a
    // per-instruction ELECT/BRA.U.ANY loop.
    {
      uint32_t a_it = 0, b_it = 0, acc_it = 0;
      long long tw_a = 0, tw_b = 0, tw_acc = 0;
      const long long t_begin = clock64();
      const uint32_t a_hi = ((160u >> 4) & 0x3FFFu) | (1u << 14);  // SBO = 160 B (next y row), version 1
      const uint32_t b_hi = ((512u >> 4) & 0x3FFFu) | (1u << 14);  // SBO = 512 B (next 8 couts)
      const uint32_t a_lbo = ((static_cast<uint32_t>(kChunkPlaneBytes) >> 4) & 0x3FFFu) << 16;
      const uint32_t b_lbo = ((128u >> 4) & 0x3FFFu) << 16;
      for (int item = blockIdx.x; item < a.num_items; item += gridDim.x) {
        const uint32_t as = acc_it % a.acc_stages;
        long long tq = clock64();
        mbar_wait(smem_u32(&sm.acc_empty[as]), ((acc_it / a.acc_stages) & 1u) ^ 1u);
        tw_acc += clock64() - tq;
        tc_fence_after_sync();
        uint32_t d_acc[PZ];
#pragma unroll
        for (int p = 0; p < PZ; ++p) d_acc[p] = tmem_base + as * acc_cols + p * a.NT;
        for (int c = 0; c < a.nchunks; ++c) {
          const uint32_t ab = a_it & 1u;
          tq = clock64();
          mbar_wait(smem_u32(&sm.a_full[ab]), (a_it >> 1) & 1u);
          tw_a += clock64() - tq;
          tc_fence_after_sync();
          const bool two_k = (c != a.nchunks - 1) || (a.last_ksteps == 2);
          // Descriptor low words (start address >> 4, LBO in the upper half) are formed once per
          // chunk / per tap; every MMA then costs one integer add per operand: the single issuing
          // thread is the pipeline's critical resource (measured: ~56 cycles per 128xNx16 MMA when
          // the issue loop is tight vs ~160 with per-MMA descriptor arithmetic).
          const uint32_t a_unit_lo = a_lbo | ((a_base + ab * a_unit_bytes<PZ>()) >> 4);
          uint32_t first = (c != 0) ? 1u : 0u;  // accumulate flag of the very first MMA of each accumulator
#pragma unroll
          for (int tap = 0; tap < 27; ++tap) {
            constexpr int kPlaneUnits = kPlaneVox;  // one 8-channel plane = 180 x 16 B
            const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
            const uint32_t bs = b_it % kNumBStages;
            tq = clock64();
            mbar_wait(smem_u32(&sm.b_full[bs]), (b_it / kNumBStages) & 1u);
            tw_b += clock64() - tq;
            tc_fence_after_sync();
            const uint32_t b_lo = b_lbo | ((b_base + bs * b_stage_bytes) >> 4);
            const uint32_t a_tap_lo = a_unit_lo + (kd * 4 * kPlaneUnits + kh * kHaloX + kw);
            if (elect_one()) {
#pragma unroll
            for (int p = 0; p < PZ; ++p) {
              const uint64_t ad0 = (static_cast<uint64_t>(a_hi) << 32) | (a_tap_lo + p * 4 * kPlaneUnits);
              const uint64_t bd0 = (static_cast<uint64_t>(b_hi) << 32) | b_lo;
              umma_bf16_ss(d_acc[p], ad0, bd0, a.idesc, tap == 0 ? first : 1u);
              if (two_k) {
                const uint64_t ad1 = (static_cast<uint64_t>(a_hi) << 32) | (a_tap_lo + (p * 4 + 2) * kPlaneUnits);
                const uint64_t bd1 = (static_cast<uint64_t>(b_hi) << 32) | (b_lo + 16);
                umma_bf16_ss(d_acc[p], ad1, bd1, a.idesc, 1u);
              }
            }
            umma_commit(smem_u32(&sm.b_empty[bs]));
            if (tap == 26) {
              umma_commit(smem_u32(&sm.a_empty[ab]));
              if (c == a.nchunks - 1) umma_commit(smem_u32(&sm.acc_full[as]));
            }
            }
            __syncwarp();
            ++b_it;
          }
          ++a_it;
        }
        ++acc_it;
      }
      if (a.dbg != nullptr && lane == 0) {
        long long* d = a.dbg + static_cast<size_t>(blockIdx.x) * 16;
        d[0] = clock64() - t_begin; d[1] = tw_a; d[2] = tw_b; d[3] = tw_acc; d[4] = acc_it;
      }
    }
  } else if (warp == 1) {
    // =========================== weight loader ===========================
    if (lane == 0) {
      uint32_t b_it = 0;
      long long tw_be = 0;
      for (int item = blockIdx.x; item < a.num_items; item += gridDim.x) {
        const ItemCoord ic = decode_item(a, item, PZ);
        for (int c = 0; c < a.nchunks; ++c) {
          for (int tap = 0; tap < 27; ++tap) {
            const uint32_t bs = b_it % kNumBStages;
            const long long tq = clock64();
            mbar_wait(smem_u32(&sm.b_empty[bs]), ((b_it / kNumBStages) & 1u) ^ 1u);
            tw_be += clock64() - tq;
            const uint8_t* src =
                a.w_packed +
                (static_cast<size_t>(c * 27 + tap) * a.cout_groups + (ic.n0 >> 3)) * 512;
            mbar_arrive_expect_tx(smem_u32(&sm.b_full[bs]), b_stage_bytes);
            bulk_g2s(b_base + bs * b_stage_bytes, src, b_stage_bytes, smem_u32(&sm.b_full[bs]));
            ++b_it;
          }
        }
      }
      if (a.dbg != nullptr) a.dbg[static_cast<size_t>(blockIdx.x) * 16 + 9] = tw_be;
    }
  } else if (warp >= kProducerWarp0) {
    // =========================== A producers ===========================
    // 8 warps.  A thread owns a fixed 8-channel group (cj) and up to 3 in-plane voxel slots; index
    // math is done once per item, then every chunk issues (planes x slots) independent 16/32-byte
    // loads before converting, so one global-memory latency is paid per batch instead of per voxel.
    const int pw = warp - kProducerWarp0;
    const int cj = lane >> 3;  // 8-channel group inside the 32-channel chunk
    const int vi = lane & 7;
    constexpr int kSlots = 3;  // ceil(23 groups of 8 voxels / 8 warps)
    constexpr int NP = PZ + 2;
    constexpr int PB = (sizeof(T) == 2) ? 3 : 2;  // planes per load batch
    const T* __restrict__ xg = reinterpret_cast<const T*>(a.x);
    const long long plane_stride = static_cast<long long>(a.H) * a.W * a.x_pitch;
    const bool has_norm = a.in_stats != nullptr;
    uint32_t a_it = 0;
    long long tw_pe = 0;
    const long long tp_begin = clock64();
    for (int item = blockIdx.x; item < a.num_items; item += gridDim.x) {
      const ItemCoord ic = decode_item(a, item, PZ);
      int soff[kSlots];
      long long goff[kSlots];
      bool ok[kSlots];
#pragma unroll
      for (int s = 0; s < kSlots; ++s) {
        const int vox = (pw + kNumProducerWarps * s) * 8 + vi;
        const int yy = vox / kHaloX, xx = vox - yy * kHaloX;
        const int y = ic.y0 - 1 + yy, xq = ic.x0 - 1 + xx;
        const bool in_tile = vox < kPlaneVox;
        ok[s] = in_tile && y >= 0 && y < a.H && xq >= 0 && xq < a.W;
        goff[s] = (static_cast<long long>(y) * a.W + xq) * a.x_pitch;
        soff[s] = in_tile ? (cj * kPlaneVox + vox) * 16 : -1;
      }
      const long long nbase = static_cast<long long>(ic.n) * a.D * plane_stride;
      for (int c = 0; c < a.nchunks; ++c) {
        const uint32_t ab = a_it & 1u;
        const long long tq = clock64();
        mbar_wait(smem_u32(&sm.a_empty[ab]), ((a_it >> 1) & 1u) ^ 1u);
        tw_pe += clock64() - tq;
        uint8_t* unit = a_buf + ab * a_unit_bytes<PZ>();
        const int ch0 = c * kChunk + cj * 8;
        const bool ch_ok = ch0 < a.Cin;
        float sc[8], sh[8];  // a = act(x * sc + sh)
#pragma unroll
        for (int j = 0; j < 8; ++j) { sc[j] = 1.f; sh[j] = 0.f; }
        if (has_norm && ch_ok) {
          const float* st = a.in_stats + (static_cast<size_t>(ic.n) * a.x_pitch + ch0) * 2;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float mean, rstd;
            stats_to_mean_rstd(st[2 * j], st[2 * j + 1], a.inv_count, a.eps, mean, rstd);
            sc[j] = rstd;
            sh[j] = -mean * rstd;
          }
        }
#pragma unroll
        for (int p0 = 0; p0 < NP; p0 += PB) {
          Raw8<T> raw[PB][kSlots];
          bool inb[PB][kSlots];
#pragma unroll
          for (int pb = 0; pb < PB; ++pb) {
            const int p = p0 + pb;
            const int z = ic.z0 - 1 + p;
            const bool zok = (p < NP) && ch_ok && z >= 0 && z < a.D;
#pragma unroll
            for (int s = 0; s < kSlots; ++s) {
              inb[pb][s] = zok && ok[s];
              if (inb[pb][s]) raw[pb][s].load(xg + nbase + z * plane_stride + goff[s] + ch0);
            }
          }
#pragma unroll
          for (int pb = 0; pb < PB; ++pb) {
            const int p = p0 + pb;
            if (p < NP) {
#pragma unroll
              for (int s = 0; s < kSlots; ++s) {
                if (soff[s] >= 0) {
                  uint4 o = make_uint4(0u, 0u, 0u, 0u);
                  if (inb[pb][s]) {
                    float f[8];
                    raw[pb][s].to_float(f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                      float t = fmaf(f[j], sc[j], sh[j]);
                      if (has_norm) t = t > 0.f ? t : t * a.slope;
                      f[j] = t;
                    }
                    o.x = pack_bf16x2(f[0], f[1]);
                    o.y = pack_bf16x2(f[2], f[3]);
                    o.z = pack_bf16x2(f[4], f[5]);
                    o.w = pack_bf16x2(f[6], f[7]);
                  }
                  *reinterpret_cast<uint4*>(unit + p * kZPlaneBytes + soff[s]) = o;
                }
              }
            }
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(smem_u32(&sm.a_full[ab]));
        ++a_it;
      }
    }
    if (a.dbg != nullptr && pw == 0 && lane == 0) {
      long long* d = a.dbg + static_cast<size_t>(blockIdx.x) * 16;
      d[5] = clock64() - tp_begin; d[6] = tw_pe;
    }
  } else {
    // =========================== epilogue ===========================
    const int ew = warp & 3;  // TMEM lane quarter this warp may access
    const int et = (warp - kEpiWarp0) * 32 + lane;
    const int row = ew * 32 + lane;
    const int ry = row >> 3, rx = row & 7;
    T* __restrict__ yg = reinterpret_cast<T*>(a.y);
    const T* __restrict__ rg = reinterpret_cast<const T*>(a.res);
    const T* __restrict__ mg = reinterpret_cast<const T*>(a.mask_x);
    const bool want_stats = (a.out_stats != nullptr) || (a.bwd_sums != nullptr);
    const bool mask_mode = a.mask_x != nullptr;
    float* stat_dst = mask_mode ? a.bwd_sums : a.out_stats;
    const long long stat_pitch = mask_mode ? a.mask_x_pitch : a.y_pitch;
    uint32_t acc_it = 0;
    long long tw_ef = 0;
    const long long te_begin = clock64();
    for (int item = blockIdx.x; item < a.num_items; item += gridDim.x) {
      const ItemCoord ic = decode_item(a, item, PZ);
      if (mask_mode) {
        named_bar_sync(1, 128);
        for (int cidx = et; cidx < a.NT; cidx += 128) {
          float mean = 0.f, rstd = 1.f;
          if (ic.n0 + cidx < a.Cout) {
            const float* st = a.mask_stats + (static_cast<size_t>(ic.n) * a.mask_x_pitch + ic.n0 + cidx) * 2;
            stats_to_mean_rstd(st[0], st[1], a.inv_count, a.eps, mean, rstd);
          }
          sm.mstat[cidx][0] = mean;
          sm.mstat[cidx][1] = rstd;
        }
        named_bar_sync(1, 128);
      }
      const uint32_t as = acc_it % a.acc_stages;
      const long long tq = clock64();
      mbar_wait(smem_u32(&sm.acc_full[as]), (acc_it / a.acc_stages) & 1u);
      tw_ef += clock64() - tq;
      tc_fence_after_sync();
      const int y = ic.y0 + ry, xq = ic.x0 + rx;
      const bool row_ok = (y < a.H) && (xq < a.W);
      // flattened (plane, 16-column chunk) iteration space; the residual / mask rows of iteration
      // i+1 are requested before the TMEM load of iteration i, so their latency overlaps the math.
      const int nch = a.NT >> 4;
      const int nplanes = min(PZ, a.D - ic.z0);
      const int nit = nplanes * nch;
      const size_t vox0 = ((static_cast<size_t>(ic.n) * a.D + ic.z0) * a.H + (row_ok ? y : 0)) * a.W + (row_ok ? xq : 0);
      const size_t vox_plane = static_cast<size_t>(a.H) * a.W;
      Raw16<T> pre_res, pre_mx;
      pre_res.zero();
      pre_mx.zero();
      auto prefetch = [&](int it) {
        const int p = it / nch;
        const int cbase = ic.n0 + (it - p * nch) * 16;
        const int nvalid = a.Cout - cbase;
        if (row_ok && nvalid > 0) {
          const size_t vox = vox0 + p * vox_plane;
          if (rg != nullptr) pre_res.load(rg + vox * a.res_pitch + cbase, nvalid > 8);
          if (mask_mode) pre_mx.load(mg + vox * a.mask_x_pitch + cbase, nvalid > 8);
        }
      };
      if (nit > 0) prefetch(0);
#pragma unroll 1
      for (int it = 0; it < nit; ++it) {
        const int p = it / nch;
        const int cc = (it - p * nch) * 16;
        const Raw16<T> cur_res = pre_res, cur_mx = pre_mx;
        if (it + 1 < nit) prefetch(it + 1);
        const size_t vox = vox0 + p * vox_plane;
        uint32_t r[16];
        __syncwarp();
        tmem_ld16(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * acc_cols + p * a.NT + cc, r);
        tmem_ld_wait();
        const int cbase = ic.n0 + cc;
        const int nvalid = a.Cout - cbase;  // multiple of 8 (Cout % 8 == 0)
        if (nvalid <= 0) continue;
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
        float xh[16];
        if (row_ok) {
          if (rg != nullptr) {
            float t[16];
            cur_res.to_float(t);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += t[j];
          }
          if (mask_mode) {
            cur_mx.to_float(xh);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float h = (xh[j] - sm.mstat[cc + j][0]) * sm.mstat[cc + j][1];
              xh[j] = h;
              v[j] = h > 0.f ? v[j] : v[j] * a.slope;
            }
          }
        }
        if (want_stats) {
          float s1[16], s2[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float val = row_ok ? v[j] : 0.f;
            s1[j] = val;
            s2[j] = mask_mode ? (row_ok ? val * xh[j] : 0.f) : val * val;
          }
          const float t1 = butterfly16(s1, lane);
          const float t2 = butterfly16(s2, lane);
          if ((lane & 1) == 0) {
            const int col = cc + butterfly_col(lane);
            sm.stat[warp - kEpiWarp0][col][0] += t1;
            sm.stat[warp - kEpiWarp0][col][1] += t2;
          }
        }
        if (row_ok) {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = v[j];
          Vec8<T>::store(yg + vox * a.y_pitch + cbase, o);
          if (nvalid > 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = v[8 + j];
            Vec8<T>::store(yg + vox * a.y_pitch + cbase + 8, o);
          }
        }
      }
      tc_fence_before_sync();
      mbar_arrive(smem_u32(&sm.acc_empty[as]));
      ++acc_it;
      if (want_stats) {
        __syncwarp();
        for (int col = lane; col < a.NT; col += 32) {
          if (ic.n0 + col < a.Cout) {
            float* dst = stat_dst + (static_cast<size_t>(ic.n) * stat_pitch + ic.n0 + col) * 2;
            atomicAdd(dst, sm.stat[warp - kEpiWarp0][col][0]);
            atomicAdd(dst + 1, sm.stat[warp - kEpiWarp0][col][1]);
          }
          sm.stat[warp - kEpiWarp0][col][0] = 0.f;
          sm.stat[warp - kEpiWarp0][col][1] = 0.f;
        }
        __syncwarp();
      }
    }
    if (a.dbg != nullptr && warp == kEpiWarp0 && lane == 0) {
      long long* d = a.dbg + static_cast<size_t>(blockIdx.x) * 16;
      d[7] = clock64() - te_begin; d[8] = tw_ef;
    }
  }

  // ---------------- teardown ----------------
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// weight packing: fp32 OIDHW -> bf16 [chunk][tap][Cout'/8][4 (k/8)][8 (o%8)][8 (k%8)]
// ---------------------------------------------------------------------------------------------
__global__ void pack_conv3_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                          int Cout, int Cin, int transpose_flip, int co_eff, int ci_eff,
                                          int co_groups, size_t total) {
  size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  size_t t = i;
  const int kk = t & 7; t >>= 3;
  const int orow = t & 7; t >>= 3;
  const int kc = t & 3; t >>= 2;
  const int og = t % co_groups; t /= co_groups;
  const int tap = t % 27; t /= 27;
  const int chunk = static_cast<int>(t);
  const int o = og * 8 + orow;
  const int k = chunk * kChunk + kc * 8 + kk;
  float v = 0.f;
  if (o < co_eff && k < ci_eff) {
    if (!transpose_flip) {
      v = w[(static_cast<size_t>(o) * Cin + k) * 27 + tap];
    } else {
      // effective conv: out channel o = original ci, in channel k = original co, taps flipped
      v = w[(static_cast<size_t>(k) * Cin + o) * 27 + (26 - tap)];
    }
  }
  out[i] = __float2bfloat16_rn(v);
}

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

template <typename T, int PZ>
static int launch_conv3(const Conv3Dev& dev, int grid, size_t smem_bytes, cudaStream_t stream) {
  auto kern = conv3_igemm_kernel<T, PZ>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem_bytes));
  if (e != cudaSuccess) {
    set_last_error("conv3: cudaFuncSetAttribute(%zu B smem) failed: %s", smem_bytes, cudaGetErrorString(e));
    return -2;
  }
  kern<<<grid, kThreads, smem_bytes, stream>>>(dev);
  return check_launch("conv3_igemm_kernel");
}

}  // namespace rsb

using namespace rsb;

static long long* g_timing_buffer = nullptr;
// Bring-up / profiling aid: when set (device pointer to >= 16 * grid int64), conv3_forward launches
// record per-CTA cycle counters: [0] MMA thread total, [1..3] its waits on a_full / b_full / acc_empty,
// [4] items, [5] producer total, [6] producer wait on a_empty, [7] epilogue total, [8] epilogue wait on
// acc_full, [9] weight-loader wait on b_empty.
extern "C" int rsb_debug_set_timing_buffer(void* device_ptr) {
  g_timing_buffer = reinterpret_cast<long long*>(device_ptr);
  return 0;
}

extern "C" size_t rsb_conv3_packed_weight_bytes(int Cout, int Cin) {
  const int nchunks = (Cin + kChunk - 1) / kChunk;
  const int co_pad = round_up(Cout, 16);
  return static_cast<size_t>(nchunks) * 27 * co_pad * 64;
}

extern "C" int rsb_conv3_pack_weights(const float* w_oidhw, void* packed, int Cout, int Cin,
                                      int transpose_flip, void* stream) {
  RSB_REQUIRE(w_oidhw && packed, "pack_weights: null pointer");
  RSB_REQUIRE(Cout > 0 && Cin > 0, "pack_weights: bad channel counts");
  const int co_eff = transpose_flip ? Cin : Cout;
  const int ci_eff = transpose_flip ? Cout : Cin;
  const int co_groups = round_up(co_eff, 16) / 8;
  const size_t total = rsb_conv3_packed_weight_bytes(co_eff, ci_eff) / 2;
  const int threads = 256;
  const unsigned blocks = static_cast<unsigned>((total + threads - 1) / threads);
  pack_conv3_weights_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      w_oidhw, reinterpret_cast<__nv_bfloat16*>(packed), Cout, Cin, transpose_flip, co_eff, ci_eff,
      co_groups, total);
  return check_launch("pack_conv3_weights_kernel");
}

extern "C" int rsb_conv3_forward(const RsbConv3Args* p, void* stream) {
  RSB_REQUIRE(p != nullptr, "conv3: null args");
  RSB_REQUIRE(p->x && p->y && p->w_packed, "conv3: null tensor pointer");
  RSB_REQUIRE(p->N > 0 && p->D > 0 && p->H > 0 && p->W > 0, "conv3: bad geometry");
  RSB_REQUIRE(p->Cin > 0 && p->Cin % 8 == 0, "conv3: Cin must be a positive multiple of 8 (got %d)", p->Cin);
  RSB_REQUIRE(p->Cout > 0 && p->Cout % 8 == 0, "conv3: Cout must be a positive multiple of 8 (got %d)", p->Cout);
  RSB_REQUIRE(p->dtype == RSB_BF16 || p->dtype == RSB_F32, "conv3: bad dtype %d", p->dtype);
  RSB_REQUIRE(p->x_pitch >= p->Cin && p->x_pitch % 8 == 0, "conv3: bad x_pitch %d", p->x_pitch);
  RSB_REQUIRE(p->y_pitch >= p->Cout && p->y_pitch % 8 == 0, "conv3: bad y_pitch %d", p->y_pitch);
  RSB_REQUIRE(!p->res || (p->res_pitch >= p->Cout && p->res_pitch % 8 == 0), "conv3: bad res_pitch");
  RSB_REQUIRE(!p->mask_x || (p->mask_stats && p->bwd_sums && p->mask_x_pitch % 8 == 0),
              "conv3: mask_x needs mask_stats and bwd_sums");
  RSB_REQUIRE(!(p->mask_x && p->out_stats), "conv3: out_stats and mask epilogue are exclusive");

  Conv3Dev d{};
  d.N = p->N; d.D = p->D; d.H = p->H; d.W = p->W; d.Cin = p->Cin; d.Cout = p->Cout;
  d.x = p->x; d.x_pitch = p->x_pitch; d.in_stats = p->in_stats;
  d.eps = p->eps; d.slope = p->slope;
  d.inv_count = 1.0f / (static_cast<float>(p->D) * p->H * p->W);
  d.w_packed = reinterpret_cast<const uint8_t*>(p->w_packed);
  d.y = p->y; d.y_pitch = p->y_pitch; d.res = p->res; d.res_pitch = p->res_pitch;
  d.out_stats = p->out_stats; d.mask_x = p->mask_x; d.mask_x_pitch = p->mask_x_pitch;
  d.mask_stats = p->mask_stats; d.bwd_sums = p->bwd_sums;
  d.dbg = g_timing_buffer;

  const int co_pad = round_up(p->Cout, 16);
  d.cout_groups = co_pad / 8;
  d.nchunks = (p->Cin + kChunk - 1) / kChunk;
  const int cin_pad16 = round_up(p->Cin, 16);
  d.last_ksteps = (cin_pad16 % kChunk == 16) ? 1 : 2;
  d.tiles_x = (p->W + kTileX - 1) / kTileX;
  d.tiles_y = (p->H + kTileY - 1) / kTileY;

  int sms = p->max_ctas > 0 ? p->max_ctas : rsb_num_sms();
  RSB_REQUIRE(sms > 0, "conv3: could not query the SM count");

  // N tile: largest divisor of co_pad that is a multiple of 16 and <= cap
  auto pick_nt = [&](int cap) {
    int best = 16;
    for (int nt = 16; nt <= cap && nt <= co_pad; nt += 16)
      if (co_pad % nt == 0) best = nt;
    return best;
  };
  int PZ = p->planes_per_item;
  int NT = p->n_tile;
  if (PZ == 0) {
    // enough z-blocks to give every SM at least ~2 items, otherwise fewer planes per item
    PZ = 4;
    while (PZ > 1) {
      const int nt = NT ? NT : pick_nt(PZ == 4 ? 128 : 256);
      const long long items = static_cast<long long>(p->N) * ((p->D + PZ - 1) / PZ) * d.tiles_y *
                              d.tiles_x * (co_pad / nt);
      if (PZ <= p->D && items >= 2LL * sms) break;
      PZ >>= 1;
    }
  }
  RSB_REQUIRE(PZ == 1 || PZ == 2 || PZ == 4, "conv3: planes_per_item must be 1, 2 or 4 (got %d)", PZ);
  if (NT == 0) NT = pick_nt(PZ == 4 ? 128 : 256);
  RSB_REQUIRE(NT % 16 == 0 && NT >= 16 && NT <= kMaxNT && co_pad % NT == 0,
              "conv3: n_tile %d must be a multiple of 16 dividing %d", NT, co_pad);
  RSB_REQUIRE(PZ * NT <= 512, "conv3: PZ*n_tile = %d exceeds the 512 TMEM columns", PZ * NT);
  d.NT = NT;
  d.ntiles = co_pad / NT;
  d.zblocks = (p->D + PZ - 1) / PZ;
  d.acc_stages = (2 * PZ * NT <= 512) ? 2 : 1;
  const long long items = static_cast<long long>(p->N) * d.zblocks * d.tiles_y * d.tiles_x * d.ntiles;
  RSB_REQUIRE(items < (1LL << 31), "conv3: too many work items");
  d.num_items = static_cast<int>(items);
  d.idesc = make_idesc_bf16(128, NT, 0, 0);

  constexpr size_t ctrl = (sizeof(Conv3Smem) + 1023) / 1024 * 1024;
  size_t smem = ctrl + 2 * static_cast<size_t>(PZ + 2) * kZPlaneBytes + static_cast<size_t>(kNumBStages) * NT * 64;
  if (smem < 120 * 1024) smem = 120 * 1024;  // force 1 CTA / SM: each CTA owns all 512 TMEM columns
  RSB_REQUIRE(smem <= 227 * 1024, "conv3: shared memory budget exceeded (%zu B)", smem);
  const int grid = static_cast<int>(items < sms ? items : sms);
  cudaStream_t st = static_cast<cudaStream_t>(stream);

#define RSB_DISPATCH(TT)                                                   \
  switch (PZ) {                                                            \
    case 1: return launch_conv3<TT, 1>(d, grid, smem, st);                 \
    case 2: return launch_conv3<TT, 2>(d, grid, smem, st);                 \
    default: return launch_conv3<TT, 4>(d, grid, smem, st);                \
  }
  if (p->dtype == RSB_BF16) {
    RSB_DISPATCH(__nv_bfloat16)
  } else {
    RSB_DISPATCH(float)
  }
#undef RSB_DISPATCH
}
