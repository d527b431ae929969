// conv3_stream.cu — plane-streaming variant of the 3x3x3 implicit-GEMM convolution for the 32-channel layers
// (Cin <= 32, 16 < Cout <= 32: the full-resolution level of the UNet, 10 of the 68 fprop / dgrad launches per step and
// a quarter of their time).  Same math, operands, packed weights and epilogue contract as conv3_fprop.cu
// (nn.Conv3d of ConvNormAct, rsuper_train/model/dim3/conv_layers.py:29-49, residual add :92, next InstanceNorm's
// statistics; with flipped weights the dgrad with the act'(xhat) mask and the InstanceNorm-backward sums).
//
// Why a second kernel: below N = 128 a tcgen05.mma (M = 128, K = 16, operands in shared memory) costs a flat 56-60
// cycles — the A operand fetch — so a 32-channel layer lives or dies by how many output planes one MMA serves.  The
// item-based kernel merges the three kd taps inside a block of PZ = 4 output planes: 6 MMAs per (tap, k-step) for 4
// planes (N = 32, 64, 96, 96, 64, 32), i.e. 360 cycles for 192 cycles of math.  Here a CTA walks a z-column instead:
//   * the packed weights of the whole layer (27 taps x 3 kd x 32 x 32 bf16 = 54 KB) stay resident in shared memory;
//   * input planes stream through a 12-slot ring (one 5-D TMA box (32 ch, 10 x, 18 y, 1 z) each, 64-byte swizzled rows);
//   * the accumulators of 16 consecutive output planes form a ring of 16 x 32 TMEM columns; input plane i feeds output
//     planes i-2, i-1, i (kd = 2, 1, 0) = three ADJACENT column blocks, so every (tap, k-step) is ONE N = 96 MMA
//     (two at the ring wrap, plus the split first touch of the new accumulator): 19 MMAs per plane instead of 27;
//   * output plane o is complete after input plane o + 2; eight epilogue warps drain it while the MMAs run 13 planes ahead.
#include "rsb_common.cuh"
#include "rsb_tma.cuh"

#include <cstdlib>

#include "../../include/rsuper_b200.h"

namespace rsb {

constexpr int kStThreads = 384;          // warp 0: TMA producer | warp 1: MMA issuer | warp 2: TMEM owner + weights | warps 4-11: epilogue
constexpr int kStEpiWarp0 = 4;
constexpr int kStEpiWarps = 8;
constexpr int kStNT = 32;                // N tile = padded Cout
constexpr int kStAcc = 16;               // accumulator ring: 16 x 32 = 512 TMEM columns
constexpr int kStRing = 12;              // input-plane ring
constexpr int kStPlaneBytes = 180 * 64;  // haloed plane: 18 x 10 voxels x 32 channels
constexpr int kStSlotBytes = 12288;      // ring stride (512-byte multiple: the swizzle is a function of address bits)
constexpr int kStTapBytes = 3 * kStNT * 64;       // one (kh, kw) tap: [kd=2 | kd=1 | kd=0][32 co][32 k]
constexpr int kStWeightBytes = 9 * kStTapBytes;   // 55296
constexpr int kStChunk = 32;             // output planes per work unit (needs 34 input planes)

struct StreamDev {
  int N, D, H, W, Cin, Cout;
  const uint8_t* w_packed;
  void* y;
  long long y_pitch;
  const void* aux;
  long long aux_pitch;
  int mask_mode;
  const float* mask_stats;
  float* stat_dst;
  long long stat_pitch;
  float eps, slope, inv_count;
  int ksteps;  // 1 if Cin <= 16 else 2
  int tiles_x, tiles_y, zchunks, num_units;
};

struct __align__(16) StreamSmem {
  uint64_t a_full[kStRing], a_empty[kStRing];
  uint64_t acc_full[kStAcc], acc_empty[kStAcc];
  uint64_t b_full;
  uint32_t tmem_base;
  uint32_t pad_;
  float stat[kStEpiWarps][16][2];
  float mstat[kStNT][2];
};
constexpr int kStCtrlBytes = (sizeof(StreamSmem) + 1023) / 1024 * 1024;

struct StUnit {
  int n, y0, x0, z0, L;
};
RSB_DEVICE StUnit st_decode(const StreamDev& a, int u) {
  StUnit r;
  int t = u;
  const int zc = t % a.zchunks; t /= a.zchunks;
  const int xt = t % a.tiles_x; t /= a.tiles_x;
  const int yt = t % a.tiles_y; t /= a.tiles_y;
  r.n = t;
  r.y0 = yt * 16;
  r.x0 = xt * 8;
  r.z0 = zc * kStChunk;
  r.L = min(kStChunk, a.D - r.z0);
  return r;
}

RSB_DEVICE float st_butterfly16(float (&v)[16], int lane) {
  float b8[8], b4[4], b2[2];
  {
    const bool up = lane & 16;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float send = up ? v[j] : v[j + 8], keep = up ? v[j + 8] : v[j];
      b8[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool up = lane & 8;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float send = up ? b8[j] : b8[j + 4], keep = up ? b8[j + 4] : b8[j];
      b4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool up = lane & 4;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float send = up ? b4[j] : b4[j + 2], keep = up ? b4[j + 2] : b4[j];
      b2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  const bool up = lane & 2;
  const float send = up ? b2[0] : b2[1], keep = up ? b2[1] : b2[0];
  float b1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  b1 += __shfl_xor_sync(0xffffffffu, b1, 1);
  return b1;
}
RSB_DEVICE int st_butterfly_col(int lane) {
  return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}

template <typename T>
__global__ void __launch_bounds__(kStThreads, 1)
conv3_stream32_kernel(const __grid_constant__ CUtensorMap tm_a, const StreamDev a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  StreamSmem& sm = *reinterpret_cast<StreamSmem*>(smem_raw);
  const uint32_t a_base = smem_u32(smem_raw) + kStCtrlBytes;     // kStRing plane slots
  const uint32_t b_base = a_base + kStRing * kStSlotBytes;       // resident packed weights
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStRing; ++i) {
      mbar_init(smem_u32(&sm.a_full[i]), 1);
      mbar_init(smem_u32(&sm.a_empty[i]), 1);
    }
    for (int i = 0; i < kStAcc; ++i) {
      mbar_init(smem_u32(&sm.acc_full[i]), 1);
      mbar_init(smem_u32(&sm.acc_empty[i]), kStEpiWarps * 32);
    }
    mbar_init(smem_u32(&sm.b_full), 1);
    mbar_fence_init();
    tma_prefetch_desc(&tm_a);
  }
  for (int i = threadIdx.x; i < kStEpiWarps * 16 * 2; i += kStThreads) (&sm.stat[0][0][0])[i] = 0.f;
  if (warp == 2) {
    tmem_alloc(smem_u32(&sm.tmem_base), 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = sm.tmem_base;

  if (warp == 2) {
    // =========================== resident weights (once) ===========================
    if (elect_one()) {
      const uint32_t bar = smem_u32(&sm.b_full);
      mbar_arrive_expect_tx(bar, kStWeightBytes);
      for (int t = 0; t < 9; ++t) bulk_g2s(b_base + t * kStTapBytes, a.w_packed + static_cast<size_t>(t) * kStTapBytes, kStTapBytes, bar);
    }
    __syncwarp();
  } else if (warp == 0) {
    // =========================== input-plane producer (TMA) ===========================
    uint32_t slot = 0, ph = 0;
    for (int u = blockIdx.x; u < a.num_units; u += gridDim.x) {
      const StUnit un = st_decode(a, u);
      for (int i = 0; i < un.L + 2; ++i) {
        mbar_wait(smem_u32(&sm.a_empty[slot]), ph ^ 1u);
        if (elect_one()) {
          const uint32_t bar = smem_u32(&sm.a_full[slot]);
          mbar_arrive_expect_tx(bar, kStPlaneBytes);
          tma_load_5d(a_base + slot * kStSlotBytes, &tm_a, 0, un.x0 - 1, un.y0 - 1, un.z0 - 1 + i, un.n, bar);
        }
        __syncwarp();
        if (++slot == kStRing) { slot = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t slot16 = (kStNT / 8) * 512u / 16u;   // one kd slot of a tap, in 16-byte units
    constexpr uint32_t tap16 = kStTapBytes / 16;
    const uint32_t a_hi = desc_hi(640, kLayoutSw64);   // 8-row groups (8 x) are one y row = 10 voxel rows apart
    const uint32_t b_hi = desc_hi(512, kLayoutNone);   // next 8 couts
    constexpr uint32_t a_lbo = 1u << 16;
    constexpr uint32_t b_lbo = ((128u >> 4) & 0x3FFFu) << 16;
    const uint32_t b_lo0 = b_lbo | (b_base >> 4);
    const uint32_t bar_a_full = smem_u32(&sm.a_full[0]), bar_a_empty = smem_u32(&sm.a_empty[0]);
    const uint32_t bar_acc_full = smem_u32(&sm.acc_full[0]), bar_acc_empty = smem_u32(&sm.acc_empty[0]);
    const int ksteps = a.ksteps;
    // (Round 2 tried the opposite organisation — the whole warp walks the schedule with warp-uniform control flow and only the
    // tcgen05 instructions under the elected lane, to trade the R2UR moves of this divergent region for uniform-datapath
    // arithmetic: the compiler still hoisted the 17 B-descriptor words into vector registers, and 32 lanes polling the
    // barriers cost more than one: 239 vs 211 us standalone on the same box.  Per plane the issuer spends ~1630 cycles for
    // 1140 of MMAs; ~2 x 90 of the difference are the two barrier polls.)
    // ONE elected thread runs the whole issue loop.  The tensor pipe queues only ~2 MMAs (~110 cycles of cover), so
    // everything this thread does between two MMAs that is not hidden behind them is pipe idle time: the barrier waits
    // of plane i + 1 are therefore taken in the MIDDLE of plane i's MMA stream (they are satisfied long before: the
    // TMA ring runs 12 planes ahead, the accumulator ring has 13 planes of slack), and a regular plane needs no
    // arithmetic beyond its accumulator column.
    if (elect_one()) {
      mbar_wait(smem_u32(&sm.b_full), 0);
      const bool two_k = ksteps == 2;
      const uint32_t i64 = make_idesc_bf16(128, 2 * kStNT, 0, 0), i32 = make_idesc_bf16(128, kStNT, 0, 0);
      const uint32_t i96 = make_idesc_bf16(128, 3 * kStNT, 0, 0);
      uint32_t slot = 0, ph = 0;   // input ring position of the CURRENT plane
      uint32_t oc = 0;             // outputs started before the current unit: accumulator slot = (oc + o) & 15
      int u = blockIdx.x, i = 0;
      int L = u < a.num_units ? st_decode(a, u).L : 0;
      auto wait_plane = [&](int wi, int wL, uint32_t woc, uint32_t wslot, uint32_t wph) {
        if (wi < wL) {
          const uint32_t on = woc + static_cast<uint32_t>(wi);
          mbar_wait(bar_acc_empty + 8u * (on & 15u), ((on >> 4) & 1u) ^ 1u);
        }
        mbar_wait(bar_a_full + 8u * wslot, wph);
        tc_fence_after_sync();
      };
      if (u < a.num_units) wait_plane(0, L, 0u, 0u, 0u);
      while (u < a.num_units) {
        // the plane after this one (possibly the first plane of the next unit)
        int nu = u, ni = i + 1, nL = L;
        uint32_t noc = oc;
        if (ni == L + 2) {
          nu = u + gridDim.x; ni = 0; noc = oc + static_cast<uint32_t>(L);
          nL = nu < a.num_units ? st_decode(a, nu).L : 0;
        }
        uint32_t nslot = slot + 1, nph = ph;
        if (nslot == kStRing) { nslot = 0; nph ^= 1u; }
        const bool has_next = nu < a.num_units;

        const uint32_t a_lo0 = a_lbo | ((a_base + slot * kStSlotBytes) >> 4);
        const uint32_t s_open = (oc + static_cast<uint32_t>(i)) & 15u;
        const int o_lo = i - 2 > 0 ? i - 2 : 0;
        const int o_hi = i < L - 1 ? i : L - 1;
        const bool opens = i < L;
        if (i >= 2 && opens && s_open >= 2u) {
          // ---- regular plane (26 of 34): outputs i-2, i-1 accumulate, output i opens, no ring wrap: the three
          // accumulators are the columns [d, d + 96)
          const uint32_t d = tmem_base + (s_open - 2u) * kStNT;
          const uint64_t ad0 = desc_join(a_hi, a_lo0);
          umma_bf16_ss(d, ad0, desc_join(b_hi, b_lo0), i64, 1u);                                    // kd = 2, 1
          umma_bf16_ss(d + 2 * kStNT, ad0, desc_join(b_hi, b_lo0 + 2u * slot16), i32, 0u);          // kd = 0: first touch
#pragma unroll
          for (int tk = 1; tk < 18; ++tk) {
            const int t = tk >> 1, ks = tk & 1;
            if (tk == 9 && has_next) wait_plane(ni, nL, noc, nslot, nph);
            if (ks == 0 || two_k) {
              const int kh = t / 3, kw = t - kh * 3;
              umma_bf16_ss(d, desc_join(a_hi, a_lo0 + static_cast<uint32_t>((kh * 10 + kw) * 4 + ks * 2)),
                           desc_join(b_hi, b_lo0 + static_cast<uint32_t>(t) * tap16 + static_cast<uint32_t>(ks) * 16u), i96, 1u);
            }
          }
        } else {
          // Steady runs: outputs that plane i only ACCUMULATES into, as at most two column-contiguous runs (the
          // accumulator ring wraps every 16 output planes).  Everything is kept in scalars: indexed arrays end up in
          // local memory and this warp is instruction-bound (a list-based version executed ~980 instructions per plane
          // for 19 MMAs — ncu: R2UR / LDL / predicate logic — and ran at 4000 cycles per plane instead of 1140).
          const int acc_hi = opens ? o_hi - 1 : o_hi;      // last output that was already opened by an earlier plane
          uint32_t d0 = 0, b0 = 0, i0 = 0, d1 = 0, b1 = 0, i1 = 0;
          int nseg = 0, n0 = 0;
          if (acc_hi >= o_lo) {
            const uint32_t s_lo = (oc + o_lo) & 15u, s_hi = (oc + acc_hi) & 15u;
            d0 = tmem_base + s_lo * kStNT;
            b0 = static_cast<uint32_t>(2 - (i - o_lo)) * slot16;
            if (s_lo <= s_hi) {
              n0 = acc_hi - o_lo + 1;
              nseg = 1;
            } else {
              n0 = 16 - static_cast<int>(s_lo);             // planes before the wrap
              d1 = tmem_base;
              b1 = static_cast<uint32_t>(2 - (i - (o_lo + n0))) * slot16;
              i1 = make_idesc_bf16(128, (acc_hi - o_lo + 1 - n0) * kStNT, 0, 0);
              nseg = 2;
            }
            i0 = make_idesc_bf16(128, n0 * kStNT, 0, 0);
          }
          const uint32_t open_slot = (oc + static_cast<uint32_t>(i)) & 15u;
          const uint32_t open_d = tmem_base + open_slot * kStNT;
          const uint32_t open_b = 2u * slot16;              // kd = 0 slot
          const uint32_t idesc1 = make_idesc_bf16(128, kStNT, 0, 0);
          // first (tap 0, k-step 0): the steady runs accumulate, the opened accumulator is overwritten
          {
            const uint64_t ad = desc_join(a_hi, a_lo0);
            if (nseg >= 1) umma_bf16_ss(d0, ad, desc_join(b_hi, b_lo0 + b0), i0, 1u);
            if (nseg == 2) umma_bf16_ss(d1, ad, desc_join(b_hi, b_lo0 + b1), i1, 1u);
            if (opens) umma_bf16_ss(open_d, ad, desc_join(b_hi, b_lo0 + open_b), idesc1, 0u);
          }
          // all other (tap, k-step): if the opened accumulator directly follows the last steady run (no ring wrap in
          // between) one wider instruction covers both
          const bool merged = opens && nseg >= 1 && open_slot != 0u;
          if (merged) {
            if (nseg == 1) i0 = make_idesc_bf16(128, (n0 + 1) * kStNT, 0, 0);
            else i1 = make_idesc_bf16(128, (acc_hi - o_lo + 1 - n0 + 1) * kStNT, 0, 0);
          }
          const bool sep_open = opens && !merged;
          const bool two = ksteps == 2;
          if (nseg + (sep_open ? 1 : 0) == 1) {
            // ---- common case: ONE instruction per (tap, k-step); fully unrolled, descriptors = base + immediate ----
            const uint32_t dd = nseg == 1 ? d0 : open_d;
            const uint32_t bb = b_lo0 + (nseg == 1 ? b0 : open_b);
            const uint32_t ii = nseg == 1 ? i0 : idesc1;
#pragma unroll
            for (int tk = 1; tk < 18; ++tk) {
              const int t = tk >> 1, ks = tk & 1;
              if (ks == 0 || two) {
                const int kh = t / 3, kw = t - kh * 3;
                umma_bf16_ss(dd, desc_join(a_hi, a_lo0 + static_cast<uint32_t>((kh * 10 + kw) * 4 + ks * 2)),
                             desc_join(b_hi, bb + static_cast<uint32_t>(t) * tap16 + static_cast<uint32_t>(ks) * 16u), ii, 1u);
              }
            }
          } else {
            // ---- ring wrap / separately opened accumulator (2 planes in 16): up to three instructions per (tap, k-step) ----
#pragma unroll 1
            for (int tk = 1; tk < 18; ++tk) {
              const int t = tk >> 1, ks = tk & 1;
              if (ks == 0 || two) {
                const int kh = t / 3, kw = t - kh * 3;
                const uint64_t ad = desc_join(a_hi, a_lo0 + static_cast<uint32_t>((kh * 10 + kw) * 4 + ks * 2));
                const uint32_t b_ks = b_lo0 + static_cast<uint32_t>(t) * tap16 + static_cast<uint32_t>(ks) * 16u;
                if (nseg >= 1) umma_bf16_ss(d0, ad, desc_join(b_hi, b_ks + b0), i0, 1u);
                if (nseg == 2) umma_bf16_ss(d1, ad, desc_join(b_hi, b_ks + b1), i1, 1u);
                if (sep_open) umma_bf16_ss(open_d, ad, desc_join(b_hi, b_ks + open_b), idesc1, 1u);
              }
            }
          }
          if (has_next) wait_plane(ni, nL, noc, nslot, nph);
        }
        umma_commit(bar_a_empty + 8u * slot);
        if (i >= 2) umma_commit(bar_acc_full + 8u * ((oc + static_cast<uint32_t>(i - 2)) & 15u));   // output i - 2 is complete
        u = nu; i = ni; L = nL; oc = noc; slot = nslot; ph = nph;
      }
    }
    __syncwarp();
  } else if (warp >= kStEpiWarp0) {
    // =========================== epilogue (8 warps) ===========================
    const int ew = warp & 3;                       // TMEM lane quarter this warp may access
    const int eidx = warp - kStEpiWarp0;
    const int eset = eidx >> 2;                    // column half: channels [16 * eset, 16 * eset + 16)
    const int et = eidx * 32 + lane;
    const int row = ew * 32 + lane;
    const int ry = row >> 3, rx = row & 7;
    const int cc = eset * 16;
    T* __restrict__ yg = reinterpret_cast<T*>(a.y);
    const T* __restrict__ xg = reinterpret_cast<const T*>(a.aux);
    const bool want_stats = a.stat_dst != nullptr;
    const bool mask_mode = a.mask_mode != 0;
    const bool has_aux = a.aux != nullptr;
    const int nvalid = a.Cout - cc;                // channels of this half that exist (multiple of 8; may be <= 0)
    float s1[16], s2[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
    int stat_n = -1;
    auto flush_stats = [&]() {
      // registers -> warp butterfly -> this warp's shared partial -> one round of global atomics per sample
      const float t1 = st_butterfly16(s1, lane), t2 = st_butterfly16(s2, lane);
      if ((lane & 1) == 0) {
        const int col = st_butterfly_col(lane);
        sm.stat[eidx][col][0] = t1;
        sm.stat[eidx][col][1] = t2;
      }
      __syncwarp();
      if (lane < 16 && cc + lane < a.Cout) {
        float* dst = a.stat_dst + (static_cast<size_t>(stat_n) * a.stat_pitch + cc + lane) * 2;
        atomicAdd(dst, sm.stat[eidx][lane][0]);
        atomicAdd(dst + 1, sm.stat[eidx][lane][1]);
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 16; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
    };
    uint32_t oc = 0;
    for (int u = blockIdx.x; u < a.num_units; u += gridDim.x) {
      const StUnit un = st_decode(a, u);
      if (un.n != stat_n) {
        if (want_stats && stat_n >= 0) flush_stats();
        stat_n = un.n;
        if (mask_mode) {   // (mean, rstd) of the masking tensor change with the sample only
          named_bar_sync(1, kStEpiWarps * 32);
          for (int cidx = et; cidx < kStNT; cidx += kStEpiWarps * 32) {
            float mean = 0.f, rstd = 1.f;
            if (cidx < a.Cout) {
              const float* st = a.mask_stats + (static_cast<size_t>(un.n) * a.aux_pitch + cidx) * 2;
              stats_to_mean_rstd(st[0], st[1], a.inv_count, a.eps, mean, rstd);
            }
            sm.mstat[cidx][0] = mean;
            sm.mstat[cidx][1] = rstd;
          }
          named_bar_sync(1, kStEpiWarps * 32);
        }
      }
      const int y = un.y0 + ry, xq = un.x0 + rx;
      const bool row_ok = (y < a.H) && (xq < a.W) && nvalid > 0;
      const size_t vox0 = ((static_cast<size_t>(un.n) * a.D + un.z0) * a.H + (row_ok ? y : 0)) * a.W + (row_ok ? xq : 0);
      const size_t vox_plane = static_cast<size_t>(a.H) * a.W;
      auto arm = [&](int o, Raw16<T>& dst) {
        dst.zero();
        if (has_aux && row_ok && o < un.L) dst.load(xg + (vox0 + o * vox_plane) * a.aux_pitch + cc, nvalid > 8);
      };
      // aux rows (residual / masking tensor) are requested two output planes ahead of their use
      Raw16<T> pre0, pre1;
      arm(0, pre0);
      arm(1, pre1);
      for (int o = 0; o < un.L; ++o) {
        const Raw16<T> cur = pre0;
        pre0 = pre1;
        arm(o + 2, pre1);
        const uint32_t on = oc + static_cast<uint32_t>(o);
        const uint32_t aslot = on & 15u;
        mbar_wait(smem_u32(&sm.acc_full[aslot]), (on >> 4) & 1u);
        tc_fence_after_sync();
        uint32_t r[16];
        __syncwarp();
        tmem_ld16(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + aslot * kStNT + cc, r);
        tmem_ld_wait();
        tc_fence_before_sync();
        mbar_arrive(smem_u32(&sm.acc_empty[aslot]));
        if (!row_ok) continue;
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
        if (has_aux) {
          float xh[16];
          cur.to_float(xh);
          if (!mask_mode) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += xh[j];
#pragma unroll
            for (int j = 0; j < 16; ++j) { s1[j] += v[j]; s2[j] = fmaf(v[j], v[j], s2[j]); }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float h = (xh[j] - sm.mstat[cc + j][0]) * sm.mstat[cc + j][1];
              v[j] = h > 0.f ? v[j] : v[j] * a.slope;
              s1[j] += v[j];
              s2[j] = fmaf(v[j], h, s2[j]);
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) { s1[j] += v[j]; s2[j] = fmaf(v[j], v[j], s2[j]); }
        }
        const size_t vox = vox0 + o * vox_plane;
        float ov[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) ov[j] = v[j];
        Vec8<T>::store(yg + vox * a.y_pitch + cc, ov);
        if (nvalid > 8) {
#pragma unroll
          for (int j = 0; j < 8; ++j) ov[j] = v[8 + j];
          Vec8<T>::store(yg + vox * a.y_pitch + cc + 8, ov);
        }
      }
      oc += static_cast<uint32_t>(un.L);
    }
    if (want_stats && stat_n >= 0) flush_stats();
  }

  // ---------------- teardown ----------------
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace rsb

using namespace rsb;

// Called by rsb_conv3_forward (conv3_fprop.cu) for eligible layers; returns 1 when the layer is not eligible.
int rsb_conv3_stream_try(const RsbConv3Args* p, void* stream) {
  // Default for the eligible layers (RSB_FPROP_STREAM=0 disables it).  Standalone it beats the item-based kernel by 8-13 % on
  // the 32-channel layers (0.212-0.234 ms vs 0.244-0.262 ms at 2 x 128^3); inside the graph-captured train step with the
  // weight gradients on the second stream: 19.97 vs 20.41 ms per step (profiles/r02_*).  (Round 1 measured no gain with
  // eagerly enqueued streams, where the kernel order depended on host timing.)
  const char* on = getenv("RSB_FPROP_STREAM");
  if (on != nullptr && on[0] == '0') return 1;
  if (p->Cin > 32 || p->Cout > 32 || p->Cout <= 16 || p->a_lo != nullptr || p->planes_per_item != 0 || p->pointwise) return 1;
  if (p->dtype != RSB_BF16 && p->dtype != RSB_F32) return 1;
  int sms = p->max_ctas > 0 ? p->max_ctas : rsb_num_sms();
  if (sms <= 0) return 1;
  StreamDev d{};
  d.N = p->N; d.D = p->D; d.H = p->H; d.W = p->W; d.Cin = p->Cin; d.Cout = p->Cout;
  d.w_packed = reinterpret_cast<const uint8_t*>(p->w_packed);
  d.y = p->y; d.y_pitch = p->y_pitch;
  d.mask_mode = p->mask_x != nullptr;
  d.aux = d.mask_mode ? p->mask_x : p->res;
  d.aux_pitch = d.mask_mode ? p->mask_x_pitch : p->res_pitch;
  d.mask_stats = p->mask_stats;
  d.stat_dst = d.mask_mode ? p->bwd_sums : p->out_stats;
  d.stat_pitch = d.mask_mode ? p->mask_x_pitch : p->y_pitch;
  d.eps = p->eps; d.slope = p->slope;
  d.inv_count = 1.0f / (static_cast<float>(p->D) * p->H * p->W);
  d.ksteps = p->Cin <= 16 ? 1 : 2;
  d.tiles_x = (p->W + 7) / 8;
  d.tiles_y = (p->H + 15) / 16;
  d.zchunks = (p->D + kStChunk - 1) / kStChunk;
  const long long units = static_cast<long long>(p->N) * d.tiles_y * d.tiles_x * d.zchunks;
  if (units >= (1LL << 31)) return 1;
  // the column walk pays two extra input planes per unit and needs enough units to balance the SMs
  if (units < 3LL * sms || p->D < 8) return 1;
  d.num_units = static_cast<int>(units);

  CUtensorMap tm;
  int rc = make_act_tensor_map(&tm, p->a, p->a_pitch, p->Cin, p->N, p->D, p->H, p->W, 32, 10, 18, 1);
  if (rc) return rc;
  const size_t smem = kStCtrlBytes + static_cast<size_t>(kStRing) * kStSlotBytes + kStWeightBytes;
  const int grid = static_cast<int>(units < sms ? units : sms);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  if (p->dtype == RSB_BF16) {
    e = cudaFuncSetAttribute(conv3_stream32_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    RSB_REQUIRE(e == cudaSuccess, "conv3 stream: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    conv3_stream32_kernel<__nv_bfloat16><<<grid, kStThreads, smem, st>>>(tm, d);
  } else {
    e = cudaFuncSetAttribute(conv3_stream32_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    RSB_REQUIRE(e == cudaSuccess, "conv3 stream: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    conv3_stream32_kernel<float><<<grid, kStThreads, smem, st>>>(tm, d);
  }
  return check_launch("conv3_stream32_kernel");
}
