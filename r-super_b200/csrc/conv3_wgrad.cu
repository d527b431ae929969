// conv3_wgrad.cu — weight gradient of the 3x3x3 convolution on tcgen05, TMA-fed.
//
//   dW[co][ci][kd,kh,kw] = sum_{n,z,y,x} dy[n,z,y,x][co] * a[n,z+kd-1,y+kh-1,x+kw-1][ci]
//   (autograd of nn.Conv3d inside ConvNormAct, rsuper_train/model/dim3/conv_layers.py:29-49;
//    a = act(instnorm(x)) is the bf16 operand tensor materialised by rsb_norm_act).
//
// GEMM view: K = voxels, both operands MN-major.  A TMA box of an NDHWC tensor lands in shared memory
// as rows of voxels whose payload is a dense run of 32 / 64 channels (SWIZZLE_64B / SWIZZLE_128B):
// that IS the MN-major UMMA operand layout, so no thread touches operand bytes.
//   A (M side) = dy tiles (16 x 8 voxels x OTc channels).  M = 128 rows = consecutive z-planes (and / or
//                64-channel sub-tiles) stacked through the descriptor's LBO: one MMA covers several kd taps
//                (dy planes z-1..z+2 against a-plane z give kd = 2,1,0 + one ignored block).  Planes live in
//                a ring whose first PM-1 slots are mirrored behind the end, so every window is contiguous.
//   B (N side) = one haloed a-plane (18 x 10 voxels x ITc channels).  N = 3 * ITc: the three kh taps are
//                stacked through LBO = one halo row (10 voxels) — the same stride that advances K by a
//                y-row — and kw is a one-row shift of the start address.
//   accumulators (window, kw) x [128 x 3*ITc] fp32 stay in TMEM for the whole kernel; every persistent
//   CTA sums over its share of the volume and dumps one partial; a second small kernel reduces the
//   partials deterministically into fp32 OIDHW.
// When (Cout, Cin, 27 taps) exceeds 512 TMEM columns the problem is split into groups
// (o-tile, i-tile, kw); CTAs are dealt round-robin to groups.  Tile shapes are chosen per layer by a
// cost model built on the measured MMA cost (profiles/r01_umma_mnmajor_layout_probe.log:
// 57 cycles for N = 96, 97 for N = 192 at M = 128, K = 16).
#include "rsb_common.cuh"
#include "rsb_tma.cuh"

#include <cstdlib>

#include "../../include/rsuper_b200.h"

namespace rsb {

constexpr int kWgThreads = 192;  // warp 0: TMA producer | warp 1: MMA issuer + TMEM owner | warps 2-5: epilogue
constexpr int kWgMaxRing = 8;
constexpr int kWgMaxStages = 4;
constexpr int kWgCtrlBytes = 1024;

struct WgradDev {
  int N, D, H, W, Cin, Cout;
  float* ws;
  // tiling
  int OTc, S, PM, KS, WL;  // dy row channels, sub-tiles per plane, planes per window, windows, planes spanned
  int OT;                  // S * OTc output channels per group
  int ITc, TS;             // a row channels (= input channels per group), kw taps per group
  int Ncols;               // 3 * ITc
  int n_otiles, n_itiles, n_tapsets, n_groups, ranks;
  int tiles_y, tiles_x, zchunks, zlen, n_units;
  int R, SA;               // dy ring entries, a-plane stages
  uint32_t rba, rbb;       // row bytes of dy / a tiles
  uint32_t dy_sub_bytes, dy_slot_bytes, a_stage_bytes, a_box_bytes;
  int piece_floats;        // 3 (kd) * TS * 3 (kh) * OT * ITc
  uint32_t idesc;
  long long* dbg;
  int dbg_mode;  // profiling experiments only: bit 0 = skip the TMA copies after the first ring fill
};

struct __align__(16) WgradSmem {
  uint64_t dy_full[kWgMaxRing], dy_empty[kWgMaxRing];
  uint64_t a_full[kWgMaxStages], a_empty[kWgMaxStages];
  uint64_t done;
  uint32_t tmem_base;
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
  r.y0 = yt * 16;
  r.x0 = xt * 8;
  r.zs = zc * a.zlen;
  r.ze = min(a.D, r.zs + a.zlen);
  return r;
}

// Template parameters = the tile shape (row bytes of the dy / a tiles, windows, kw taps per group): every descriptor
// stride of the MMA issue loop is a compile-time constant, and the profiling hooks (Dbg) are compiled out of the
// production instantiation.  The issuing warp is instruction-bound — the tensor pipe's queue holds ~2 MMAs, so each of
// its ~17 integer / uniform-move instructions per MMA showed up as pipe idle time (profiles/r01h_wgrad_*).
template <int RBA, int RBB, int KS, int TS, bool Dbg>
__global__ void __launch_bounds__(kWgThreads, 1)
conv3_wgrad_kernel(const __grid_constant__ CUtensorMap tm_dy, const __grid_constant__ CUtensorMap tm_a, const WgradDev a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  WgradSmem& sm = *reinterpret_cast<WgradSmem*>(smem_raw);
  const uint32_t dy_base = smem_u32(smem_raw) + kWgCtrlBytes;
  const uint32_t a_base = dy_base + (a.R + a.PM - 1) * a.dy_slot_bytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = blockIdx.x % a.n_groups;
  const int rank = blockIdx.x / a.n_groups;
  const int tapset = group % a.n_tapsets;
  const int itile = (group / a.n_tapsets) % a.n_itiles;
  const int otile = group / (a.n_tapsets * a.n_itiles);
  const int o0 = otile * a.OT, i0 = itile * a.ITc;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kWgMaxRing; ++i) {
      mbar_init(smem_u32(&sm.dy_full[i]), 1);
      mbar_init(smem_u32(&sm.dy_empty[i]), 1);
    }
    for (int i = 0; i < kWgMaxStages; ++i) {
      mbar_init(smem_u32(&sm.a_full[i]), 1);
      mbar_init(smem_u32(&sm.a_empty[i]), 1);
    }
    mbar_init(smem_u32(&sm.done), 1);
    mbar_fence_init();
    tma_prefetch_desc(&tm_dy);
    tma_prefetch_desc(&tm_a);
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
    // =========================== TMA producer ===========================
    // The whole warp walks the schedule (all state is warp-uniform, so it lives in uniform registers and no
    // per-instruction election loop is needed); one elected lane issues.  Ring positions are carried
    // incrementally: an integer division per plane in this thread was a measurable stall.
    if (!(a.dbg_mode & 2)) {
      uint32_t dslot = 0, dphase = 0;  // dy ring write position
      uint32_t ast = 0, aphase = 0;    // a-plane stage write position
      bool dy_live = true, a_live = true;
      auto load_dy = [&](const WgUnit& un, int z) {
        mbar_wait(smem_u32(&sm.dy_empty[dslot]), dphase ^ 1u);
        const bool mirror = static_cast<int>(dslot) < a.PM - 1;
        const uint32_t bar = smem_u32(&sm.dy_full[dslot]);
        if (elect_one()) {
          if (!dy_live) {
            mbar_arrive(bar);
          } else {
            mbar_arrive_expect_tx(bar, a.dy_slot_bytes * (mirror ? 2u : 1u));
            for (int s = 0; s < a.S; ++s) {
              const uint32_t dst = dy_base + dslot * a.dy_slot_bytes + s * a.dy_sub_bytes;
              tma_load_5d(dst, &tm_dy, o0 + s * a.OTc, un.x0, un.y0, z, un.n, bar);
              if (mirror) tma_load_5d(dst + a.R * a.dy_slot_bytes, &tm_dy, o0 + s * a.OTc, un.x0, un.y0, z, un.n, bar);
            }
          }
        }
        __syncwarp();
        if (++dslot == static_cast<uint32_t>(a.R)) {
          dslot = 0;
          dphase ^= 1u;
          if (a.dbg_mode & 1) dy_live = false;
        }
      };
      for (int u = rank; u < a.n_units; u += a.ranks) {
        const WgUnit un = wg_decode_unit(a, u);
        for (int i = 0; i < a.WL - 1; ++i) load_dy(un, un.zs - 1 + i);
        for (int zb = un.zs; zb < un.ze; ++zb) {
          load_dy(un, zb + a.WL - 2);
          mbar_wait(smem_u32(&sm.a_empty[ast]), aphase ^ 1u);
          const uint32_t bar = smem_u32(&sm.a_full[ast]);
          if (elect_one()) {
            if (!a_live) {
              mbar_arrive(bar);
            } else {
              mbar_arrive_expect_tx(bar, a.a_box_bytes);
              tma_load_5d(a_base + ast * a.a_stage_bytes, &tm_a, i0, un.x0 - 1, un.y0 - 1, zb, un.n, bar);
            }
          }
          __syncwarp();
          if (++ast == static_cast<uint32_t>(a.SA)) {
            ast = 0;
            aphase ^= 1u;
            if (a.dbg_mode & 1) a_live = false;
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    // warp-converged waits; one elected lane issues (uniform-datapath UTCHMMA, no per-instruction branch).
    // The tensor pipe's instruction queue is shallow: whatever this thread does between two MMAs is NOT hidden
    // behind the MMAs in flight, so ring positions are carried incrementally (no division, no multiplication, no
    // clock reads unless Dbg) and every descriptor is base + compile-time constant.
    constexpr uint32_t PM = KS == 1 ? 4u : (KS == 2 ? 2u : 1u);
    constexpr uint32_t WL = PM * KS;
    constexpr uint32_t S = 128u / (PM * (RBA / 2));            // 64-channel sub-tiles per plane
    constexpr uint32_t lt_a = RBA == 128 ? kLayoutSw128 : kLayoutSw64;
    constexpr uint32_t lt_b = RBB == 128 ? kLayoutSw128 : kLayoutSw64;
    constexpr uint32_t dy_sub_bytes = 128u * RBA;
    constexpr uint32_t dy_slot16 = (S * dy_sub_bytes) >> 4;
    constexpr uint32_t a_stage16 = ((180u * RBB + 1023u) / 1024u * 1024u) >> 4;
    constexpr uint32_t a_kstep = (16u * RBA) >> 4;             // 16 voxels = 2 y rows of the dy tile
    constexpr uint32_t b_kstep = (20u * RBB) >> 4;             // ... = 2 halo rows
    constexpr uint32_t b_kw16 = RBB >> 4;
    constexpr uint32_t Ncols = 3u * (RBB / 2);
    const uint32_t a_hi = desc_hi(8 * RBA, lt_a);    // dy: K groups = 8 voxels (one y row of the tile)
    const uint32_t b_hi = desc_hi(10 * RBB, lt_b);   // a:  K groups = next y row of the haloed plane
    constexpr uint32_t a_lbo = ((dy_sub_bytes >> 4) & 0x3FFFu) << 16;  // M groups: next sub-tile / next plane
    constexpr uint32_t b_lbo = (((10u * RBB) >> 4) & 0x3FFFu) << 16;   // N groups: kh -> next halo row
    const uint32_t a_ring_lo = a_lbo | (dy_base >> 4);
    const uint32_t b_ring_lo = b_lbo | ((a_base >> 4) + ((static_cast<uint32_t>(tapset * TS) * RBB) >> 4));
    const uint32_t R = a.R, SA = a.SA;
    const uint32_t idesc = a.idesc;
    const uint32_t bar_dy_full = smem_u32(&sm.dy_full[0]), bar_dy_empty = smem_u32(&sm.dy_empty[0]);
    const uint32_t bar_a_full = smem_u32(&sm.a_full[0]), bar_a_empty = smem_u32(&sm.a_empty[0]);
    const bool no_wait = Dbg && (a.dbg_mode & 2) != 0, no_commit = Dbg && (a.dbg_mode & 4) != 0;
    uint32_t head = 0;                  // ring position of the oldest dy plane of the current window (offset 0)
    uint32_t nw = 0, nw_phase = 0;      // ring position / phase of the plane the next wait is for
    uint32_t st = 0, st_phase = 0;      // a-plane stage
    uint32_t first = 0;                 // 0 until the accumulators have been written once
    uint32_t t = 0;
    long long tw_a = 0, tw_dy = 0;
    const long long t_begin = Dbg ? clock64() : 0;
    unsigned long long ns_begin = 0;
    if (Dbg) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns_begin));
    for (int u = rank; u < a.n_units; u += a.ranks) {
      const WgUnit un = wg_decode_unit(a, u);
      const int nsteps = un.ze - un.zs;
      // column start: the first WL - 1 planes of the window (afterwards every step waits for ONE new plane)
      if (!no_wait) {
        for (uint32_t i = 0; i + 1 < WL; ++i) {
          mbar_wait(bar_dy_full + 8u * nw, nw_phase);
          if (++nw == R) { nw = 0; nw_phase ^= 1u; }
        }
      }
      for (int j = 0; j < nsteps; ++j) {
        long long tq = Dbg ? clock64() : 0;
        if (!no_wait) mbar_wait(bar_dy_full + 8u * nw, nw_phase);
        if (Dbg) { const long long now = clock64(); tw_dy += now - tq; tq = now; }
        if (!no_wait) mbar_wait(bar_a_full + 8u * st, st_phase);
        if (Dbg) tw_a += clock64() - tq;
        tc_fence_after_sync();
        if (elect_one()) {
          const uint32_t b_stage_lo = b_ring_lo + st * a_stage16;
          uint32_t wslot = head;
          uint32_t d_col = tmem_base;
#pragma unroll
          for (int w = 0; w < KS; ++w) {
            const uint32_t a_win_lo = a_ring_lo + wslot * dy_slot16;
#pragma unroll
            for (int ti = 0; ti < TS; ++ti) {
#pragma unroll
              for (int ks = 0; ks < 8; ++ks) {
                umma_bf16_ss(d_col, desc_join(a_hi, a_win_lo + ks * a_kstep),
                             desc_join(b_hi, b_stage_lo + ti * b_kw16 + ks * b_kstep), idesc, ks == 0 ? first : 1u);
              }
              d_col += Ncols;
            }
            wslot += PM;
            if (wslot >= R) wslot -= R;
          }
          if (!no_commit) {
            umma_commit(bar_a_empty + 8u * st);
            umma_commit(bar_dy_empty + 8u * head);
            if (j == nsteps - 1) {
              uint32_t s2 = head;
              for (uint32_t i = 1; i < WL; ++i) {
                if (++s2 == R) s2 = 0;
                umma_commit(bar_dy_empty + 8u * s2);
              }
            }
          }
        }
        __syncwarp();
        first = 1u;
        ++t;
        if (++st == SA) { st = 0; st_phase ^= 1u; }
        if (++head == R) head = 0;
        if (++nw == R) { nw = 0; nw_phase ^= 1u; }
      }
      // the WL-1 trailing planes of this column are consumed too (nw already points past them)
      head += WL - 1;
      if (head >= R) head -= R;
    }
    if (elect_one()) umma_commit(smem_u32(&sm.done));
    __syncwarp();
    // Only this warp polls the final mbarrier; the four epilogue warps sleep in a hardware named barrier meanwhile.
    mbar_wait(smem_u32(&sm.done), 0);
    named_bar_sync(1, 160);
    if (Dbg && a.dbg != nullptr && lane == 0) {
      long long* d = a.dbg + static_cast<size_t>(blockIdx.x) * 16;
      unsigned long long ns_end;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns_end));
      d[0] = clock64() - t_begin; d[1] = tw_a; d[2] = t; d[3] = tw_dy; d[4] = static_cast<long long>(ns_end - ns_begin);
    }
  } else {
    // =========================== epilogue (warps 2..5) ===========================
    named_bar_sync(1, 160);
    tc_fence_after_sync();
    __syncwarp();
    const int ew = warp & 3;  // TMEM lane quarter this warp may access
    const int row = ew * 32 + lane;
    const int g = row / a.OTc, co_in = row % a.OTc;
    const int j = g / a.S, co = (g % a.S) * a.OTc + co_in;
    float* piece = a.ws + static_cast<size_t>(blockIdx.x) * a.piece_floats;
    for (int w = 0; w < a.KS; ++w) {
      const int off = w * a.PM + j;  // dy plane offset relative to zb-1  ->  kd = 2 - off
      for (int ti = 0; ti < a.TS; ++ti) {
        for (int cc = 0; cc < a.Ncols; cc += 16) {
          uint32_t r[16];
          tmem_ld16(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + (w * a.TS + ti) * a.Ncols + cc, r);
          tmem_ld_wait();
          if (off <= 2) {
            const int kh = cc / a.ITc, ci = cc % a.ITc;
            float* dst = piece + (((static_cast<size_t>(off) * a.TS + ti) * 3 + kh) * a.OT + co) * a.ITc + ci;
#pragma unroll
            for (int q = 0; q < 16; q += 4)
              *reinterpret_cast<float4*>(dst + q) = make_float4(__uint_as_float(r[q]), __uint_as_float(r[q + 1]),
                                                                __uint_as_float(r[q + 2]), __uint_as_float(r[q + 3]));
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

// Deterministic reduction of the per-CTA partials into fp32 OIDHW (fixed summation order => bit-reproducible).
// One thread per element walks the ranks with four independent accumulators so that several loads are in flight.
__global__ void wgrad_reduce_kernel(const WgradDev a, float* __restrict__ dw, int accumulate) {
  const long long total = static_cast<long long>(a.n_groups) * a.piece_floats;
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int group = static_cast<int>(e / a.piece_floats);
    const size_t pe = static_cast<size_t>(e % a.piece_floats);
    int r = static_cast<int>(pe);
    const int ci = r % a.ITc; r /= a.ITc;
    const int co = r % a.OT; r /= a.OT;
    const int kh = r % 3; r /= 3;
    const int ti = r % a.TS; r /= a.TS;
    const int off = r;  // 0..2
    const int tapset = group % a.n_tapsets;
    const int itile = (group / a.n_tapsets) % a.n_itiles;
    const int otile = group / (a.n_tapsets * a.n_itiles);
    const int o = otile * a.OT + co, i = itile * a.ITc + ci;
    if (o >= a.Cout || i >= a.Cin) continue;
    const int tap = (2 - off) * 9 + kh * 3 + tapset * a.TS + ti;
    const float* src = a.ws + static_cast<size_t>(group) * a.piece_floats + pe;
    const size_t rstride = static_cast<size_t>(a.n_groups) * a.piece_floats;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int rk = 0;
    for (; rk + 4 <= a.ranks; rk += 4) {
      a0 += src[static_cast<size_t>(rk) * rstride];
      a1 += src[static_cast<size_t>(rk + 1) * rstride];
      a2 += src[static_cast<size_t>(rk + 2) * rstride];
      a3 += src[static_cast<size_t>(rk + 3) * rstride];
    }
    for (; rk < a.ranks; ++rk) a0 += src[static_cast<size_t>(rk) * rstride];
    const float acc = (a0 + a1) + (a2 + a3);
    float* dst = dw + (static_cast<size_t>(o) * a.Cin + i) * 27 + tap;
    *dst = accumulate ? (*dst + acc) : acc;
  }
}

template <int RBA, int RBB, int KS, int TS, bool Dbg>
static int launch_wgrad(const CUtensorMap& tm_dy, const CUtensorMap& tm_a, const WgradDev& d, int grid, size_t smem, cudaStream_t st) {
  auto kern = conv3_wgrad_kernel<RBA, RBB, KS, TS, Dbg>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) {
    set_last_error("wgrad: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    return -2;
  }
  kern<<<grid, kWgThreads, smem, st>>>(tm_dy, tm_a, d);
  return check_launch("conv3_wgrad_kernel");
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// Tiling plan shared by the workspace query and the launcher.  Candidates (OTc, S, ITc, TS):
//   A (32,1,32,3): M = 4 planes x 32 co, N = 96, all 9 in-plane taps per CTA     (C = 32 layers)
//   B (32,1,64,1): N = 192                                                         (Cout <= 32, wide Cin)
//   C (64,1,64,1): M = 2 planes x 64 co (two windows), N = 192                     (64-channel layers)
//   D (64,2,32,1): M = 128 co of one plane (three windows), N = 96                 (Cout >= 128)
//   E (64,1,32,1): M = 2 planes x 64 co, N = 96
static int wgrad_plan(int Cout, int Cin, int N, int D, int H, int W, int max_ctas, WgradDev& best) {
  struct Cand { int OTc, S, ITc, TS; };
  const Cand cands[5] = {{32, 1, 32, 3}, {32, 1, 64, 1}, {64, 1, 64, 1}, {64, 2, 32, 1}, {64, 1, 32, 1}};
  double best_cost = -1;
  const char* force = getenv("RSB_WGRAD_CAND");   // experiments: restrict the search to one candidate (0..4)
  for (int ci = 0; ci < 5; ++ci) {
    const Cand& c = cands[ci];
    if (force != nullptr && force[0] >= '0' && force[0] <= '4' && ci != force[0] - '0') continue;
    WgradDev d = best;
    d.OTc = c.OTc; d.S = c.S; d.ITc = c.ITc; d.TS = c.TS;
    d.OT = d.S * d.OTc;
    d.PM = 128 / d.OT;
    d.KS = cdiv(3, d.PM);
    d.WL = d.PM * d.KS;
    d.Ncols = 3 * d.ITc;
    if (d.KS * d.TS * d.Ncols > 512) continue;
    d.rba = d.OTc * 2; d.rbb = d.ITc * 2;
    d.dy_sub_bytes = 128 * d.rba;
    d.dy_slot_bytes = d.S * d.dy_sub_bytes;
    d.a_box_bytes = 180 * d.rbb;
    d.a_stage_bytes = (d.a_box_bytes + 1023) / 1024 * 1024;
    // ring / stages: as deep as 227 KB allows, preferring >= 3 a-plane stages
    d.R = 0;
    for (int min_sa = 3; min_sa >= 2 && d.R == 0; --min_sa)
      for (int R = kWgMaxRing; R >= d.WL + 1 && d.R == 0; --R)
        for (int SA = kWgMaxStages; SA >= min_sa; --SA) {
          const size_t need = kWgCtrlBytes + static_cast<size_t>(R + d.PM - 1) * d.dy_slot_bytes + static_cast<size_t>(SA) * d.a_stage_bytes;
          if (need <= 226 * 1024) { d.R = R; d.SA = SA; break; }
        }
    if (d.R == 0) continue;
    d.n_otiles = cdiv(Cout, d.OT);
    d.n_itiles = cdiv(Cin, d.ITc);
    d.n_tapsets = 3 / d.TS;
    d.n_groups = d.n_otiles * d.n_itiles * d.n_tapsets;
    d.piece_floats = 3 * d.TS * 3 * d.OT * d.ITc;
    d.tiles_y = cdiv(H, 16);
    d.tiles_x = cdiv(W, 8);
    int ranks = max_ctas / d.n_groups;
    if (ranks < 1) ranks = 1;
    // z chunking: enough units for every rank of a group, but keep chunks long (each pays WL-1 extra dy planes)
    const long long cols = static_cast<long long>(N) * d.tiles_y * d.tiles_x;
    int zlen = D;
    while (zlen > 8 && cols * cdiv(D, zlen) < 4LL * ranks) zlen = (zlen + 1) / 2;
    d.zlen = zlen;
    d.zchunks = cdiv(D, zlen);
    const long long units = cols * d.zchunks;
    if (units >= (1LL << 31)) continue;
    d.n_units = static_cast<int>(units);
    if (ranks > d.n_units) ranks = d.n_units;
    d.ranks = ranks;
    d.idesc = make_idesc_bf16(128, d.Ncols, 1, 1);
    // cost model (cycles): MMA time vs TMA time per z-step, times steps per CTA, times waves
    const double t_mma = d.KS * d.TS * 8.0 * (d.Ncols <= 96 ? 58.0 : 97.0);
    const double t_tma = 180.0 * (d.rbb == 128 ? 2.0 : 1.5) +
                         d.S * 128.0 * (d.rba == 128 ? 2.0 : 1.5) * (1.0 + static_cast<double>(d.PM - 1) / d.R);
    const double step = t_mma > t_tma ? t_mma : t_tma;
    const double steps_per_cta = static_cast<double>(cdiv(d.n_units, d.ranks)) * (d.zlen + 0.25 * (d.WL - 1));
    const double waves = cdiv(d.n_groups * d.ranks, max_ctas);
    const double cost = steps_per_cta * step * waves + 3000.0 * d.KS * d.TS;  // + epilogue dump
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = d; }
  }
  return best_cost < 0 ? -1 : 0;
}

}  // namespace rsb

using namespace rsb;

static long long* g_wgrad_timing_buffer = nullptr;
// profiling aid (see rsb_debug_set_timing_buffer): [0] MMA warp total cycles, [1] its wait on a-planes,
// [2] steps, [3] its wait on dy planes
extern "C" int rsb_debug_set_wgrad_timing_buffer(void* device_ptr) {
  g_wgrad_timing_buffer = reinterpret_cast<long long*>(device_ptr);
  return 0;
}

extern "C" size_t rsb_conv3_wgrad_workspace_bytes(int Cout, int Cin, int max_ctas) {
  if (max_ctas <= 0) max_ctas = rsb_num_sms();
  if (max_ctas <= 0) max_ctas = 148;
  // geometry only influences WHICH candidate tiling wins; size for the largest (CTAs x piece) over all of them
  size_t worst = 0;
  const int cand[5][4] = {{32, 1, 32, 3}, {32, 1, 64, 1}, {64, 1, 64, 1}, {64, 2, 32, 1}, {64, 1, 32, 1}};
  for (const auto& c : cand) {
    const int OT = c[0] * c[1], ITc = c[2], TS = c[3];
    const size_t groups = static_cast<size_t>((Cout + OT - 1) / OT) * ((Cin + ITc - 1) / ITc) * (3 / TS);
    const size_t ctas = groups > static_cast<size_t>(max_ctas) ? groups : static_cast<size_t>(max_ctas);
    const size_t b = ctas * (static_cast<size_t>(27) * TS * OT * ITc) * sizeof(float);
    if (b > worst) worst = b;
  }
  return worst;
}

extern "C" int rsb_conv3_wgrad(const RsbConv3WgradArgs* p, void* stream) {
  RSB_REQUIRE(p != nullptr, "wgrad: null args");
  RSB_REQUIRE(p->a && p->dy && p->dw_oidhw && p->workspace, "wgrad: null pointer");
  RSB_REQUIRE(p->N > 0 && p->D > 0 && p->H > 0 && p->W > 0, "wgrad: bad geometry");
  RSB_REQUIRE(p->Cin > 0 && p->Cin % 8 == 0 && p->Cout > 0 && p->Cout % 8 == 0,
              "wgrad: channel counts must be positive multiples of 8 (Cin=%d Cout=%d)", p->Cin, p->Cout);
  RSB_REQUIRE(p->a_pitch >= p->Cin && p->a_pitch % 8 == 0 && p->dy_pitch >= p->Cout && p->dy_pitch % 8 == 0,
              "wgrad: bad pitch");
  int sms = p->max_ctas > 0 ? p->max_ctas : rsb_num_sms();
  RSB_REQUIRE(sms > 0, "wgrad: could not query the SM count");

  WgradDev d{};
  d.N = p->N; d.D = p->D; d.H = p->H; d.W = p->W; d.Cin = p->Cin; d.Cout = p->Cout;
  d.ws = reinterpret_cast<float*>(p->workspace);
  d.dbg = g_wgrad_timing_buffer;
  if (d.dbg != nullptr) {
    const char* m = getenv("RSB_WGRAD_DEBUG_MODE");
    d.dbg_mode = m ? atoi(m) : 0;
  }
  RSB_REQUIRE(wgrad_plan(p->Cout, p->Cin, p->N, p->D, p->H, p->W, sms, d) == 0, "wgrad: no feasible tiling");
  const int grid = d.ranks * d.n_groups;
  const size_t need = static_cast<size_t>(grid) * d.piece_floats * sizeof(float);
  RSB_REQUIRE(p->workspace_bytes >= need, "wgrad: workspace too small (%zu < %zu)", p->workspace_bytes, need);

  CUtensorMap tm_dy, tm_a;
  int rc = make_act_tensor_map(&tm_dy, p->dy, p->dy_pitch, p->Cout, p->N, p->D, p->H, p->W, d.OTc, 8, 16, 1);
  if (rc) return rc;
  rc = make_act_tensor_map(&tm_a, p->a, p->a_pitch, p->Cin, p->N, p->D, p->H, p->W, d.ITc, 10, 18, 1);
  if (rc) return rc;

  size_t smem = kWgCtrlBytes + static_cast<size_t>(d.R + d.PM - 1) * d.dy_slot_bytes + static_cast<size_t>(d.SA) * d.a_stage_bytes;
  if (smem < 120 * 1024) smem = 120 * 1024;  // 1 CTA / SM: each CTA owns all 512 TMEM columns
  RSB_REQUIRE(smem <= 227 * 1024, "wgrad: shared memory budget exceeded (%zu)", smem);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool dbg = d.dbg != nullptr;
  rc = -1;
#define RSB_WG_CASE(RBA_, RBB_, KS_, TS_)                                                                         \
  if (rc == -1 && d.rba == RBA_ && d.rbb == RBB_ && d.KS == KS_ && d.TS == TS_)                                    \
    rc = dbg ? launch_wgrad<RBA_, RBB_, KS_, TS_, true>(tm_dy, tm_a, d, grid, smem, st)                            \
             : launch_wgrad<RBA_, RBB_, KS_, TS_, false>(tm_dy, tm_a, d, grid, smem, st);
  RSB_WG_CASE(64, 64, 1, 3)     // A
  RSB_WG_CASE(64, 128, 1, 1)    // B
  RSB_WG_CASE(128, 128, 2, 1)   // C
  RSB_WG_CASE(128, 64, 3, 1)    // D
  RSB_WG_CASE(128, 64, 2, 1)    // E
#undef RSB_WG_CASE
  RSB_REQUIRE(rc != -1, "wgrad: no kernel instantiation for tiling (rba %u rbb %u KS %d TS %d)", d.rba, d.rbb, d.KS, d.TS);
  if (rc) return rc;
  const long long total = static_cast<long long>(d.n_groups) * d.piece_floats;
  int rblocks = static_cast<int>((total + 127) / 128);
  if (rblocks > sms * 16) rblocks = sms * 16;
  wgrad_reduce_kernel<<<rblocks, 128, 0, st>>>(d, p->dw_oidhw, p->accumulate);
  return check_launch("wgrad_reduce_kernel");
}
