// umma_mn_probe.cu — micro-experiment for the wgrad operand layouts: tcgen05.mma with BOTH operands MN-major in
// SWIZZLE_64B / SWIZZLE_128B shared-memory tiles whose rows are voxels (K) and whose row payload is a dense run of
// channels (what a TMA box of an NDHWC tensor lands as):
//   A = dy planes   [plane j][16 y][8 x][OTc ch]  -> M index = (plane j, channel), planes stacked through LBO
//   B = haloed a    [18 yy][10 xx][ITc ch]        -> N index = (kh, channel): the +1-row-of-y shift IS the LBO stride,
//                                                    kw = start-address shift of one row, K-groups (next y) through SBO
// Verifies the descriptor semantics (which of LBO / SBO strides MN-groups vs K-groups in swizzled MN-major mode) and
// that row-shifted start addresses keep the absolute-address swizzle consistent.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/umma_mn_probe tools/umma_mn_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../r-super_b200/csrc/rsb_common.cuh"

namespace rsb {
void set_last_error(const char*, ...) {}
int check_launch(const char*) { return 0; }
}  // namespace rsb
using namespace rsb;

struct Params {
  int otc, itc;    // channels per row of dy / a  (32 -> SWIZZLE_64B rows of 64 B, 64 -> SWIZZLE_128B rows of 128 B)
  int kw, ks;      // in-plane x shift of the a view, K step (16 voxels = 2 y rows)
  int swap;        // 0: LBO = MN-group stride, SBO = K-group stride;  1: swapped
  int nkh;         // kh shifts stacked along N (N = nkh * itc)
  int iters;
};

__device__ __forceinline__ uint64_t mk_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, int mode) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(mode & 7) << 61;
  return d;
}

// byte offset of (row, channel) in a tile of `rb`-byte rows with the TMA swizzle pattern (absolute address bits)
__device__ __host__ inline uint32_t swz_off(uint32_t base_low, uint32_t row, int c, uint32_t rb) {
  const uint32_t raddr = base_low + row * rb;
  uint32_t chunk = c / 8;
  chunk ^= (raddr >> 7) & (rb == 128 ? 7u : 3u);
  return row * rb + chunk * 16 + (c % 8) * 2;
}

__global__ void __launch_bounds__(128, 1) probe_kernel(Params p, const float* __restrict__ dy_src, const float* __restrict__ a_src,
                                                       float* __restrict__ d_out, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem_al = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  uint8_t* dy_buf = smem_al;               // PM planes x 128 rows
  uint8_t* a_buf = smem_al + 64 * 1024;    // 180 rows
  const uint32_t dy_base = smem_u32(dy_buf), a_base = smem_u32(a_buf);
  const uint32_t rba = p.otc * 2, rbb = p.itc * 2;
  const int PM = 128 / p.otc;
  for (int i = threadIdx.x; i < PM * 128 * p.otc; i += blockDim.x) {
    const int c = i % p.otc, row = i / p.otc;  // row = plane*128 + vox
    *reinterpret_cast<__nv_bfloat16*>(dy_buf + swz_off(dy_base, row, c, rba)) = __float2bfloat16(dy_src[i]);
  }
  for (int i = threadIdx.x; i < 180 * p.itc; i += blockDim.x) {
    const int c = i % p.itc, row = i / p.itc;
    *reinterpret_cast<__nv_bfloat16*>(a_buf + swz_off(a_base, row, c, rbb)) = __float2bfloat16(a_src[i]);
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const int N = p.nkh * p.itc;
  const uint32_t idesc = make_idesc_bf16(128, N, 1, 1);
  if (threadIdx.x == 0) {
    const int moda = rba == 128 ? 2 : 4, modb = rbb == 128 ? 2 : 4;
    const uint32_t a_mn = 128 * rba, a_k = 8 * rba;  // plane stride / 8-voxel (one y row) stride
    const uint32_t b_mn = 10 * rbb, b_k = 10 * rbb;  // +1 y row of the halo plane, both as kh stacking and as K group
    const uint64_t ad = mk_desc(dy_base + p.ks * 16 * rba, p.swap ? a_k : a_mn, p.swap ? a_mn : a_k, moda);
    const uint64_t bd = mk_desc(a_base + (2 * p.ks * 10 + p.kw) * rbb, p.swap ? b_k : b_mn, p.swap ? b_mn : b_k, modb);
    uint32_t parity = 0;
    for (int rep = 0; rep < 2; ++rep) {
      const int iters = rep == 0 ? 1 : p.iters;
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) umma_bf16_ss(tmem, ad, bd, idesc, it > 0 ? 1u : 0u);
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), parity);
      parity ^= 1;
      if (rep == 1) cycles[blockIdx.x * 2] = clock64() - t0;
    }
    // streaming pattern of the wgrad kernel: 3 kw x 8 k-steps of distinct descriptors, three accumulators
    {
      const long long t0 = clock64();
      for (int it = 0; it < p.iters / 24; ++it) {
        for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t ad2 = mk_desc(dy_base + ks * 16 * rba, a_mn, a_k, moda);
            const uint64_t bd2 = mk_desc(a_base + (2 * ks * 10 + kw) * rbb, b_mn, b_k, modb);
            umma_bf16_ss(tmem + (N <= 96 ? 128 + kw * 96 : 256), ad2, bd2, idesc, 1u);
          }
        }
      }
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), parity);
      parity ^= 1;
      cycles[blockIdx.x * 2 + 1] = clock64() - t0;
    }
    // issue-gap experiment: G back-to-back MMAs, then the issuing thread is busy for X cycles (what barrier waits /
    // index arithmetic between pipeline stages cost): how much of X does the tensor pipe's queue hide?
    if (blockIdx.x == 0) {
      const int gaps[6] = {0, 50, 100, 200, 400, 800};
      const int groups[3] = {4, 12, 24};
      for (int gi = 0; gi < 3; ++gi)
        for (int xi = 0; xi < 6; ++xi) {
          const int G = groups[gi], X = gaps[xi];
          const long long t0 = clock64();
          for (int it = 0; it < 2400 / G; ++it) {
            for (int q = 0; q < G; ++q) umma_bf16_ss(tmem, ad, bd, idesc, 1u);
            if (X > 0) {
              const long long ts = clock64();
              while (clock64() - ts < X) {}
            }
          }
          umma_commit(smem_u32(&bar));
          mbar_wait(smem_u32(&bar), parity);
          parity ^= 1;
          cycles[512 + gi * 6 + xi] = clock64() - t0;
        }
    }
  }
  __syncthreads();
  tc_fence_after_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int cc = 0; cc < N; cc += 16) {
    uint32_t r[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + cc, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) d_out[(warp * 32 + lane) * N + cc + j] = __uint_as_float(r[j]) / (float)p.iters;  // (main pass only reads columns [0, N))
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  std::vector<float> hdy(4 * 128 * 64), ha(180 * 64);
  float *ddy, *da, *dd;
  long long* dc;
  cudaMalloc(&ddy, hdy.size() * 4); cudaMalloc(&da, ha.size() * 4); cudaMalloc(&dd, 128 * 256 * 4); cudaMalloc(&dc, 1024 * 8);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 130 * 1024);
  struct Case { int otc, itc, kw, ks, swap, nkh; };
  std::vector<Case> cases = {
      {32, 32, 0, 0, 0, 1}, {32, 32, 0, 0, 1, 1}, {32, 32, 1, 3, 0, 3}, {32, 32, 2, 7, 0, 3}, {32, 32, 1, 3, 1, 3},
      {64, 64, 0, 0, 0, 1}, {64, 64, 0, 0, 1, 1}, {64, 64, 1, 3, 0, 3}, {64, 64, 2, 7, 0, 3}, {64, 64, 1, 3, 1, 3},
      {32, 64, 1, 2, 0, 3}, {64, 32, 2, 5, 0, 3}, {64, 64, 1, 4, 0, 4},
  };
  for (const Case& c : cases) {
    Params p{c.otc, c.itc, c.kw, c.ks, c.swap, c.nkh, 264 * 50};
    const int PM = 128 / c.otc, N = c.nkh * c.itc;
    for (size_t i = 0; i < (size_t)PM * 128 * c.otc; ++i) hdy[i] = (float)((int)((i * 7 + i / 13) % 9) - 4);
    for (size_t i = 0; i < (size_t)180 * c.itc; ++i) ha[i] = (float)((int)((i * 5 + i / 11) % 7) - 3);
    cudaMemcpy(ddy, hdy.data(), hdy.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(da, ha.data(), ha.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dd, 0, 128 * 256 * 4);
    probe_kernel<<<1, 128, 130 * 1024>>>(p, ddy, da, dd, dc);
    cudaError_t e = cudaDeviceSynchronize();
    long long hc1[2];
    cudaMemcpy(hc1, dc, 16, cudaMemcpyDeviceToHost);
    probe_kernel<<<148, 128, 130 * 1024>>>(p, ddy, da, dd, dc);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> hd(128 * N);
    long long hc[2 * 148];
    cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hc, dc, 16 * 148, cudaMemcpyDeviceToHost);
    long long mx0 = 0, mx1 = 0;
    for (int i = 0; i < 148; ++i) { mx0 = hc[2 * i] > mx0 ? hc[2 * i] : mx0; mx1 = hc[2 * i + 1] > mx1 ? hc[2 * i + 1] : mx1; }
    int bad = 0;
    for (int m = 0; m < 128; ++m) {
      const int j = m / c.otc, co = m % c.otc;
      for (int n = 0; n < N; ++n) {
        const int kh = n / c.itc, ci = n % c.itc;
        double ref = 0;
        for (int k = 0; k < 16; ++k) {
          const int y = 2 * c.ks + k / 8, x = k % 8;
          const int arow = (y + kh) * 10 + x + c.kw;
          if (arow >= 180) continue;
          ref += (double)hdy[((size_t)j * 128 + y * 8 + x) * c.otc + co] * ha[(size_t)arow * c.itc + ci];
        }
        if (fabs(ref - hd[m * N + n]) > 1e-3) ++bad;
      }
    }
    printf("dy %2dch/row (PM=%d) x a %2dch/row, N=%3d (kh x%d) kw=%d ks=%d %s: %s (bad %5d of %5d) | %.1f cyc/MMA\n", c.otc, PM, c.itc, N, c.nkh,
           c.kw, c.ks, c.swap ? "LBO=K-group SBO=MN-group" : "LBO=MN-group SBO=K-group", bad == 0 ? "CORRECT" : "WRONG  ", bad, 128 * N,
           (double)hc1[0] / 13200);
    {
      long long hg[18];
      cudaMemcpy(hg, dc + 512, 18 * 8, cudaMemcpyDeviceToHost);
      for (int gi = 0; gi < 3; ++gi) {
        printf("      gap test G=%2d:", gi == 0 ? 4 : (gi == 1 ? 12 : 24));
        for (int xi = 0; xi < 6; ++xi) printf("  X=%d: %.1f", xi == 0 ? 0 : (50 << (xi - 1)) , (double)hg[gi * 6 + xi] / 2400);
        printf("  cyc/MMA\n");
      }
    }
    printf("      1 CTA: same-desc %.1f, streaming %.1f cyc/MMA | 148 CTAs: same-desc %.1f, streaming %.1f cyc/MMA\n", (double)hc1[0] / 13200,
           (double)hc1[1] / 13200, (double)mx0 / 13200, (double)mx1 / 13200);
  }
  return 0;
}
