// umma_layout_probe.cu — micro-experiment: tcgen05.mma cost and correctness for different shared-memory
// operand layouts, in particular "tap views" (row-shifted start addresses, non-canonical group stride)
// inside a SWIZZLE_128B / SWIZZLE_64B K-major halo tile.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_layout_probe tools/umma_layout_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../r-super_b200/csrc/rsb_common.cuh"

namespace rsb {
void set_last_error(const char*, ...) {}
int check_launch(const char*) { return 0; }
}  // namespace rsb
using namespace rsb;

struct Params {
  int mode;        // 0 = no swizzle, 2 = 128B swizzle, 4 = 64B swizzle
  int start_row;   // first row of the view
  int group_rows;  // rows between consecutive 8-row groups (8 = dense/canonical, 10 = halo tile)
  int use_base_offset;
  int N;           // MMA N
  int ksteps;      // K = 16 * ksteps per row
  int iters;       // timing repetitions
  int a_rows;
  int nacc;        // number of rotating accumulators (independent MMAs between dependent ones)
};

__device__ __forceinline__ uint64_t mk_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, int mode, int base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)(mode & 7) << 61;
  return d;
}

// element (row r, k) byte offset inside the A region for each layout
__device__ __host__ inline uint32_t a_offset(int mode, int r, int k, int row_elems, uint32_t base_addr_low) {
  if (mode == 0) {
    // [k/8][row][8]: 16 B per (row, k-group); k-group planes of a_rows*16 bytes are handled by caller
    return 0;
  }
  const uint32_t row_bytes = row_elems * 2;  // 128 (mode 2) or 64 (mode 4)
  const uint32_t raddr = base_addr_low + r * row_bytes;
  uint32_t chunk = (k / 8);
  if (mode == 2) chunk ^= (raddr >> 7) & 7;
  if (mode == 4) chunk ^= (raddr >> 7) & 3;
  return r * row_bytes + chunk * 16 + (k % 8) * 2;
}

__global__ void __launch_bounds__(128, 1) probe_kernel(Params p, const float* __restrict__ a_src, const float* __restrict__ b_src,
                                                       float* __restrict__ d_out, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int K = 16 * p.ksteps;
  uint8_t* smem_al = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);   // force 1024 B alignment
  uint8_t* a_buf = smem_al;                      // up to 64 KB
  uint8_t* b_buf = smem_al + 64 * 1024;          // 1024-aligned
  const uint32_t a_base = smem_u32(a_buf), b_base = smem_u32(b_buf);
  const int row_elems = (p.mode == 2) ? 64 : (p.mode == 4 ? 32 : K);
  // ---- fill A ----
  for (int i = threadIdx.x; i < p.a_rows * K; i += blockDim.x) {
    const int r = i / K, k = i % K;
    const __nv_bfloat16 v = __float2bfloat16(a_src[i]);
    uint32_t off;
    if (p.mode == 0) off = ((k / 8) * p.a_rows + r) * 16 + (k % 8) * 2;
    else off = a_offset(p.mode, r, k, row_elems, a_base);
    *reinterpret_cast<__nv_bfloat16*>(a_buf + off) = v;
  }
  // ---- fill B (canonical, dense groups) ----
  for (int i = threadIdx.x; i < p.N * K; i += blockDim.x) {
    const int n = i / K, k = i % K;
    const __nv_bfloat16 v = __float2bfloat16(b_src[i]);
    uint32_t off;
    if (p.mode == 0) off = ((n / 8) * (K / 8) + (k / 8)) * 128 + (n % 8) * 16 + (k % 8) * 2;
    else off = a_offset(p.mode, n, k, row_elems, b_base);
    *reinterpret_cast<__nv_bfloat16*>(b_buf + off) = v;
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = make_idesc_bf16(128, p.N, 0, 0);
  const uint32_t row_bytes = row_elems * 2;
  if (threadIdx.x == 0) {
    uint32_t parity = 0;
    for (int rep = 0; rep < 2; ++rep) {
      const int iters = rep == 0 ? 1 : p.iters;
      // descriptors are precomputed so the timed loop is ONLY tcgen05.mma issue (+ one AND/ADD)
      uint64_t adv[4], bdv[4];
      for (int s = 0; s < p.ksteps; ++s) {
        if (p.mode == 0) {
          const uint32_t aaddr = a_base + (2 * s * p.a_rows + p.start_row) * 16;
          adv[s] = mk_desc(aaddr, p.a_rows * 16, p.group_rows * 16, 0, 0);
          bdv[s] = mk_desc(b_base + s * 256, 128, (K / 8) * 128, 0, 0);
        } else {
          const uint32_t aaddr = a_base + p.start_row * row_bytes + s * 32;
          const int bo = p.use_base_offset ? ((aaddr >> 7) & 7) : 0;
          adv[s] = mk_desc(aaddr, 16, p.group_rows * row_bytes, p.mode, bo);
          bdv[s] = mk_desc(b_base + s * 32, 16, 8 * row_bytes, p.mode, 0);
        }
      }
      const uint32_t accmask = p.nacc - 1;  // nacc is a power of two
      const uint32_t nN = p.N;
      const long long t0 = clock64();
      if (p.ksteps == 2) {
#pragma unroll 4
        for (int it = 0; it < iters; ++it) {
          const uint32_t d = tmem + (it & accmask) * nN;
          const uint32_t acc = (it > (int)accmask) ? 1u : 0u;
          umma_bf16_ss(d, adv[0], bdv[0], idesc, acc);
          umma_bf16_ss(d, adv[1], bdv[1], idesc, 1u);
        }
      } else {
#pragma unroll 4
        for (int it = 0; it < iters; ++it) {
          const uint32_t d = tmem + (it & accmask) * nN;
          const uint32_t acc = (it > (int)accmask) ? 1u : 0u;
          umma_bf16_ss(d, adv[0], bdv[0], idesc, acc);
          umma_bf16_ss(d, adv[1], bdv[1], idesc, 1u);
          umma_bf16_ss(d, adv[2], bdv[2], idesc, 1u);
          umma_bf16_ss(d, adv[3], bdv[3], idesc, 1u);
        }
      }
      umma_commit(smem_u32(&bar));
      const long long t1 = clock64();
      mbar_wait(smem_u32(&bar), parity);
      parity ^= 1;
      const long long t2 = clock64();
      if (rep == 1) { cycles[0] = t1 - t0; cycles[1] = t2 - t0; }
      if (rep == 0) {
        // result of the single pass is read out below; re-zero by the next pass's first MMA
      }
      if (rep == 0) cycles[2] = t2 - t0;
    }
  }
  __syncthreads();
  tc_fence_after_sync();
  // read D (after iters accumulations of the same product: value = iters * product; report / iters)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int cc = 0; cc < p.N; cc += 16) {
    uint32_t r[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + cc, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) d_out[(warp * 32 + lane) * p.N + cc + j] = __uint_as_float(r[j]) / (float)(p.iters / p.nacc);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  const int a_rows = 220;
  std::vector<float> ha(a_rows * 64), hb(256 * 64);
  for (int r = 0; r < a_rows; ++r)
    for (int k = 0; k < 64; ++k) ha[r * 64 + k] = (float)(((r * 7 + k * 3) % 13) - 6);
  for (int n = 0; n < 256; ++n)
    for (int k = 0; k < 64; ++k) hb[n * 64 + k] = (float)(((n * 5 + k) % 7) - 3);
  float *da, *db, *dd;
  long long* dc;
  cudaMalloc(&da, ha.size() * 4); cudaMalloc(&db, hb.size() * 4); cudaMalloc(&dd, 128 * 256 * 4); cudaMalloc(&dc, 64);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 130 * 1024);
  struct Case { int mode, start, group, bo, N, ksteps; const char* name; int nacc = 1; };
  std::vector<Case> cases = {
      {0, 0, 8, 0, 32, 2, "noswz dense (SBO=128B)          "},
      {0, 0, 10, 0, 32, 2, "noswz halo  (SBO=160B) start 0  "},
      {0, 11, 10, 0, 32, 2, "noswz halo  (SBO=160B) start 11 "},
      {0, 11, 10, 0, 128, 2, "noswz halo  N=128     start 11 "},
      {2, 0, 8, 0, 32, 4, "sw128 canonical (SBO=1024)      "},
      {2, 0, 10, 0, 32, 4, "sw128 halo SBO=1280 start 0     "},
      {2, 1, 10, 0, 32, 4, "sw128 halo start 1  bo=0        "},
      {2, 1, 10, 1, 32, 4, "sw128 halo start 1  bo=addr     "},
      {2, 11, 10, 0, 32, 4, "sw128 halo start 11 bo=0        "},
      {2, 11, 10, 1, 32, 4, "sw128 halo start 11 bo=addr     "},
      {2, 22, 10, 0, 32, 4, "sw128 halo start 22 bo=0        "},
      {2, 11, 10, 0, 64, 4, "sw128 halo start 11 N=64        "},
      {2, 11, 10, 0, 128, 4, "sw128 halo start 11 N=128       "},
      {2, 11, 10, 0, 256, 4, "sw128 halo start 11 N=256       "},
      {4, 0, 8, 0, 32, 2, "sw64 canonical (SBO=512)        "},
      {4, 11, 10, 0, 32, 2, "sw64 halo SBO=640 start 11 bo=0 "},
      {4, 11, 10, 1, 32, 2, "sw64 halo SBO=640 start 11 bo=ad"},
      {4, 3, 10, 0, 64, 2, "sw64 halo start 3  N=64         "},
      {2, 11, 10, 0, 32, 4, "sw128 halo N=32  2 accumulators ", 2},
      {2, 11, 10, 0, 32, 4, "sw128 halo N=32  4 accumulators ", 4},
      {2, 11, 10, 0, 32, 4, "sw128 halo N=32  8 accumulators ", 8},
      {2, 11, 10, 0, 64, 4, "sw128 halo N=64  4 accumulators ", 4},
      {2, 11, 10, 0, 128, 4, "sw128 halo N=128 4 accumulators ", 4},
      {0, 11, 10, 0, 32, 2, "noswz halo N=32  4 accumulators ", 4},
      {0, 11, 10, 0, 32, 2, "noswz halo N=32  8 accumulators ", 8},
      {0, 11, 10, 0, 128, 2, "noswz halo N=128 4 accumulators ", 4},
      {4, 11, 10, 0, 32, 2, "sw64  halo N=32  8 accumulators ", 8},
      {0, 11, 10, 0, 64, 2, "noswz halo N=64                 "},
      {0, 11, 10, 0, 96, 2, "noswz halo N=96                 "},
      {0, 11, 10, 0, 192, 2, "noswz halo N=192                "},
      {0, 11, 10, 0, 256, 2, "noswz halo N=256                "},
      {0, 11, 10, 0, 96, 2, "noswz halo N=96  4 accumulators ", 4},
  };
  for (const Case& c : cases) {
    Params p{c.mode, c.start, c.group, c.bo, c.N, c.ksteps, 512, a_rows, c.nacc};
    const int K = 16 * p.ksteps;
    // repack sources to K columns
    std::vector<float> pa(a_rows * K), pb(256 * K);
    for (int r = 0; r < a_rows; ++r) for (int k = 0; k < K; ++k) pa[r * K + k] = ha[r * 64 + k];
    for (int n = 0; n < 256; ++n) for (int k = 0; k < K; ++k) pb[n * K + k] = hb[n * 64 + k];
    cudaMemcpy(da, pa.data(), pa.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db, pb.data(), pb.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dd, 0, 128 * 256 * 4);
    probe_kernel<<<1, 128, 130 * 1024>>>(p, da, db, dd, dc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", c.name, cudaGetErrorString(e)); return 1; }
    std::vector<float> hd(128 * c.N);
    long long hc[3];
    cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hc, dc, 24, cudaMemcpyDeviceToHost);
    int bad = 0;
    double maxerr = 0;
    for (int m = 0; m < 128; ++m) {
      const int r = c.start + (m / 8) * c.group + (m % 8);
      for (int n = 0; n < c.N; ++n) {
        double ref = 0;
        for (int k = 0; k < K; ++k) ref += (double)pa[r * K + k] * pb[n * K + k];
        const double err = fabs(ref - hd[m * c.N + n]);
        if (err > 1e-3) ++bad;
        if (err > maxerr) maxerr = err;
      }
    }
    const int nmma = 512 * c.ksteps;
    printf("%s N=%3d K=%2d: %s (bad %5d, maxerr %.1f) | issue %.1f cyc/MMA, complete %.1f cyc/MMA (single pass %lld cyc)\n", c.name, c.N, K,
           bad == 0 ? "CORRECT" : "WRONG  ", bad, maxerr, (double)hc[0] / nmma, (double)hc[1] / nmma, hc[2]);
  }
  return 0;
}
