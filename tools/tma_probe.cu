// tma_probe.cu — micro-experiment: 5-D tiled TMA loads of an NDHWC (channel-pitched) bf16 activation
// straight into the SWIZZLE_NONE UMMA core-matrix layouts the conv kernels consume:
//   order A (fprop / wgrad dy):  box dims (c8, x, y, cg, z')  -> smem [cg][y][x][8ch]
//   order B (wgrad a-plane):     box dims (c8, x, cg, y, z')  -> smem [y][cg][x][8ch]
// Checks (1) cuTensorMapEncodeTiled accepts the non-monotonic strides (cg stride 16 B < x stride),
// (2) the landing layout incl. zero fill for negative / out-of-range coordinates, (3) cycles per box.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/tma_probe tools/tma_probe.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) tma_kernel(const __grid_constant__ CUtensorMap tm, int c1, int c2, int c3, int c4,
                                                     uint32_t box_bytes, int iters, int depth, uint16_t* dump, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[8];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int it = 0; it < iters + depth; ++it) {
      if (it >= depth) {
        const int j = it - depth;
        const uint32_t b = smem_u32(&bar[j % depth]);
        const uint32_t par = (j / depth) & 1;
        uint32_t ok = 0;
        while (!ok) {
          asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
                       : "=r"(ok) : "r"(b), "r"(par) : "memory");
        }
      }
      if (it < iters) {
        const uint32_t b = smem_u32(&bar[it % depth]);
        const uint32_t dst = smem_u32(smem) + (it % depth) * ((box_bytes + 1023) / 1024 * 1024);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(box_bytes) : "memory");
        // vary z' with the iteration and the CTA so loads are not all the same lines
        const int z = c4 + ((it + blockIdx.x) & 1);  // second sample / plane alternately
        asm volatile(
            "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
            ::"r"(dst), "l"(&tm), "r"(0), "r"(c1), "r"(c2), "r"(c3), "r"(iters == 1 ? c4 : z), "r"(b)
            : "memory");
      }
    }
    cyc[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  if (dump != nullptr && blockIdx.x == 0)
    for (uint32_t i = threadIdx.x; i < box_bytes / 2; i += blockDim.x) dump[i] = reinterpret_cast<uint16_t*>(smem)[i];
}

int main() {
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres) != cudaSuccess || !encode) {
    printf("cuTensorMapEncodeTiled not found\n");
    return 1;
  }
  const int N = 2, D = 6, H = 20, W = 12, pitch = 48, c0 = 8, C = 32;
  const size_t elems = (size_t)N * D * H * W * pitch;
  std::vector<uint16_t> h(elems);
  // value = exact small integer in bf16: encode (z', y, x, c) compactly: bf16 holds 8 bits mantissa -> use a hash < 256
  auto val = [&](int zp, int y, int x, int c) { return (float)(((zp * 31 + y * 17 + x * 7 + c * 3) % 251) + 1); };
  for (int zp = 0; zp < N * D; ++zp)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x)
        for (int c = 0; c < pitch; ++c) {
          __nv_bfloat16 b = __float2bfloat16(val(zp, y, x, c));
          h[(((size_t)zp * H + y) * W + x) * pitch + c] = *reinterpret_cast<uint16_t*>(&b);
        }
  uint16_t* d;
  cudaMalloc(&d, elems * 2);
  cudaMemcpy(d, h.data(), elems * 2, cudaMemcpyHostToDevice);
  uint16_t* ddump;
  long long* dcyc;
  cudaMalloc(&ddump, 256 * 1024);
  cudaMalloc(&dcyc, 148 * 8);
  cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);

  for (int order = 0; order < 2; ++order) {
    // order 0: (c8, x, y, cg, z')   order 1: (c8, x, cg, y, z')
    const int bx = 10, by = 18, bcg = 4, bz = (order == 0) ? 2 : 1;
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
    gdim[0] = 8; box[0] = 8;
    gdim[1] = W; gstr[0] = pitch * 2; box[1] = bx;
    if (order == 0) {
      gdim[2] = H; gstr[1] = (cuuint64_t)W * pitch * 2; box[2] = by;
      gdim[3] = C / 8; gstr[2] = 16; box[3] = bcg;
    } else {
      gdim[2] = C / 8; gstr[1] = 16; box[2] = bcg;
      gdim[3] = H; gstr[2] = (cuuint64_t)W * pitch * 2; box[3] = by;
    }
    gdim[4] = N * D; gstr[3] = (cuuint64_t)H * W * pitch * 2; box[4] = bz;
    CUtensorMap tm;
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d + c0, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("order %d: encode -> %d\n", order, (int)r);
    if (r != CUDA_SUCCESS) continue;
    const uint32_t box_bytes = 8 * bx * by * bcg * bz * 2;
    // correctness: box starting at x=-1, y=7 (runs past H), z'=3
    const int x0 = -1, y0 = 7, z0 = 3;
    int c1 = x0, c2 = order == 0 ? y0 : 0, c3 = order == 0 ? 0 : y0;
    tma_kernel<<<1, 128, 200 * 1024>>>(tm, c1, c2, c3, z0, box_bytes, 1, 1, ddump, dcyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("  kernel error %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<uint16_t> out(box_bytes / 2);
    cudaMemcpy(out.data(), ddump, box_bytes, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int z = 0; z < bz; ++z)
      for (int cg = 0; cg < bcg; ++cg)
        for (int yy = 0; yy < by; ++yy)
          for (int xx = 0; xx < bx; ++xx)
            for (int c = 0; c < 8; ++c) {
              size_t idx = order == 0 ? ((((size_t)z * bcg + cg) * by + yy) * bx + xx) * 8 + c
                                      : ((((size_t)z * by + yy) * bcg + cg) * bx + xx) * 8 + c;
              const int x = x0 + xx, y = y0 + yy, zp = z0 + z;
              float want = 0.f;
              if (x >= 0 && x < W && y >= 0 && y < H && zp < N * D) want = val(zp, y, x, c0 + cg * 8 + c);
              __nv_bfloat16 b = *reinterpret_cast<__nv_bfloat16*>(&out[idx]);
              if (__bfloat162float(b) != want) ++bad;
            }
    printf("  layout check: %s (bad %d of %u)\n", bad == 0 ? "CORRECT" : "WRONG", bad, box_bytes / 2);
    // throughput: 1 CTA and 148 CTAs, pipeline depth 4
    for (int grid : {1, 148}) {
      for (int depth : {2, 4, 8}) {
        const int iters = 2000;
        tma_kernel<<<grid, 128, 200 * 1024>>>(tm, 0, order == 0 ? 1 : 0, order == 0 ? 0 : 1, 2, box_bytes, iters, depth, nullptr, dcyc);
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("  kernel error %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<long long> hc(grid);
        cudaMemcpy(hc.data(), dcyc, grid * 8, cudaMemcpyDeviceToHost);
        double mx = 0;
        for (long long v : hc) mx = v > mx ? v : mx;
        printf("  grid %3d depth %d: %.0f cycles per %u-byte box (%.1f B/clk/SM, %u 16-B pieces)\n", grid, depth, mx / iters, box_bytes,
               box_bytes / (mx / iters), box_bytes / 16);
      }
    }
  }
  // dense channel rows (c, x, y, z, n) with hardware swizzle: rows of 128 B (64 ch, SWIZZLE_128B) / 64 B (32 ch, SWIZZLE_64B)
  // expected landing pattern: 16-byte chunk index XORed with absolute smem address bits [7:9] (SW128) / [7:8] (SW64)
  for (int mode = 0; mode < 3; ++mode) {
    const int CC = (mode == 1) ? 32 : 64;  // mode 0: 64ch SW128, 1: 32ch SW64, 2: 64ch no swizzle
    const int pitch2 = 96, cc0 = 16;
    const size_t elems2 = (size_t)N * D * H * W * pitch2;
    std::vector<uint16_t> h2(elems2);
    for (int zp = 0; zp < N * D; ++zp)
      for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
          for (int c = 0; c < pitch2; ++c) {
            __nv_bfloat16 b = __float2bfloat16(val(zp, y, x, c));
            h2[(((size_t)zp * H + y) * W + x) * pitch2 + c] = *reinterpret_cast<uint16_t*>(&b);
          }
    uint16_t* d2;
    cudaMalloc(&d2, elems2 * 2);
    cudaMemcpy(d2, h2.data(), elems2 * 2, cudaMemcpyHostToDevice);
    const int bx = 10, by = 18, bz = 3;
    cuuint64_t gdim[5] = {(cuuint64_t)CC, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t gstr[4] = {(cuuint64_t)pitch2 * 2, (cuuint64_t)W * pitch2 * 2, (cuuint64_t)H * W * pitch2 * 2, (cuuint64_t)D * H * W * pitch2 * 2};
    cuuint32_t box[5] = {(cuuint32_t)CC, bx, by, bz, 1}, estr[5] = {1, 1, 1, 1, 1};
    CUtensorMap tm;
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d2 + cc0, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        mode == 0 ? CU_TENSOR_MAP_SWIZZLE_128B : (mode == 1 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE),
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("dense rows mode %d (%d ch): encode -> %d\n", mode, CC, (int)r);
    if (r != CUDA_SUCCESS) continue;
    const uint32_t row_bytes = CC * 2;
    const uint32_t box_bytes = row_bytes * bx * by * bz;
    const int x0 = -1, y0 = 7, z0 = 4, n0 = 1;  // z runs past D=6: planes 4,5 valid, 6 zero-filled (not the next sample)
    tma_kernel<<<1, 128, 200 * 1024>>>(tm, x0, y0, z0, n0, box_bytes, 1, 1, ddump, dcyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("  kernel error %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<uint16_t> out(box_bytes / 2);
    cudaMemcpy(out.data(), ddump, box_bytes, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int z = 0; z < bz; ++z)
      for (int yy = 0; yy < by; ++yy)
        for (int xx = 0; xx < bx; ++xx)
          for (int c = 0; c < CC; ++c) {
            const uint32_t row = (z * by + yy) * bx + xx;
            const uint32_t raddr = row * row_bytes;  // box base is 1024-aligned -> relative == absolute low bits
            uint32_t chunk = c / 8;
            if (mode == 0) chunk ^= (raddr >> 7) & 7;
            if (mode == 1) chunk ^= (raddr >> 7) & 3;
            const size_t idx = (raddr + chunk * 16) / 2 + (c % 8);
            const int x = x0 + xx, y = y0 + yy, zz = z0 + z;
            float want = 0.f;
            if (x >= 0 && x < W && y >= 0 && y < H && zz < D) want = val(n0 * D + zz, y, x, cc0 + c);
            __nv_bfloat16 b = *reinterpret_cast<__nv_bfloat16*>(&out[idx]);
            if (__bfloat162float(b) != want) ++bad;
          }
    printf("  layout check (absolute-address XOR swizzle, per-sample z zero fill): %s (bad %d of %u)\n", bad == 0 ? "CORRECT" : "WRONG", bad,
           box_bytes / 2);
    for (int grid : {1, 148}) {
      const int iters = 2000, depth = 2;
      tma_kernel<<<grid, 128, 200 * 1024>>>(tm, 0, 1, 2, 0, box_bytes, iters, depth, nullptr, dcyc);
      e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("  kernel error %s\n", cudaGetErrorString(e)); return 1; }
      std::vector<long long> hc(grid);
      cudaMemcpy(hc.data(), dcyc, grid * 8, cudaMemcpyDeviceToHost);
      double mx = 0;
      for (long long v : hc) mx = v > mx ? v : mx;
      printf("  grid %3d depth %d: %.0f cycles per %u-byte box (%.1f B/clk/SM, %u rows of %u B)\n", grid, depth, mx / iters, box_bytes,
             box_bytes / (mx / iters), box_bytes / row_bytes, row_bytes);
    }
  }
  return 0;
}
