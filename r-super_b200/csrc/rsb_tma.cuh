// rsb_tma.cuh — tensor-map (TMA) plumbing for the conv kernels.
//
// Activations are channels-last bf16 [N][D][H][W][pitch] views (a channel slice of a wider buffer is
// fine: the slice start is the base address, the pitch is the voxel stride).  They are described to
// the TMA unit as a 5-D tensor (c, x, y, z, n) and fetched as boxes (CB, bx, by, bz, 1) whose rows
// (CB = 32 or 64 channels = 64 / 128 bytes) land in shared memory with the hardware 64B / 128B
// swizzle — exactly the K-major (fprop / dgrad A operand) resp. MN-major (wgrad operands) UMMA
// layouts, so that no thread ever touches the operand bytes.  Out-of-range coordinates (negative
// or past the end, per dimension and therefore per SAMPLE in z) are zero-filled by the hardware:
// that is the conv's zero padding and the ragged-tile handling.
//
// Measured on B200 (profiles/r01_tma_probe_*.log): 128-byte rows stream at 64 B/clk/SM, 64-byte
// rows at 43 B/clk/SM; the 16-byte-piece boxes needed for un-swizzled core-matrix layouts only
// reach 16 B/clk/SM (7 with all SMs active), which is why the swizzled row layouts are used.
#pragma once
#include <cuda.h>  // CUtensorMap + enums only; the encoder is resolved at run time (no libcuda link)
#include <cuda_runtime.h>
#include <stdint.h>

#include "rsb_common.cuh"

namespace rsb {

// ---- host ------------------------------------------------------------------------------------
// Encodes the 5-D activation map.  Returns 0 or a negative code (rsb_last_error is set).
int make_act_tensor_map(CUtensorMap* tm, const void* base, long long pitch_elems, int C, int N, int D, int H, int W,
                        int box_c, int box_x, int box_y, int box_z);

// ---- device ----------------------------------------------------------------------------------
RSB_DEVICE void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

// 5-D tiled load global -> shared, completion (bytes) on an mbarrier.  SASS: UTMALDG.5D
RSB_DEVICE void tma_load_5d(uint32_t dst_smem, const CUtensorMap* tm, int c, int x, int y, int z, int n, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(dst_smem), "l"(tm), "r"(c), "r"(x), "r"(y), "r"(z), "r"(n), "r"(bar)
      : "memory");
}

// Shared-memory matrix descriptor for the swizzled layouts (sm_100: version 1 at bit 46,
// layout_type [61,64): 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 0 = none).
//   K-major  (rows = M/N index, row payload = K):   SBO = byte stride between 8-row groups, LBO unused
//   MN-major (rows = K index, row payload = M/N):   LBO = byte stride between MN groups (one row payload
//            each), SBO = byte stride between 8-row K groups          (profiles/r01_umma_mnmajor_layout_probe.log)
// The swizzle is a function of absolute shared-memory address bits, so a descriptor may start at any row
// of a TMA-written tile (a filter tap is a row shift) — profiles/r01_umma_kmajor_layout_probe.log.
RSB_DEVICE uint32_t desc_hi(uint32_t sbo_bytes, uint32_t layout_type) {
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (layout_type << 29);
}
RSB_DEVICE uint32_t desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
RSB_DEVICE uint64_t desc_join(uint32_t hi, uint32_t lo) { return (static_cast<uint64_t>(hi) << 32) | lo; }

constexpr uint32_t kLayoutSw128 = 2, kLayoutSw64 = 4, kLayoutNone = 0;

}  // namespace rsb
