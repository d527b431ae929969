// api.cu — library-level entry points and error plumbing of librsuper_b200.so.
// Contract: 0 / negative return codes, no exceptions across the boundary, no exit,
// thread-local last-error string (forward runs on the Python main thread, backward on the autograd
// engine's worker thread).
#include "rsb_common.cuh"
#include "rsb_tma.cuh"

#include <cstdarg>
#include <cstdio>

#include "../../include/rsuper_b200.h"

namespace rsb {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return -3;
  }
  return 0;
}

// ---- TMA tensor maps ------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encoder() {
  // resolved once per process through the runtime (no link-time dependency on libcuda.so)
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
    return q == cudaDriverEntryPointSuccess ? reinterpret_cast<EncodeTiledFn>(p) : nullptr;
  }();
  return fn;
}

int make_act_tensor_map(CUtensorMap* tm, const void* base, long long pitch, int C, int N, int D, int H, int W, int box_c,
                        int box_x, int box_y, int box_z) {
  EncodeTiledFn enc = get_encoder();
  RSB_REQUIRE(enc != nullptr, "TMA: cuTensorMapEncodeTiled is not available from the driver");
  RSB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && pitch % 8 == 0, "TMA: activation base / pitch must be 16-byte aligned");
  RSB_REQUIRE(box_c == 32 || box_c == 64, "TMA: box_c must be 32 (SWIZZLE_64B) or 64 (SWIZZLE_128B)");
  const cuuint64_t gdim[5] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                              static_cast<cuuint64_t>(D), static_cast<cuuint64_t>(N)};
  const cuuint64_t p2 = static_cast<cuuint64_t>(pitch) * 2;
  const cuuint64_t gstr[4] = {p2, p2 * W, p2 * W * H, p2 * W * H * D};
  const cuuint32_t box[5] = {static_cast<cuuint32_t>(box_c), static_cast<cuuint32_t>(box_x), static_cast<cuuint32_t>(box_y),
                             static_cast<cuuint32_t>(box_z), 1u};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, box_c == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RSB_REQUIRE(r == CUDA_SUCCESS, "TMA: cuTensorMapEncodeTiled failed (CUresult %d; C=%d N=%d D=%d H=%d W=%d pitch=%lld box=%d,%d,%d,%d)",
              static_cast<int>(r), C, N, D, H, W, pitch, box_c, box_x, box_y, box_z);
  return 0;
}

}  // namespace rsb

extern "C" const char* rsb_version(void) { return "rsuper_b200 0.1.0 (sm_100a)"; }

extern "C" const char* rsb_last_error(void) { return rsb::g_last_error; }

extern "C" int rsb_num_sms(void) {
  static thread_local int cached_dev = -1;
  static thread_local int cached_sms = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (dev != cached_dev) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    cached_dev = dev;
    cached_sms = sms;
  }
  return cached_sms;
}
