// infer.cu — sliding-window inference post-processing and bit-packed mask batch assembly (§8f N2).
//
//   sigmoid_window_accumulate   out[:, :, win] += sigmoid(pred); count[:, :, win] += 1      inference/inference3d.py:80-100
//   blend_finalize              prob = out / count ; mask = prob > threshold                  inference3d.py:102, predict_abdomenatlas.py:672
//   dilate_box3                 ndi.binary_dilation(organ, structure = ones((3, 3, 3)))       predict_abdomenatlas.py:675
//   gate                        lesion_prob *= dilated organ                                   predict_abdomenatlas.py:680
//   cc_*                        face-connected components (sitk.ConnectedComponentImageFilter default) and
//                               keep_largest_component                                         predict_abdomenatlas.py:686-710
//   unpack_masks                np.unpackbits(packed, axis = 0)[:C] of the on-disk crop format dataset_abdomenatlas_UFO.py:1006-1015
//
// All of it is HBM-bound byte / index work: one thread per voxel (or per 4 voxels), coalesced along x.  Connected
// components use the lock-free union-find of Komura / Playne-Hawick: every foreground voxel starts as its own root,
// is united with its -x, -y, -z foreground neighbours by atomicMin on the larger root, and a final pass flattens every
// voxel to its root.  Roots are the SMALLEST linear index of their component, so ordering the roots by index reproduces
// the raster-scan label numbering of scipy.ndimage.label / SimpleITK, and "first component of maximal size"
// (predict_abdomenatlas.py:699-705, strict '>') is the maximal (size, -root) key.
#include "rsb_common.cuh"

#include "../../include/rsuper_b200.h"

namespace rsb {

constexpr int INF_THREADS = 256;

static inline int grid_for(long long n, int threads = INF_THREADS) {
  const long long b = (n + threads - 1) / threads;
  const long long cap = 148LL * 32;
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

// ---- sliding-window blend ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(INF_THREADS) sigmoid_window_accumulate_kernel(
    const float* __restrict__ pred, float* __restrict__ out, float* __restrict__ count, int B, int C, int D, int H, int W, int wd,
    int wh, int ww, int d0, int h0, int w0) {
  const long long wv = static_cast<long long>(wd) * wh * ww;
  const long long total = static_cast<long long>(B) * C * wv;
  const long long V = static_cast<long long>(D) * H * W;
  for (long long i = blockIdx.x * static_cast<long long>(INF_THREADS) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * INF_THREADS) {
    const long long bc = i / wv;
    const long long r = i - bc * wv;
    const int z = static_cast<int>(r / (static_cast<long long>(wh) * ww));
    const int rem = static_cast<int>(r - static_cast<long long>(z) * wh * ww);
    const int y = rem / ww, x = rem - y * ww;
    const long long vox = (static_cast<long long>(d0 + z) * H + (h0 + y)) * W + (w0 + x);
    const float p = 1.f / (1.f + expf(-pred[i]));  // torch.sigmoid (ATen: 1 / (1 + exp(-x)) in fp32)
    out[bc * V + vox] += p;
    const int c = static_cast<int>(bc % C);
    if (c == 0) count[(bc / C) * V + vox] += 1.f;
  }
}

__global__ void __launch_bounds__(INF_THREADS) count_window_kernel(float* __restrict__ count, int B, int D, int H, int W, int wd,
                                                                    int wh, int ww, int d0, int h0, int w0) {
  // a gated-out window contributes zeros but still counts in the blend (inference3d.py:92-100)
  const long long wv = static_cast<long long>(wd) * wh * ww;
  const long long total = static_cast<long long>(B) * wv;
  const long long V = static_cast<long long>(D) * H * W;
  for (long long i = blockIdx.x * static_cast<long long>(INF_THREADS) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * INF_THREADS) {
    const long long b = i / wv;
    const long long r = i - b * wv;
    const int z = static_cast<int>(r / (static_cast<long long>(wh) * ww));
    const int rem = static_cast<int>(r - static_cast<long long>(z) * wh * ww);
    const int y = rem / ww, x = rem - y * ww;
    count[b * V + (static_cast<long long>(d0 + z) * H + (h0 + y)) * W + (w0 + x)] += 1.f;
  }
}

__global__ void __launch_bounds__(INF_THREADS) blend_finalize_kernel(const float* __restrict__ acc, const float* __restrict__ count,
                                                                      float* __restrict__ prob, uint8_t* __restrict__ mask,
                                                                      float threshold, int B, int C, long long V) {
  const long long total = static_cast<long long>(B) * C * V;
  for (long long i = blockIdx.x * static_cast<long long>(INF_THREADS) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * INF_THREADS) {
    const long long bc = i / V;
    const long long v = i - bc * V;
    const float p = acc[i] / count[(bc / C) * V + v];
    if (prob != nullptr) prob[i] = p;
    if (mask != nullptr) mask[i] = p > threshold ? 1 : 0;
  }
}

// ---- organ gating --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(INF_THREADS) dilate_box3_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                                   int n_vol, int D, int H, int W) {
  const long long V = static_cast<long long>(D) * H * W;
  const long long total = static_cast<long long>(n_vol) * V;
  for (long long i = blockIdx.x * static_cast<long long>(INF_THREADS) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * INF_THREADS) {
    const long long n = i / V;
    const long long v = i - n * V;
    const int z = static_cast<int>(v / (static_cast<long long>(H) * W));
    const int rem = static_cast<int>(v - static_cast<long long>(z) * H * W);
    const int y = rem / W, x = rem - y * W;
    const uint8_t* s = src + n * V;
    uint8_t hit = 0;
    for (int dz = -1; dz <= 1 && !hit; ++dz) {
      const int zz = z + dz;
      if (zz < 0 || zz >= D) continue;  // border_value = 0 (scipy default)
      for (int dy = -1; dy <= 1 && !hit; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= H) continue;
        const uint8_t* row = s + (static_cast<long long>(zz) * H + yy) * W;
        if ((x > 0 && row[x - 1]) || row[x] || (x + 1 < W && row[x + 1])) hit = 1;
      }
    }
    dst[i] = hit;
  }
}

__global__ void __launch_bounds__(INF_THREADS) gate_kernel(float* __restrict__ prob, const uint8_t* __restrict__ organ, long long n) {
  for (long long i = blockIdx.x * static_cast<long long>(INF_THREADS) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * INF_THREADS)
    prob[i] = organ[i] ? prob[i] : 0.f * prob[i];  // organ.astype(float) * prob: NaN / sign of zero kept like the product
}

// ---- connected components ---------------------------------------------------------------------------------------
RSB_DEVICE int cc_find(const int* L, int i) {
  // parents only ever decrease (atomicMin), so a stale value is still an ancestor; ld.cg keeps the walk off the
  // non-coherent L1 so that it sees the other SMs' links as early as possible
  int p = __ldcg(L + i);
  while (p != i) {
    i = p;
    p = __ldcg(L + i);
  }
  return i;
}

RSB_DEVICE void cc_unite(int* L, int a, int b) {
  bool done;
  do {
    a = cc_find(L, a);
    b = cc_find(L, b);
    if (a < b) {
      const int old = atomicMin(&L[b], a);
      done = (old == b);
      b = old;
    } else if (b < a) {
      const int old = atomicMin(&L[a], b);
      done = (old == a);
      a = old;
    } else {
      done = true;
    }
  } while (!done);
}

__global__ void __launch_bounds__(INF_THREADS) cc_init_kernel(const uint8_t* __restrict__ mask, int* __restrict__ L, int* sizes,
                                                               unsigned long long* best, int* n_comp, int V) {
  for (int i = blockIdx.x * INF_THREADS + threadIdx.x; i < V; i += gridDim.x * INF_THREADS) {
    L[i] = mask[i] ? i : -1;
    if (sizes != nullptr) sizes[i] = 0;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *n_comp = 0;
    if (best != nullptr) *best = 0ull;
  }
}

__global__ void __launch_bounds__(INF_THREADS) cc_merge_kernel(int* L, int D, int H, int W) {
  const int V = D * H * W;
  for (int i = blockIdx.x * INF_THREADS + threadIdx.x; i < V; i += gridDim.x * INF_THREADS) {
    if (L[i] < 0) continue;
    const int x = i % W, y = (i / W) % H, z = i / (W * H);
    if (x > 0 && L[i - 1] >= 0) cc_unite(L, i, i - 1);
    if (y > 0 && L[i - W] >= 0) cc_unite(L, i, i - W);
    if (z > 0 && L[i - W * H] >= 0) cc_unite(L, i, i - W * H);
  }
}

__global__ void __launch_bounds__(INF_THREADS) cc_flatten_kernel(int* L, int* sizes, int* n_comp, int V) {
  // after the merge kernel has COMPLETED every parent chain ends in the component's minimum index; flattening only ever
  // replaces L[i] by that root, so concurrent readers of L[i] still reach the same root
  for (int i = blockIdx.x * INF_THREADS + threadIdx.x; i < V; i += gridDim.x * INF_THREADS) {
    if (L[i] < 0) continue;
    const int r = cc_find(L, i);
    L[i] = r;
    if (r == i) atomicAdd(n_comp, 1);
    if (sizes != nullptr) atomicAdd(&sizes[r], 1);
  }
}

__global__ void __launch_bounds__(INF_THREADS) cc_best_kernel(const int* __restrict__ L, const int* __restrict__ sizes,
                                                               unsigned long long* best, int V) {
  unsigned long long local = 0ull;
  for (int i = blockIdx.x * INF_THREADS + threadIdx.x; i < V; i += gridDim.x * INF_THREADS)
    if (L[i] == i) {  // a root
      const unsigned long long key = (static_cast<unsigned long long>(static_cast<unsigned int>(sizes[i])) << 32) |
                                     (0xFFFFFFFFull - static_cast<unsigned int>(i));
      local = key > local ? key : local;
    }
  if (local) atomicMax(best, local);
}

__global__ void __launch_bounds__(INF_THREADS) cc_select_kernel(const int* __restrict__ L, const unsigned long long* __restrict__ best,
                                                                 uint8_t* __restrict__ out, int V) {
  const unsigned long long key = *best;
  const int root = key ? static_cast<int>(0xFFFFFFFFull - (key & 0xFFFFFFFFull)) : -2;  // no component: all zeros
  for (int i = blockIdx.x * INF_THREADS + threadIdx.x; i < V; i += gridDim.x * INF_THREADS) out[i] = L[i] == root ? 1 : 0;
}

// ---- bit-packed masks ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(INF_THREADS) unpack_masks_kernel(const uint8_t* __restrict__ packed, uint8_t* __restrict__ out,
                                                                    int B, int C, int Cp, long long V, int invert) {
  // numpy packs big-endian: channel c lives in byte plane c >> 3, bit 7 - (c & 7); one thread = one (b, byte plane, voxel):
  // one packed byte read, up to 8 channel bytes written (each channel row coalesced across the warp)
  const long long total = static_cast<long long>(B) * Cp * V;
  for (long long i = blockIdx.x * static_cast<long long>(INF_THREADS) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * INF_THREADS) {
    const long long bp = i / V;
    const long long v = i - bp * V;
    const int b = static_cast<int>(bp / Cp), plane = static_cast<int>(bp - static_cast<long long>(b) * Cp);
    uint32_t bits = packed[i];
    if (invert) bits = ~bits;
    const int c0 = plane * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (c0 + k < C) out[(static_cast<long long>(b) * C + c0 + k) * V + v] = (bits >> (7 - k)) & 1u;
  }
}

}  // namespace rsb

using namespace rsb;

extern "C" int rsb_sigmoid_window_accumulate(const float* pred, float* out, float* count, int B, int C, int D, int H, int W, int wd,
                                             int wh, int ww, int d0, int h0, int w0, void* stream) {
  RSB_REQUIRE(out != nullptr && count != nullptr, "sigmoid_window_accumulate: null accumulator");
  RSB_REQUIRE(B > 0 && C > 0 && wd > 0 && wh > 0 && ww > 0, "sigmoid_window_accumulate: empty window");
  RSB_REQUIRE(d0 >= 0 && h0 >= 0 && w0 >= 0 && d0 + wd <= D && h0 + wh <= H && w0 + ww <= W,
              "sigmoid_window_accumulate: window (%d,%d,%d)+(%d,%d,%d) leaves the volume (%d,%d,%d)", d0, h0, w0, wd, wh, ww, D, H, W);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (pred != nullptr) {
    sigmoid_window_accumulate_kernel<<<grid_for(static_cast<long long>(B) * C * wd * wh * ww), INF_THREADS, 0, st>>>(
        pred, out, count, B, C, D, H, W, wd, wh, ww, d0, h0, w0);
    return check_launch("sigmoid_window_accumulate_kernel");
  }
  count_window_kernel<<<grid_for(static_cast<long long>(B) * wd * wh * ww), INF_THREADS, 0, st>>>(count, B, D, H, W, wd, wh, ww, d0, h0, w0);
  return check_launch("count_window_kernel");
}

extern "C" int rsb_blend_finalize(const float* acc, const float* count, float* prob, uint8_t* mask, float threshold, int B, int C,
                                  long long V, void* stream) {
  RSB_REQUIRE(acc != nullptr && count != nullptr && (prob != nullptr || mask != nullptr), "blend_finalize: null pointer");
  RSB_REQUIRE(B > 0 && C > 0 && V > 0, "blend_finalize: empty volume");
  blend_finalize_kernel<<<grid_for(static_cast<long long>(B) * C * V), INF_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      acc, count, prob, mask, threshold, B, C, V);
  return check_launch("blend_finalize_kernel");
}

extern "C" int rsb_dilate_box3(const uint8_t* src, uint8_t* dst, int n_vol, int D, int H, int W, void* stream) {
  RSB_REQUIRE(src != nullptr && dst != nullptr && src != dst, "dilate_box3: needs distinct source and destination");
  RSB_REQUIRE(n_vol > 0 && D > 0 && H > 0 && W > 0, "dilate_box3: empty volume");
  dilate_box3_kernel<<<grid_for(static_cast<long long>(n_vol) * D * H * W), INF_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      src, dst, n_vol, D, H, W);
  return check_launch("dilate_box3_kernel");
}

extern "C" int rsb_gate_by_mask(float* prob, const uint8_t* organ, long long n, void* stream) {
  RSB_REQUIRE(prob != nullptr && organ != nullptr && n > 0, "gate_by_mask: null pointer / empty volume");
  gate_kernel<<<grid_for(n), INF_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(prob, organ, n);
  return check_launch("gate_kernel");
}

extern "C" size_t rsb_cc_workspace_bytes(int D, int H, int W) {
  const size_t V = static_cast<size_t>(D) * H * W;
  return V * sizeof(int) + 16;  // component sizes + the 64-bit arg-max key
}

extern "C" int rsb_cc_label(const uint8_t* mask, int* labels, int* n_components, uint8_t* largest, void* workspace, int D, int H, int W,
                            void* stream) {
  RSB_REQUIRE(mask != nullptr && labels != nullptr && n_components != nullptr, "cc_label: null pointer");
  RSB_REQUIRE(D > 0 && H > 0 && W > 0 && static_cast<long long>(D) * H * W <= (1LL << 30), "cc_label: volume must have 1 .. 2^30 voxels (32-bit labels)");
  RSB_REQUIRE(largest == nullptr || workspace != nullptr, "cc_label: keep-largest needs the workspace (rsb_cc_workspace_bytes)");
  RSB_REQUIRE(workspace == nullptr || (reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "cc_label: workspace must be 8-byte aligned");
  const int V = D * H * W;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned long long* best = static_cast<unsigned long long*>(workspace);
  int* sizes = workspace != nullptr ? reinterpret_cast<int*>(static_cast<char*>(workspace) + 16) : nullptr;
  const int grid = grid_for(V);
  cc_init_kernel<<<grid, INF_THREADS, 0, st>>>(mask, labels, sizes, best, n_components, V);
  if (int rc = check_launch("cc_init_kernel")) return rc;
  cc_merge_kernel<<<grid, INF_THREADS, 0, st>>>(labels, D, H, W);
  if (int rc = check_launch("cc_merge_kernel")) return rc;
  cc_flatten_kernel<<<grid, INF_THREADS, 0, st>>>(labels, sizes, n_components, V);
  if (int rc = check_launch("cc_flatten_kernel")) return rc;
  if (largest != nullptr) {
    cc_best_kernel<<<grid, INF_THREADS, 0, st>>>(labels, sizes, best, V);
    if (int rc = check_launch("cc_best_kernel")) return rc;
    cc_select_kernel<<<grid, INF_THREADS, 0, st>>>(labels, best, largest, V);
    if (int rc = check_launch("cc_select_kernel")) return rc;
  }
  return 0;
}

extern "C" int rsb_unpack_masks(const uint8_t* packed, uint8_t* out, int B, int C, long long V, int invert, void* stream) {
  RSB_REQUIRE(packed != nullptr && out != nullptr, "unpack_masks: null pointer");
  RSB_REQUIRE(B > 0 && C > 0 && V > 0, "unpack_masks: empty batch");
  const int Cp = (C + 7) / 8;
  unpack_masks_kernel<<<grid_for(static_cast<long long>(B) * Cp * V), INF_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      packed, out, B, C, Cp, V, invert);
  return check_launch("unpack_masks_kernel");
}
