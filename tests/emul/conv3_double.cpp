// conv3_double.cpp — TEST INFRASTRUCTURE ONLY: a naive host stand-in for the tcgen05 / TMA kernels (rsb_conv3_forward,
// rsb_conv3_wgrad, rsb_conv3_pack_weights[_batched]) behind the SAME C-ABI, so that the whole train step — the engine's host
// logic, every HBM-bound kernel (run from its real source through cuda_emul.h), the losses and the fused optimizer — can be
// executed and checked against the oracle in the GPU-less container.  The tensor-core kernels themselves are NOT tested by
// this (their parity evidence is the GPU suite); the double only reproduces their contract as include/rsuper_b200.h states it:
// bf16 operands, fp32 accumulation, split-precision parts, residual / statistics / dgrad-mask epilogues.  The "packed" image
// is opaque to the host code, so the double uses its own layout inside it: planes [hi][lo] of bf16 [Cout'][Cin'][27].
#include "cuda_emul.h"

#include "../../include/rsuper_b200.h"

namespace {

inline float bf(const __nv_bfloat16* p, long long i) { return __bfloat162float(p[i]); }

void pack_one(const float* w_a, const float* w_b, int rows_a, int Cout, int Cin, int flip, int parts, __nv_bfloat16* out, int pointwise = 0) {
  const int co_eff = flip ? Cin : Cout, ci_eff = flip ? Cout : Cin;
  const long long plane = 27LL * co_eff * ci_eff;
  const int taps = pointwise ? 1 : 27;   // 1x1x1 sources are embedded at the centre tap (13), the other taps are zero
  for (int co = 0; co < Cout; ++co) {
    const float* row = co < rows_a ? w_a + static_cast<long long>(co) * Cin * taps : w_b + static_cast<long long>(co - rows_a) * Cin * taps;
    for (int ci = 0; ci < Cin; ++ci)
      for (int t = 0; t < 27; ++t) {
        const float w = pointwise ? (t == 13 ? row[ci] : 0.f) : row[ci * 27 + t];
        // dgrad image: conv with Cin' = Cout, Cout' = Cin and the taps flipped in all three axes
        const long long dst = flip ? (static_cast<long long>(ci) * ci_eff + co) * 27 + (26 - t) : (static_cast<long long>(co) * ci_eff + ci) * 27 + t;
        const __nv_bfloat16 hi = __float2bfloat16_rn(w);
        out[dst] = hi;
        if (parts >= 3) out[plane + dst] = __float2bfloat16_rn(w - __bfloat162float(hi));
      }
  }
}

template <typename T> inline float ld(const T* p, long long i);
template <> inline float ld<float>(const float* p, long long i) { return p[i]; }
template <> inline float ld<__nv_bfloat16>(const __nv_bfloat16* p, long long i) { return __bfloat162float(p[i]); }
template <typename T> inline void st(T* p, long long i, float v);
template <> inline void st<float>(float* p, long long i, float v) { p[i] = v; }
template <> inline void st<__nv_bfloat16>(__nv_bfloat16* p, long long i, float v) { p[i] = __float2bfloat16_rn(v); }

template <typename T>
int conv_forward(const RsbConv3Args* a) {
  const int N = a->N, D = a->D, H = a->H, W = a->W, Ci = a->Cin, Co = a->Cout;
  const long long V = static_cast<long long>(D) * H * W;
  const __nv_bfloat16* A = static_cast<const __nv_bfloat16*>(a->a);
  const __nv_bfloat16* Al = static_cast<const __nv_bfloat16*>(a->a_lo);
  const __nv_bfloat16* Wh = static_cast<const __nv_bfloat16*>(a->w_packed);
  const __nv_bfloat16* Wl = Wh + 27LL * Co * Ci;
  T* Y = static_cast<T*>(a->y);
  const T* aux = static_cast<const T*>(a->mask_x ? a->mask_x : a->res);
  const long long auxp = a->mask_x ? a->mask_x_pitch : a->res_pitch;
  float* stat = a->mask_x ? a->bwd_sums : a->out_stats;
  const long long statp = a->mask_x ? a->mask_x_pitch : a->y_pitch;
  std::vector<float> acc(Co);
  for (int n = 0; n < N; ++n)
    for (int z = 0; z < D; ++z)
      for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
          std::fill(acc.begin(), acc.end(), 0.f);
          for (int kd = 0; kd < 3; ++kd) {
            const int zz = z + kd - 1;
            if (zz < 0 || zz >= D) continue;
            for (int kh = 0; kh < 3; ++kh) {
              const int yy = y + kh - 1;
              if (yy < 0 || yy >= H) continue;
              for (int kw = 0; kw < 3; ++kw) {
                const int xx = x + kw - 1;
                if (xx < 0 || xx >= W) continue;
                const int t = kd * 9 + kh * 3 + kw;
                const long long src = ((static_cast<long long>(n) * D + zz) * H + yy) * W + xx;
                for (int ci = 0; ci < Ci; ++ci) {
                  const float ah = bf(A, src * a->a_pitch + ci);
                  const float al = Al ? bf(Al, src * a->a_pitch + ci) : 0.f;
                  if (ah == 0.f && al == 0.f) continue;
                  for (int co = 0; co < Co; ++co) {
                    const long long wi = (static_cast<long long>(co) * Ci + ci) * 27 + t;
                    float p = ah * bf(Wh, wi);
                    if (Al) p += al * bf(Wh, wi) + ah * bf(Wl, wi);   // a_hi*w_hi + a_lo*w_hi + a_hi*w_lo
                    acc[co] += p;
                  }
                }
              }
            }
          }
          const long long vox = ((static_cast<long long>(n) * D + z) * H + y) * W + x;
          for (int co = 0; co < Co; ++co) {
            float v = acc[co];
            if (a->mask_x) {
              const float* ms = a->mask_stats + (static_cast<long long>(n) * a->mask_x_pitch + co) * 2;
              const float mean = ms[0] / V;
              const float var = std::max(ms[1] / V - mean * mean, 0.f);
              const float h = (ld<T>(aux, vox * auxp + co) - mean) / std::sqrt(var + a->eps);
              v = h > 0.f ? v : v * a->slope;
              stat[(static_cast<long long>(n) * statp + co) * 2 + 0] += v;
              stat[(static_cast<long long>(n) * statp + co) * 2 + 1] += v * h;
            } else {
              if (a->res) v += ld<T>(aux, vox * auxp + co);
              if (stat) {
                stat[(static_cast<long long>(n) * statp + co) * 2 + 0] += v;
                stat[(static_cast<long long>(n) * statp + co) * 2 + 1] += v * v;
              }
            }
            st<T>(Y, vox * a->y_pitch + co, v);
          }
        }
  return 0;
}

}  // namespace

extern "C" int rsb_conv3_pack_weights(const float* w, void* packed, int Cout, int Cin, int transpose_flip, int parts, void*) {
  RSB_REQUIRE(w && packed && (parts == 1 || parts == 3), "conv3 double: pack supports parts 1 and 3");
  pack_one(w, nullptr, Cout, Cout, Cin, transpose_flip, parts, static_cast<__nv_bfloat16*>(packed));
  return 0;
}

extern "C" int rsb_conv3_pack_weights_batched(const RsbPackJob* jobs, int n_jobs, unsigned int, void*) {
  RSB_REQUIRE(jobs && n_jobs > 0, "conv3 double: null job table");
  for (int j = 0; j < n_jobs; ++j) {
    const RsbPackJob& b = jobs[j];
    RSB_REQUIRE(b.parts == 1 || b.parts == 3, "conv3 double: pack supports parts 1 and 3 (got %d)", b.parts);
    pack_one(b.w_a, b.w_b, b.rows_a, b.Cout, b.Cin, b.transpose_flip, b.parts, static_cast<__nv_bfloat16*>(b.packed), b.pointwise);
  }
  return 0;
}

extern "C" int rsb_conv3_forward(const RsbConv3Args* a, void*) {
  RSB_REQUIRE(a && a->a && a->w_packed && a->y, "conv3 double: null pointer");
  RSB_REQUIRE(!a->a_lo2, "conv3 double: the three-piece product is not modelled");
  RSB_REQUIRE(!(a->mask_x && (a->out_stats || a->res)), "conv3: the dgrad mask epilogue excludes out_stats / res");
  return a->dtype == RSB_BF16 ? conv_forward<__nv_bfloat16>(a) : conv_forward<float>(a);
}

extern "C" int rsb_conv3_wgrad(const RsbConv3WgradArgs* g, void*) {
  RSB_REQUIRE(g && g->a && g->dy && g->dw_oidhw, "conv3 double: null pointer");
  const int N = g->N, D = g->D, H = g->H, W = g->W, Ci = g->Cin, Co = g->Cout;
  const __nv_bfloat16* A = static_cast<const __nv_bfloat16*>(g->a);
  const __nv_bfloat16* DY = static_cast<const __nv_bfloat16*>(g->dy);
  std::vector<double> dw(static_cast<size_t>(Co) * Ci * 27, 0.0);
  for (int n = 0; n < N; ++n)
    for (int z = 0; z < D; ++z)
      for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
          const long long vox = ((static_cast<long long>(n) * D + z) * H + y) * W + x;
          for (int kd = 0; kd < 3; ++kd) {
            const int zz = z + kd - 1;
            if (zz < 0 || zz >= D) continue;
            for (int kh = 0; kh < 3; ++kh) {
              const int yy = y + kh - 1;
              if (yy < 0 || yy >= H) continue;
              for (int kw = 0; kw < 3; ++kw) {
                const int xx = x + kw - 1;
                if (xx < 0 || xx >= W) continue;
                const int t = kd * 9 + kh * 3 + kw;
                const long long src = ((static_cast<long long>(n) * D + zz) * H + yy) * W + xx;
                for (int co = 0; co < Co; ++co) {
                  const float d = bf(DY, vox * g->dy_pitch + co);
                  if (d == 0.f) continue;
                  for (int ci = 0; ci < Ci; ++ci) dw[(static_cast<size_t>(co) * Ci + ci) * 27 + t] += static_cast<double>(d) * bf(A, src * g->a_pitch + ci);
                }
              }
            }
          }
        }
  for (size_t i = 0; i < dw.size(); ++i) g->dw_oidhw[i] = (g->accumulate ? g->dw_oidhw[i] : 0.f) + static_cast<float>(dw[i]);
  return 0;
}
