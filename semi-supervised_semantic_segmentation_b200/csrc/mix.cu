// Fused CowMix mixing: images and predictions are mixed with the same mask in one pass.
// Reference semantics: cowmix.py:72-73, called twice per step (train.py:82 and :84-86); each call
// is 4 ATen kernels with 3 temporaries there (~28 B/element moved), here 12 B/element + the mask
// once per pixel.
//
// Roofline: HBM-bound, 4*(3*c0 + 3*c1 + 1) bytes per pixel.
#include "common.cuh"
#include "bilinear.cuh"

namespace b200ssl {

constexpr int kMixThreads = 256;
constexpr int kMixChanUnroll = 2;

// RN(RN(a*m) + RN(b*RN(1-m))): the exact op sequence of `a*mask + b*(1.-mask)`; no FMA contraction.
__device__ __forceinline__ float mix_one(float a, float b, float m, float om) {
  return __fadd_rn(__fmul_rn(a, m), __fmul_rn(b, om));
}

template <int VEC>
struct Pack;
template <>
struct Pack<4> {
  float4 v;
  __device__ __forceinline__ void load(const float* p) { v = ld_stream_f4(p); }
  __device__ __forceinline__ void store(float* p) const { st_stream_f4(p, v); }
  __device__ __forceinline__ float& at(int i) { return (&v.x)[i]; }
};
template <>
struct Pack<1> {
  float v;
  __device__ __forceinline__ void load(const float* p) { v = ld_stream_f1(p); }
  __device__ __forceinline__ void store(float* p) const { *p = v; }
  __device__ __forceinline__ float& at(int) { return v; }
};

template <int VEC, bool CHANNEL_MASK>
__device__ __forceinline__ void mix_tensor(const float* __restrict__ a, const float* __restrict__ b,
                                           float* __restrict__ out, int c, long long n_idx,
                                           long long off, long long hw, Pack<VEC> m, Pack<VEC> om,
                                           const float* __restrict__ mask) {
  const long long base = n_idx * c * hw + off;
  for (int j0 = 0; j0 < c; j0 += kMixChanUnroll) {
    Pack<VEC> av[kMixChanUnroll], bv[kMixChanUnroll], mv[kMixChanUnroll];
#pragma unroll
    for (int u = 0; u < kMixChanUnroll; ++u) {
      if (j0 + u < c) {
        av[u].load(a + base + (long long)(j0 + u) * hw);
        bv[u].load(b + base + (long long)(j0 + u) * hw);
        if (CHANNEL_MASK) mv[u].load(mask + base + (long long)(j0 + u) * hw);
      }
    }
#pragma unroll
    for (int u = 0; u < kMixChanUnroll; ++u) {
      if (j0 + u < c) {
        Pack<VEC> o;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          if (CHANNEL_MASK) {
            const float mm = mv[u].at(e);
            o.at(e) = mix_one(av[u].at(e), bv[u].at(e), mm, __fsub_rn(1.0f, mm));
          } else {
            o.at(e) = mix_one(av[u].at(e), bv[u].at(e), m.at(e), om.at(e));
          }
        }
        o.store(out + base + (long long)(j0 + u) * hw);
      }
    }
  }
}

// FIELD: `mask` is the smoothed field S and the {0,1} mask is formed on the fly as S > tau[image]
// (cowmix.py:68) and written to mask_out -- the threshold pass fused into the mix.
template <int VEC, bool CHANNEL_MASK, bool FIELD>
__global__ void __launch_bounds__(kMixThreads, 4)
mix2_kernel(const float* __restrict__ a0, const float* __restrict__ b0, float* __restrict__ out0,
            int c0, const float* __restrict__ a1, const float* __restrict__ b1,
            float* __restrict__ out1, int c1, const float* __restrict__ mask, long long n,
            long long hw, const float* __restrict__ tau, float* __restrict__ mask_out) {
  const long long per_img = hw / VEC;
  const long long total = n * per_img;
  for (long long q = (long long)blockIdx.x * kMixThreads + threadIdx.x; q < total;
       q += (long long)gridDim.x * kMixThreads) {
    const long long n_idx = q / per_img;
    const long long off = (q - n_idx * per_img) * VEC;
    Pack<VEC> m, om;
    if (!CHANNEL_MASK) {
      m.load(mask + n_idx * hw + off);
      if (FIELD) {
        const float t = __ldg(tau + n_idx);
#pragma unroll
        for (int e = 0; e < VEC; ++e) m.at(e) = m.at(e) > t ? 1.0f : 0.0f;
        m.store(mask_out + n_idx * hw + off);
      }
#pragma unroll
      for (int e = 0; e < VEC; ++e) om.at(e) = __fsub_rn(1.0f, m.at(e));
    }
    mix_tensor<VEC, CHANNEL_MASK>(a0, b0, out0, c0, n_idx, off, hw, m, om, mask);
    if (c1 > 0) mix_tensor<VEC, false>(a1, b1, out1, c1, n_idx, off, hw, m, om, mask);
  }
}

// ---- row N2: bilinear upsampling of the teacher predictions fused into the mix -----------
// train.py:71-75 materialises F.interpolate(ema_pred, image size, 'bilinear', align_corners=False) for
// both teacher predictions (a full-resolution [N,C,H,W] write + read each) before mixing them
// (train.py:82).  Here the mix reads the LOW-resolution predictions (16x fewer bytes at HRNet's stride 4,
// L2-resident) and interpolates in registers.  Arithmetic = ATen's upsample_bilinear2d, bit for bit on
// up-sampling and same-size shapes [probed against the CPU kernel]:
//   s  = max(fma(in/out, dst + 0.5, -0.5), 0);  i0 = min(floor(s), in-1);  l = clamp(s - i0, 0, 1)
//   v  = fma(1-ly, fma(1-lx, v00, RN(lx*v01)), RN(ly * fma(1-lx, v10, RN(lx*v11))))
// One thread: VEC consecutive pixels of one output row.  a0/b0 (images) are full resolution; a1/b1 are
// [n, c1, h_in, w_in].  a1 only (b1 == nullptr, out0 unused): plain up-sampling into out1.
template <int VEC, bool FIELD>
__global__ void __launch_bounds__(kMixThreads, 2)
mix2_upsampled_kernel(const float* __restrict__ a0, const float* __restrict__ b0, float* __restrict__ out0, int c0,
                      const float* __restrict__ a1, const float* __restrict__ b1, float* __restrict__ out1, int c1,
                      int h_in, int w_in, const float* __restrict__ mask, long long n, int h, int w,
                      const float* __restrict__ tau, float* __restrict__ mask_out) {
  const long long hw = (long long)h * w, hw_in = (long long)h_in * w_in;
  const int per_row = w / VEC;
  const long long total = n * h * per_row;
  const float sy = (float)h_in / (float)h, sx = (float)w_in / (float)w;
  for (long long q = (long long)blockIdx.x * kMixThreads + threadIdx.x; q < total;
       q += (long long)gridDim.x * kMixThreads) {
    const long long row = q / per_row;
    const int x = (int)(q - row * per_row) * VEC;
    const long long n_idx = row / h;
    const int y = (int)(row - n_idx * h);
    const long long off = (long long)y * w + x;
    Pack<VEC> m, om;
    if (mask) {
      m.load(mask + n_idx * hw + off);
      if (FIELD) {
        const float t = __ldg(tau + n_idx);
#pragma unroll
        for (int e = 0; e < VEC; ++e) m.at(e) = m.at(e) > t ? 1.0f : 0.0f;
        m.store(mask_out + n_idx * hw + off);
      }
#pragma unroll
      for (int e = 0; e < VEC; ++e) om.at(e) = __fsub_rn(1.0f, m.at(e));
    }
    if (c0 > 0) mix_tensor<VEC, false>(a0, b0, out0, c0, n_idx, off, hw, m, om, mask);
    const AxisTap ty = axis_tap(y, h_in, sy);
    AxisTap tx[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) tx[e] = axis_tap(x + e, w_in, sx);
    for (int j = 0; j < c1; ++j) {
      const long long plane = (n_idx * c1 + j) * hw_in;
      const float* a_r0 = a1 + plane + (long long)ty.i0 * w_in;
      const float* a_r1 = a1 + plane + (long long)ty.i1 * w_in;
      Pack<VEC> o;
      if (b1) {
        const float* b_r0 = b1 + plane + (long long)ty.i0 * w_in;
        const float* b_r1 = b1 + plane + (long long)ty.i1 * w_in;
#pragma unroll
        for (int e = 0; e < VEC; ++e)
          o.at(e) = mix_one(bilerp(a_r0, a_r1, tx[e], ty.w0, ty.w1), bilerp(b_r0, b_r1, tx[e], ty.w0, ty.w1),
                            m.at(e), om.at(e));
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) o.at(e) = bilerp(a_r0, a_r1, tx[e], ty.w0, ty.w1);
      }
      o.store(out1 + (n_idx * c1 + j) * hw + off);
    }
  }
}

}  // namespace b200ssl

static int mix2_upsampled_launch(const float* a0, const float* b0, float* out0, int c0, const float* a1,
                                 const float* b1, float* out1, int c1, int h_in, int w_in, const float* mask,
                                 const float* tau, float* mask_out, int64_t n, int h, int w,
                                 b200ssl_stream_t stream, const char* who) {
  using namespace b200ssl;
  B200SSL_REQUIRE(n >= 0 && h >= 0 && w >= 0 && c0 >= 0 && c1 >= 0, "%s: negative extent", who);
  if (n == 0 || h == 0 || w == 0) return 0;
  B200SSL_REQUIRE(c1 == 0 || (a1 && out1 && h_in >= 1 && w_in >= 1), "%s: null or empty low-resolution tensor", who);
  B200SSL_REQUIRE(c0 == 0 || (a0 && b0 && out0), "%s: null tensor 0", who);
  B200SSL_REQUIRE(mask || (!b1 && c0 == 0), "%s: null mask", who);
  B200SSL_REQUIRE(!tau || mask_out, "%s: null mask_out", who);
  bool vec = (w % 4 == 0) && aligned16(out1) && (!mask || aligned16(mask));
  if (c0) vec = vec && aligned16(a0) && aligned16(b0) && aligned16(out0);
  if (tau) vec = vec && aligned16(mask_out);
  const long long work = (long long)n * h * (vec ? w / 4 : w);
  long long blocks = (work + kMixThreads - 1) / kMixThreads;
  const long long cap = (long long)kNumSMs * 8 * 16;
  if (blocks > cap) blocks = cap;
  cudaStream_t s = (cudaStream_t)stream;
#define LAUNCH(V, F) \
  mix2_upsampled_kernel<V, F><<<(int)blocks, kMixThreads, 0, s>>>(a0, b0, out0, c0, a1, b1, out1, c1, h_in, w_in, mask, n, \
                                                                  h, w, tau, mask_out)
  prof_begin(b1 ? (tau ? "mix2_upsampled_threshold" : "mix2_upsampled") : "upsample_bilinear", s);
  if (tau) {
    if (vec) LAUNCH(4, true); else LAUNCH(1, true);
  } else {
    if (vec) LAUNCH(4, false); else LAUNCH(1, false);
  }
#undef LAUNCH
  return check_launch(who);
}

extern "C" int b200ssl_mix2_upsampled(const float* a0, const float* b0, float* out0, int c0, const float* a1_lo,
                                      const float* b1_lo, float* out1, int c1, int h_in, int w_in,
                                      const float* mask, const float* tau, float* mask_out, int64_t n, int h, int w,
                                      b200ssl_stream_t stream) {
  using namespace b200ssl;
  B200SSL_REQUIRE(c1 == 0 || b1_lo, "mix2_upsampled: null second low-resolution tensor");
  return mix2_upsampled_launch(a0, b0, out0, c0, a1_lo, b1_lo, out1, c1, h_in, w_in, mask, tau, mask_out, n, h, w, stream,
                               "mix2_upsampled");
}

extern "C" int b200ssl_upsample_bilinear(const float* in, int64_t planes, int h_in, int w_in, float* out, int h, int w,
                                         b200ssl_stream_t stream) {
  return mix2_upsampled_launch(nullptr, nullptr, nullptr, 0, in, nullptr, out, 1, h_in, w_in, nullptr, nullptr, nullptr,
                               planes, h, w, stream, "upsample_bilinear");
}

static int mix2_launch(const float* a0, const float* b0, float* out0, int c0, const float* a1, const float* b1,
                       float* out1, int c1, const float* mask, int mask_channels, int64_t n, int64_t hw,
                       const float* tau, float* mask_out, b200ssl_stream_t stream, const char* who) {
  using namespace b200ssl;
  B200SSL_REQUIRE(n >= 0 && hw >= 0 && c0 >= 0 && c1 >= 0, "%s: negative extent", who);
  if (n == 0 || hw == 0 || (c0 == 0 && c1 == 0 && !tau)) return 0;
  B200SSL_REQUIRE(mask != nullptr, "%s: null mask", who);
  B200SSL_REQUIRE(c0 == 0 || (a0 && b0 && out0), "%s: null tensor 0", who);
  B200SSL_REQUIRE(c1 == 0 || (a1 && b1 && out1), "%s: null tensor 1", who);
  const bool chan_mask = (mask_channels != 1);
  B200SSL_REQUIRE(!chan_mask || (mask_channels == c0 && c1 == 0 && !tau),
                  "%s: per-channel mask needs mask_channels == c0 and no second tensor", who);
  bool vec = (hw % 4 == 0) && aligned16(mask);
  if (c0) vec = vec && aligned16(a0) && aligned16(b0) && aligned16(out0);
  if (c1) vec = vec && aligned16(a1) && aligned16(b1) && aligned16(out1);
  if (tau) vec = vec && aligned16(mask_out);
  const long long work = (long long)n * (vec ? hw / 4 : hw);
  long long blocks = (work + kMixThreads - 1) / kMixThreads;
  const long long cap = (long long)kNumSMs * 8 * 16;  // grid-stride beyond 16 waves of 8 CTAs/SM
  if (blocks > cap) blocks = cap;
  cudaStream_t s = (cudaStream_t)stream;
#define LAUNCH(V, CM, F) \
  mix2_kernel<V, CM, F><<<(int)blocks, kMixThreads, 0, s>>>(a0, b0, out0, c0, a1, b1, out1, c1, mask, n, hw, tau, mask_out)
  prof_begin(tau ? "mix2_threshold" : "mix2", s);
  if (tau) {
    if (vec) LAUNCH(4, false, true); else LAUNCH(1, false, true);
  } else if (vec) {
    if (chan_mask) LAUNCH(4, true, false); else LAUNCH(4, false, false);
  } else {
    if (chan_mask) LAUNCH(1, true, false); else LAUNCH(1, false, false);
  }
#undef LAUNCH
  return check_launch(who);
}

extern "C" int b200ssl_mix2(const float* a0, const float* b0, float* out0, int c0, const float* a1,
                            const float* b1, float* out1, int c1, const float* mask,
                            int mask_channels, int64_t n, int64_t hw, b200ssl_stream_t stream) {
  return mix2_launch(a0, b0, out0, c0, a1, b1, out1, c1, mask, mask_channels, n, hw, nullptr, nullptr, stream, "mix2");
}

extern "C" int b200ssl_mix2_field(const float* a0, const float* b0, float* out0, int c0, const float* a1,
                                  const float* b1, float* out1, int c1, const float* field, const float* tau,
                                  float* mask_out, int64_t n, int64_t hw, b200ssl_stream_t stream) {
  using namespace b200ssl;
  B200SSL_REQUIRE(tau != nullptr && mask_out != nullptr, "mix2_field: null tau / mask_out");
  return mix2_launch(a0, b0, out0, c0, a1, b1, out1, c1, field, 1, n, hw, tau, mask_out, stream, "mix2_field");
}
