// Row N3: the soft-max either side of lovasz_softmax, without materialising probabilities.
//   lovasz.py:155-160's contract is lovasz_softmax(F.softmax(logits, 1), labels); autograd then runs
//   _softmax_backward_data on the gradient.  Here:
//   softmax_stats_kernel     one pass over the logits -> per-pixel (max, sum exp(x - max)), 8 B/pixel;
//                            the key-build forms p = exp(x - max) / sum in registers (lovasz.cu)
//   softmax_backward_kernel  dL/dz_c = (g_c - sum_j g_j p_j) * p_c with p recomputed from the logits and
//                            the statistics, in place on the gradient buffer
// Arithmetic follows ATen's soft-max: max over channels, exp of the difference, channel-order fp32 sum,
// IEEE division; backward (grad - dot) * output.  HBM-bound: 4C+8 B/pixel forward, 12C+8 backward.
#include "common.cuh"

namespace b200ssl {

constexpr int kSmThreads = 256;

template <int VEC>
__global__ void __launch_bounds__(kSmThreads)
softmax_stats_kernel(const float* __restrict__ x, int n, int C, long long hw, float* __restrict__ smax,
                     float* __restrict__ ssum) {
  const long long per_img = hw / VEC;
  const long long total = (long long)n * per_img;
  for (long long q = (long long)blockIdx.x * kSmThreads + threadIdx.x; q < total; q += (long long)gridDim.x * kSmThreads) {
    const long long img = q / per_img;
    const long long off = (q - img * per_img) * VEC;
    const float* base = x + img * C * hw + off;
    float m[VEC], s[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) { m[e] = -INFINITY; s[e] = 0.f; }
    for (int c = 0; c < C; ++c) {
      float v[VEC];
      if (VEC == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(base + (long long)c * hw));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      } else {
        v[0] = __ldg(base + (long long)c * hw);
      }
#pragma unroll
      for (int e = 0; e < VEC; ++e) m[e] = (v[e] > m[e] || v[e] != v[e]) ? v[e] : m[e];   // NaN propagates like torch.max
    }
    for (int c = 0; c < C; ++c) {   // second sweep hits L1/L2
      float v[VEC];
      if (VEC == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(base + (long long)c * hw));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      } else {
        v[0] = __ldg(base + (long long)c * hw);
      }
#pragma unroll
      for (int e = 0; e < VEC; ++e) s[e] = __fadd_rn(s[e], expf(__fsub_rn(v[e], m[e])));
    }
    if (VEC == 4) {
      *reinterpret_cast<float4*>(smax + img * hw + off) = make_float4(m[0], m[1], m[2], m[3]);
      *reinterpret_cast<float4*>(ssum + img * hw + off) = make_float4(s[0], s[1], s[2], s[3]);
    } else {
      smax[img * hw + off] = m[0];
      ssum[img * hw + off] = s[0];
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(kSmThreads)
softmax_backward_kernel(const float* __restrict__ x, const float* __restrict__ smax, const float* __restrict__ ssum,
                        float* grad, int n, int C, long long hw) {
  const long long per_img = hw / VEC;
  const long long total = (long long)n * per_img;
  for (long long q = (long long)blockIdx.x * kSmThreads + threadIdx.x; q < total; q += (long long)gridDim.x * kSmThreads) {
    const long long img = q / per_img;
    const long long off = (q - img * per_img) * VEC;
    const float* xb = x + img * C * hw + off;
    float* gb = grad + img * C * hw + off;
    float m[VEC], s[VEC], dot[VEC];
    if (VEC == 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(smax + img * hw + off));
      const float4 b = __ldg(reinterpret_cast<const float4*>(ssum + img * hw + off));
      m[0] = a.x; m[1] = a.y; m[2] = a.z; m[3] = a.w;
      s[0] = b.x; s[1] = b.y; s[2] = b.z; s[3] = b.w;
    } else {
      m[0] = __ldg(smax + img * hw + off);
      s[0] = __ldg(ssum + img * hw + off);
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) dot[e] = 0.f;
    for (int c = 0; c < C; ++c) {
      float v[VEC], g[VEC];
      if (VEC == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(xb + (long long)c * hw));
        const float4 u = *reinterpret_cast<const float4*>(gb + (long long)c * hw);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        g[0] = u.x; g[1] = u.y; g[2] = u.z; g[3] = u.w;
      } else {
        v[0] = __ldg(xb + (long long)c * hw);
        g[0] = gb[(long long)c * hw];
      }
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const float pr = __fdiv_rn(expf(__fsub_rn(v[e], m[e])), s[e]);
        dot[e] = __fadd_rn(dot[e], __fmul_rn(g[e], pr));
      }
    }
    for (int c = 0; c < C; ++c) {
      float v[VEC], g[VEC], o[VEC];
      if (VEC == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(xb + (long long)c * hw));
        const float4 u = *reinterpret_cast<const float4*>(gb + (long long)c * hw);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        g[0] = u.x; g[1] = u.y; g[2] = u.z; g[3] = u.w;
      } else {
        v[0] = __ldg(xb + (long long)c * hw);
        g[0] = gb[(long long)c * hw];
      }
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const float pr = __fdiv_rn(expf(__fsub_rn(v[e], m[e])), s[e]);
        o[e] = __fmul_rn(__fsub_rn(g[e], dot[e]), pr);
      }
      if (VEC == 4) *reinterpret_cast<float4*>(gb + (long long)c * hw) = make_float4(o[0], o[1], o[2], o[3]);
      else gb[(long long)c * hw] = o[0];
    }
  }
}

// ---- register-resident variants for C <= 32: every logit (and gradient) of a pixel is read from memory
//      exactly once (the generic kernels above sweep the channels twice, and the second sweep misses L2
//      once the planes in flight exceed it: 20C instead of 12C bytes per pixel in the backward pass).
//      One pixel per thread; a warp reads 128 contiguous bytes of every channel plane.
constexpr int kSmRegC = 32;

__global__ void __launch_bounds__(kSmThreads)
softmax_stats_reg_kernel(const float* __restrict__ x, int n, int C, long long hw, float* __restrict__ smax,
                         float* __restrict__ ssum) {
  const long long total = (long long)n * hw;
  for (long long q = (long long)blockIdx.x * kSmThreads + threadIdx.x; q < total; q += (long long)gridDim.x * kSmThreads) {
    const long long img = q / hw;
    const long long off = q - img * hw;
    const float* base = x + img * C * hw + off;
    float v[kSmRegC];
#pragma unroll
    for (int c = 0; c < kSmRegC; ++c) v[c] = (c < C) ? ld_stream_f1(base + (long long)c * hw) : -INFINITY;
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < kSmRegC; ++c)
      if (c < C) m = (v[c] > m || v[c] != v[c]) ? v[c] : m;
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < kSmRegC; ++c)
      if (c < C) s = __fadd_rn(s, expf(__fsub_rn(v[c], m)));
    smax[q] = m;
    ssum[q] = s;
  }
}

__global__ void __launch_bounds__(kSmThreads)
softmax_backward_reg_kernel(const float* __restrict__ x, const float* __restrict__ smax, const float* __restrict__ ssum,
                            float* grad, int n, int C, long long hw) {
  const long long total = (long long)n * hw;
  for (long long q = (long long)blockIdx.x * kSmThreads + threadIdx.x; q < total; q += (long long)gridDim.x * kSmThreads) {
    const long long img = q / hw;
    const long long off = q - img * hw;
    const float* xb = x + img * C * hw + off;
    float* gb = grad + img * C * hw + off;
    const float m = __ldg(smax + q), s = __ldg(ssum + q);
    float pr[kSmRegC], g[kSmRegC];
#pragma unroll
    for (int c = 0; c < kSmRegC; ++c) {
      pr[c] = (c < C) ? ld_stream_f1(xb + (long long)c * hw) : 0.f;
      g[c] = (c < C) ? gb[(long long)c * hw] : 0.f;
    }
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < kSmRegC; ++c)
      if (c < C) {
        pr[c] = __fdiv_rn(expf(__fsub_rn(pr[c], m)), s);
        dot = __fadd_rn(dot, __fmul_rn(g[c], pr[c]));
      }
#pragma unroll
    for (int c = 0; c < kSmRegC; ++c)
      if (c < C) gb[(long long)c * hw] = __fmul_rn(__fsub_rn(g[c], dot), pr[c]);
  }
}

// ---- round 2: F.softmax written out once + the backward from the stored probabilities.  With the exact tail
//      pruning and the one-byte label copy on the probability path, `soft-max -> lovasz_softmax -> soft-max
//      backward` through these two kernels (8C + 12C bytes per pixel) beats the never-materialising logits
//      front end above (0.58 vs 0.94 ms at 4x21x512x512); lovasz_softmax_with_logits takes this route by default.
//      p = exp(x - max) / sum exactly as softmax_stats + the logits key-build form it.
__global__ void __launch_bounds__(kSmThreads)
softmax_forward_reg_kernel(const float* __restrict__ x, int n, int C, long long hw, float* __restrict__ probas) {
  const long long total = (long long)n * hw;
  for (long long q = (long long)blockIdx.x * kSmThreads + threadIdx.x; q < total; q += (long long)gridDim.x * kSmThreads) {
    const long long img = q / hw;
    const long long off = q - img * hw;
    const float* base = x + img * C * hw + off;
    float* out = probas + img * C * hw + off;
    float v[kSmRegC];
#pragma unroll
    for (int c = 0; c < kSmRegC; ++c) v[c] = (c < C) ? ld_stream_f1(base + (long long)c * hw) : -INFINITY;
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < kSmRegC; ++c)
      if (c < C) m = (v[c] > m || v[c] != v[c]) ? v[c] : m;
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < kSmRegC; ++c)
      if (c < C) {
        v[c] = expf(__fsub_rn(v[c], m));
        s = __fadd_rn(s, v[c]);
      }
#pragma unroll
    for (int c = 0; c < kSmRegC; ++c)
      if (c < C) out[(long long)c * hw] = __fdiv_rn(v[c], s);
  }
}

// generic C: three sweeps over the channels of a pixel (max, sum, write); the planes of one pixel column stay in L1/L2
__global__ void __launch_bounds__(kSmThreads)
softmax_forward_kernel(const float* __restrict__ x, int n, int C, long long hw, float* __restrict__ probas) {
  const long long total = (long long)n * hw;
  for (long long q = (long long)blockIdx.x * kSmThreads + threadIdx.x; q < total; q += (long long)gridDim.x * kSmThreads) {
    const long long img = q / hw;
    const long long off = q - img * hw;
    const float* base = x + img * C * hw + off;
    float* out = probas + img * C * hw + off;
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) {
      const float v = __ldg(base + (long long)c * hw);
      m = (v > m || v != v) ? v : m;
    }
    float s = 0.f;
    for (int c = 0; c < C; ++c) s = __fadd_rn(s, expf(__fsub_rn(__ldg(base + (long long)c * hw), m)));
    for (int c = 0; c < C; ++c) out[(long long)c * hw] = __fdiv_rn(expf(__fsub_rn(__ldg(base + (long long)c * hw), m)), s);
  }
}

// grad <- (grad - sum_c grad_c p_c) * p, in place, p read from the stored probabilities
template <bool REG>
__global__ void __launch_bounds__(kSmThreads)
softmax_backward_probas_kernel(const float* __restrict__ probas, float* grad, int n, int C, long long hw) {
  const long long total = (long long)n * hw;
  for (long long q = (long long)blockIdx.x * kSmThreads + threadIdx.x; q < total; q += (long long)gridDim.x * kSmThreads) {
    const long long img = q / hw;
    const long long off = q - img * hw;
    const float* pb = probas + img * C * hw + off;
    float* gb = grad + img * C * hw + off;
    if (REG) {
      float pr[kSmRegC], g[kSmRegC];
#pragma unroll
      for (int c = 0; c < kSmRegC; ++c) {
        pr[c] = (c < C) ? ld_stream_f1(pb + (long long)c * hw) : 0.f;
        g[c] = (c < C) ? gb[(long long)c * hw] : 0.f;
      }
      float dot = 0.f;
#pragma unroll
      for (int c = 0; c < kSmRegC; ++c)
        if (c < C) dot = __fadd_rn(dot, __fmul_rn(g[c], pr[c]));
#pragma unroll
      for (int c = 0; c < kSmRegC; ++c)
        if (c < C) gb[(long long)c * hw] = __fmul_rn(__fsub_rn(g[c], dot), pr[c]);
    } else {
      float dot = 0.f;
      for (int c = 0; c < C; ++c) dot = __fadd_rn(dot, __fmul_rn(gb[(long long)c * hw], __ldg(pb + (long long)c * hw)));
      for (int c = 0; c < C; ++c)
        gb[(long long)c * hw] = __fmul_rn(__fsub_rn(gb[(long long)c * hw], dot), __ldg(pb + (long long)c * hw));
    }
  }
}

static int sm_grid(long long work) {
  long long blocks = (work + kSmThreads - 1) / kSmThreads;
  const long long cap = (long long)kNumSMs * 8 * 16;
  return (int)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

}  // namespace b200ssl

extern "C" {

int b200ssl_softmax_stats(const float* logits, int n, int c, int64_t hw, float* softmax_max, float* softmax_sum,
                          b200ssl_stream_t stream) {
  using namespace b200ssl;
  B200SSL_REQUIRE(n >= 0 && c >= 1 && hw >= 0, "softmax_stats: bad extents");
  if (n == 0 || hw == 0) return 0;
  B200SSL_REQUIRE(logits && softmax_max && softmax_sum, "softmax_stats: null argument");
  const bool vec = (hw % 4 == 0) && aligned16(logits) && aligned16(softmax_max) && aligned16(softmax_sum);
  cudaStream_t s = (cudaStream_t)stream;
  prof_begin("softmax_stats", s);
  if (c <= kSmRegC) softmax_stats_reg_kernel<<<sm_grid((long long)n * hw), kSmThreads, 0, s>>>(logits, n, c, hw, softmax_max, softmax_sum);
  else if (vec) softmax_stats_kernel<4><<<sm_grid((long long)n * hw / 4), kSmThreads, 0, s>>>(logits, n, c, hw, softmax_max, softmax_sum);
  else softmax_stats_kernel<1><<<sm_grid((long long)n * hw), kSmThreads, 0, s>>>(logits, n, c, hw, softmax_max, softmax_sum);
  return check_launch("softmax_stats");
}

int b200ssl_softmax_backward(const float* logits, const float* softmax_max, const float* softmax_sum, float* grad,
                             int n, int c, int64_t hw, b200ssl_stream_t stream) {
  using namespace b200ssl;
  B200SSL_REQUIRE(n >= 0 && c >= 1 && hw >= 0, "softmax_backward: bad extents");
  if (n == 0 || hw == 0) return 0;
  B200SSL_REQUIRE(logits && softmax_max && softmax_sum && grad, "softmax_backward: null argument");
  const bool vec = (hw % 4 == 0) && aligned16(logits) && aligned16(softmax_max) && aligned16(softmax_sum) && aligned16(grad);
  cudaStream_t s = (cudaStream_t)stream;
  prof_begin("softmax_backward", s);
  if (c <= kSmRegC) softmax_backward_reg_kernel<<<sm_grid((long long)n * hw), kSmThreads, 0, s>>>(logits, softmax_max, softmax_sum, grad, n, c, hw);
  else if (vec) softmax_backward_kernel<4><<<sm_grid((long long)n * hw / 4), kSmThreads, 0, s>>>(logits, softmax_max, softmax_sum, grad, n, c, hw);
  else softmax_backward_kernel<1><<<sm_grid((long long)n * hw), kSmThreads, 0, s>>>(logits, softmax_max, softmax_sum, grad, n, c, hw);
  return check_launch("softmax_backward");
}

int b200ssl_softmax_forward(const float* logits, int n, int c, int64_t hw, float* probas, b200ssl_stream_t stream) {
  using namespace b200ssl;
  B200SSL_REQUIRE(n >= 0 && c >= 1 && hw >= 0, "softmax_forward: bad extents");
  if (n == 0 || hw == 0) return 0;
  B200SSL_REQUIRE(logits && probas, "softmax_forward: null argument");
  cudaStream_t s = (cudaStream_t)stream;
  prof_begin("softmax_forward", s);
  if (c <= kSmRegC) softmax_forward_reg_kernel<<<sm_grid((long long)n * hw), kSmThreads, 0, s>>>(logits, n, c, hw, probas);
  else softmax_forward_kernel<<<sm_grid((long long)n * hw), kSmThreads, 0, s>>>(logits, n, c, hw, probas);
  return check_launch("softmax_forward");
}

int b200ssl_softmax_backward_probas(const float* probas, float* grad, int n, int c, int64_t hw, b200ssl_stream_t stream) {
  using namespace b200ssl;
  B200SSL_REQUIRE(n >= 0 && c >= 1 && hw >= 0, "softmax_backward_probas: bad extents");
  if (n == 0 || hw == 0) return 0;
  B200SSL_REQUIRE(probas && grad, "softmax_backward_probas: null argument");
  cudaStream_t s = (cudaStream_t)stream;
  prof_begin("softmax_backward_probas", s);
  if (c <= kSmRegC) softmax_backward_probas_kernel<true><<<sm_grid((long long)n * hw), kSmThreads, 0, s>>>(probas, grad, n, c, hw);
  else softmax_backward_probas_kernel<false><<<sm_grid((long long)n * hw), kSmThreads, 0, s>>>(probas, grad, n, c, hw);
  return check_launch("softmax_backward_probas");
}

}  // extern "C"
