// Confidence-masked consistency loss of the semi-supervised branch, forward and backward
// ("next" row N1 of SURVEY 8f).  Reference semantics: train.py:98-107 (inline code):
//     t = sigmoid(mixed_ema_pred);  s = sigmoid(mixed_student_pred)
//     conf = (t.max(dim=1).values > confidence_threshold).to(t)                      [N,H,W]
//     loss = (pow(s - t, 2).sum(dim=1) * conf).sum() / conf.sum();  conf_mean = conf.mean()
// and autograd's backward w.r.t. the student logits:
//     d loss / d x[n,c,i] = go * 2 (s - t) * conf[n,i] / conf.sum() * s (1 - s)
// The reference runs ~12 elementwise/reduction kernels with 6 full-size temporaries; here the forward
// is ONE pass over both logit tensors (8 B/element) and the backward one more (12 B/element).
//
// Roofline: HBM-bound.  Sums are accumulated in fp64 and reduced in a fixed order (deterministic).
#include "bilinear.cuh"
#include "common.cuh"

namespace b200ssl {

constexpr int kConsThreads = 256;

__device__ __forceinline__ float sigmoidf_rn(float x) {
  // 1 / (1 + exp(-x)) with IEEE division; expf is the accurate (non fast-math) version
  return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
}

// grid = (blocks, n): each thread walks quads of pixels of image n, all C channels per quad
__global__ void __launch_bounds__(kConsThreads)
consistency_partial_kernel(const float* __restrict__ student, const float* __restrict__ teacher, int C,
                           long long hw, float thr, double* __restrict__ partials, bool vec) {
  const int n = blockIdx.y;
  const float* __restrict__ sp = student + (long long)n * C * hw;
  const float* __restrict__ tp = teacher + (long long)n * C * hw;
  double s_conf = 0.0, s_loss = 0.0;
  const long long quads = (hw + 3) / 4;
  for (long long q = (long long)blockIdx.x * kConsThreads + threadIdx.x; q < quads;
       q += (long long)gridDim.x * kConsThreads) {
    const long long i0 = q * 4;
    float tmax[4] = {-1.f, -1.f, -1.f, -1.f}, sq[4] = {0.f, 0.f, 0.f, 0.f};
    const bool full = vec && (i0 + 4 <= hw);
    for (int c = 0; c < C; ++c) {
      float sv[4] = {0.f, 0.f, 0.f, 0.f}, tv[4] = {0.f, 0.f, 0.f, 0.f};
      if (full) {
        const float4 a = ld_stream_f4(sp + (long long)c * hw + i0);
        const float4 b = ld_stream_f4(tp + (long long)c * hw + i0);
        sv[0] = a.x; sv[1] = a.y; sv[2] = a.z; sv[3] = a.w;
        tv[0] = b.x; tv[1] = b.y; tv[2] = b.z; tv[3] = b.w;
      } else {
        for (int e = 0; e < 4; ++e)
          if (i0 + e < hw) { sv[e] = sp[(long long)c * hw + i0 + e]; tv[e] = tp[(long long)c * hw + i0 + e]; }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float s = sigmoidf_rn(sv[e]), t = sigmoidf_rn(tv[e]);
        const float d = __fsub_rn(s, t);
        sq[e] = __fadd_rn(sq[e], __fmul_rn(d, d));          // pow(.,2).sum(dim=1): channels in order
        tmax[e] = fmaxf(tmax[e], t);
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (i0 + e < hw && tmax[e] > thr) {
        s_conf += 1.0;
        s_loss += (double)sq[e];
      }
    }
  }
  __shared__ double red[2][kConsThreads / 32];
  s_conf = warp_sum(s_conf);
  s_loss = warp_sum(s_loss);
  if (lane_id() == 0) { red[0][threadIdx.x >> 5] = s_conf; red[1][threadIdx.x >> 5] = s_loss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < kConsThreads / 32; ++w) { a += red[0][w]; b += red[1][w]; }
    const long long blk = (long long)n * gridDim.x + blockIdx.x;
    partials[2 * blk + 0] = a;
    partials[2 * blk + 1] = b;
  }
}

// stats[0] = loss, stats[1] = conf.sum(), stats[2] = conf.mean()
__global__ void consistency_final_kernel(const double* __restrict__ partials, int n_partials, double n_pixels,
                                         float* __restrict__ stats) {
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < n_partials; i += 32) { a += partials[2 * i]; b += partials[2 * i + 1]; }
  a = warp_sum(a);
  b = warp_sum(b);
  if (threadIdx.x == 0) {
    const float conf_sum = (float)a;
    stats[0] = __fdiv_rn((float)b, conf_sum);      // 0/0 = NaN when no pixel is confident, as the reference
    stats[1] = conf_sum;
    stats[2] = (float)(a / n_pixels);
  }
}

__global__ void __launch_bounds__(kConsThreads)
consistency_grad_kernel(const float* __restrict__ student, const float* __restrict__ teacher, int C, long long hw,
                        float thr, const float* __restrict__ stats, const float* __restrict__ grad_out,
                        float* __restrict__ grad, bool vec) {
  const int n = blockIdx.y;
  const float* __restrict__ sp = student + (long long)n * C * hw;
  const float* __restrict__ tp = teacher + (long long)n * C * hw;
  float* __restrict__ gp = grad + (long long)n * C * hw;
  // autograd: DivBackward -> go / conf_sum ; MulBackward (* conf) ; PowBackward 2 * d ; SigmoidBackward s (1 - s)
  const float scale = __fdiv_rn(grad_out[0], stats[1]);
  const long long quads = (hw + 3) / 4;
  for (long long q = (long long)blockIdx.x * kConsThreads + threadIdx.x; q < quads;
       q += (long long)gridDim.x * kConsThreads) {
    const long long i0 = q * 4;
    const bool full = vec && (i0 + 4 <= hw);
    // pass 1 over the channels: confidence of the 4 pixels (teacher only)
    // (raw logits here, so the sentinel is -inf: the forward kernel's -1 is only safe for sigmoid values)
    float tmax[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    for (int c = 0; c < C; ++c) {
      if (full) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(tp + (long long)c * hw + i0));
        tmax[0] = fmaxf(tmax[0], b.x); tmax[1] = fmaxf(tmax[1], b.y);
        tmax[2] = fmaxf(tmax[2], b.z); tmax[3] = fmaxf(tmax[3], b.w);
      } else {
        for (int e = 0; e < 4; ++e)
          if (i0 + e < hw) tmax[e] = fmaxf(tmax[e], tp[(long long)c * hw + i0 + e]);
      }
    }
    float conf[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) conf[e] = sigmoidf_rn(tmax[e]) > thr ? 1.0f : 0.0f;  // sigmoid is monotone
    // pass 2: gradient per channel (teacher planes come back from L1/L2)
    for (int c = 0; c < C; ++c) {
      float sv[4] = {0.f, 0.f, 0.f, 0.f}, tv[4] = {0.f, 0.f, 0.f, 0.f}, g[4];
      if (full) {
        const float4 a = ld_stream_f4(sp + (long long)c * hw + i0);
        const float4 b = __ldg(reinterpret_cast<const float4*>(tp + (long long)c * hw + i0));
        sv[0] = a.x; sv[1] = a.y; sv[2] = a.z; sv[3] = a.w;
        tv[0] = b.x; tv[1] = b.y; tv[2] = b.z; tv[3] = b.w;
      } else {
        for (int e = 0; e < 4; ++e)
          if (i0 + e < hw) { sv[e] = sp[(long long)c * hw + i0 + e]; tv[e] = tp[(long long)c * hw + i0 + e]; }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float s = sigmoidf_rn(sv[e]), t = sigmoidf_rn(tv[e]);
        const float d = __fsub_rn(s, t);
        const float up = __fmul_rn(__fmul_rn(scale, conf[e]), __fmul_rn(2.0f, d));
        g[e] = __fmul_rn(up, __fmul_rn(s, __fsub_rn(1.0f, s)));
      }
      if (full) {
        st_stream_f4(gp + (long long)c * hw + i0, make_float4(g[0], g[1], g[2], g[3]));
      } else {
        for (int e = 0; e < 4; ++e)
          if (i0 + e < hw) gp[(long long)c * hw + i0 + e] = g[e];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// The same loss with the TEACHER formed on the fly (SURVEY 8f N1 "fusing mix(pred) + sigmoid + confidence saves a
// full [N,C,H,W] round trip", N2): train.py:69-82 up-samples the two teacher predictions, mixes them with the CowMix
// mask into mixed_ema_pred, and train.py:98-107 is that tensor's only consumer.  Here
//     mixed_ema_pred[n,c,y,x] = RN(RN(A*m) + RN(B*RN(1-m))),  A / B = ema_pred_a / _b at (y,x)
// is evaluated in registers -- A and B read at full resolution (LOWRES = false) or bilinearly interpolated from the
// network's low-resolution logits with ATen's arithmetic (bilinear.cuh) -- so mixed_ema_pred is never written or
// read: 4C + 4 + 8C/s^2 bytes per pixel instead of 12C + 4 + 8C/s^2 for mix + loss.  The values are bit-identical
// to b200ssl_mix2(_upsampled) followed by the kernels above (tests/test_gpu_parity.py).
// One thread: VEC consecutive pixels of one row; grid = (blocks, n).
// ------------------------------------------------------------------------------------------
template <int VEC>
struct TeacherTaps {
  AxisTap ty;
  AxisTap tx[VEC];
};

template <int VEC, bool LOWRES>
__device__ __forceinline__ void mixed_teacher(const float* __restrict__ ap, const float* __restrict__ bp, int c,
                                              long long t_plane, int tw, long long off, const TeacherTaps<VEC>& tp,
                                              const float (&m)[VEC], const float (&om)[VEC], float (&t)[VEC]) {
  if (LOWRES) {
    const float* a0 = ap + (long long)c * t_plane + (long long)tp.ty.i0 * tw;
    const float* a1 = ap + (long long)c * t_plane + (long long)tp.ty.i1 * tw;
    const float* b0 = bp + (long long)c * t_plane + (long long)tp.ty.i0 * tw;
    const float* b1 = bp + (long long)c * t_plane + (long long)tp.ty.i1 * tw;
#pragma unroll
    for (int e = 0; e < VEC; ++e)
      t[e] = __fadd_rn(__fmul_rn(bilerp(a0, a1, tp.tx[e], tp.ty.w0, tp.ty.w1), m[e]),
                       __fmul_rn(bilerp(b0, b1, tp.tx[e], tp.ty.w0, tp.ty.w1), om[e]));
  } else if (VEC == 4) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(ap + (long long)c * t_plane + off));
    const float4 b = __ldg(reinterpret_cast<const float4*>(bp + (long long)c * t_plane + off));
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int e = 0; e < VEC; ++e) t[e] = __fadd_rn(__fmul_rn(av[e], m[e]), __fmul_rn(bv[e], om[e]));
  } else {
    t[0] = __fadd_rn(__fmul_rn(__ldg(ap + (long long)c * t_plane + off), m[0]),
                     __fmul_rn(__ldg(bp + (long long)c * t_plane + off), om[0]));
  }
}

template <int VEC>
__device__ __forceinline__ void load_row_vec(const float* __restrict__ p, float (&v)[VEC]) {
  if (VEC == 4) {
    const float4 a = ld_stream_f4(p);
    v[0] = a.x; v[VEC > 1 ? 1 : 0] = a.y; v[VEC > 2 ? 2 : 0] = a.z; v[VEC > 3 ? 3 : 0] = a.w;
  } else {
    v[0] = ld_stream_f1(p);
  }
}

template <int VEC, bool LOWRES>
__global__ void __launch_bounds__(kConsThreads, 3)
consistency_mixed_partial_kernel(const float* __restrict__ student, const float* __restrict__ ta,
                                 const float* __restrict__ tb, const float* __restrict__ mask, int C, int h, int w,
                                 int th, int tw, float thr, double* __restrict__ partials,
                                 unsigned char* __restrict__ conf_out) {
  const int n = blockIdx.y;
  const long long hw = (long long)h * w, t_plane = (long long)th * tw;
  const float* __restrict__ sp = student + (long long)n * C * hw;
  const float* __restrict__ ap = ta + (long long)n * C * t_plane;
  const float* __restrict__ bp = tb + (long long)n * C * t_plane;
  const float* __restrict__ mp = mask + (long long)n * hw;
  const float sy = (float)th / (float)h, sx = (float)tw / (float)w;
  const int per_row = w / VEC;
  const long long total = (long long)h * per_row;
  double s_conf = 0.0, s_loss = 0.0;
  for (long long q = (long long)blockIdx.x * kConsThreads + threadIdx.x; q < total;
       q += (long long)gridDim.x * kConsThreads) {
    const int y = (int)(q / per_row);
    const int x = (int)(q - (long long)y * per_row) * VEC;
    const long long off = (long long)y * w + x;
    float m[VEC], om[VEC];
    load_row_vec<VEC>(mp + off, m);
#pragma unroll
    for (int e = 0; e < VEC; ++e) om[e] = __fsub_rn(1.0f, m[e]);
    TeacherTaps<VEC> taps;
    if (LOWRES) {
      taps.ty = axis_tap(y, th, sy);
#pragma unroll
      for (int e = 0; e < VEC; ++e) taps.tx[e] = axis_tap(x + e, tw, sx);
    }
    float tmax[VEC], sq[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) { tmax[e] = -1.f; sq[e] = 0.f; }
    for (int c = 0; c < C; ++c) {
      float sv[VEC], tv[VEC];
      load_row_vec<VEC>(sp + (long long)c * hw + off, sv);
      mixed_teacher<VEC, LOWRES>(ap, bp, c, t_plane, tw, off, taps, m, om, tv);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const float s = sigmoidf_rn(sv[e]), t = sigmoidf_rn(tv[e]);
        const float d = __fsub_rn(s, t);
        sq[e] = __fadd_rn(sq[e], __fmul_rn(d, d));
        tmax[e] = fmaxf(tmax[e], t);
      }
    }
    unsigned char cf[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      cf[e] = tmax[e] > thr ? 1 : 0;
      if (cf[e]) {
        s_conf += 1.0;
        s_loss += (double)sq[e];
      }
    }
    if (conf_out) {   // the confidence decision, kept for the backward pass (one byte per pixel)
      if (VEC == 4) *reinterpret_cast<uchar4*>(conf_out + (long long)n * hw + off) = make_uchar4(cf[0], cf[VEC > 1 ? 1 : 0], cf[VEC > 2 ? 2 : 0], cf[VEC > 3 ? 3 : 0]);
      else conf_out[(long long)n * hw + off] = cf[0];
    }
  }
  __shared__ double red[2][kConsThreads / 32];
  s_conf = warp_sum(s_conf);
  s_loss = warp_sum(s_loss);
  if (lane_id() == 0) { red[0][threadIdx.x >> 5] = s_conf; red[1][threadIdx.x >> 5] = s_loss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int wv = 0; wv < kConsThreads / 32; ++wv) { a += red[0][wv]; b += red[1][wv]; }
    const long long blk = (long long)n * gridDim.x + blockIdx.x;
    partials[2 * blk + 0] = a;
    partials[2 * blk + 1] = b;
  }
}

template <int VEC, bool LOWRES>
__global__ void __launch_bounds__(kConsThreads, 3)
consistency_mixed_grad_kernel(const float* __restrict__ student, const float* __restrict__ ta,
                              const float* __restrict__ tb, const float* __restrict__ mask, int C, int h, int w,
                              int th, int tw, float thr, const float* __restrict__ stats,
                              const float* __restrict__ grad_out, float* __restrict__ grad,
                              const unsigned char* __restrict__ conf_in) {
  const int n = blockIdx.y;
  const long long hw = (long long)h * w, t_plane = (long long)th * tw;
  const float* __restrict__ sp = student + (long long)n * C * hw;
  const float* __restrict__ ap = ta + (long long)n * C * t_plane;
  const float* __restrict__ bp = tb + (long long)n * C * t_plane;
  const float* __restrict__ mp = mask + (long long)n * hw;
  float* __restrict__ gp = grad + (long long)n * C * hw;
  const float sy = (float)th / (float)h, sx = (float)tw / (float)w;
  const float scale = __fdiv_rn(grad_out[0], stats[1]);
  const int per_row = w / VEC;
  const long long total = (long long)h * per_row;
  for (long long q = (long long)blockIdx.x * kConsThreads + threadIdx.x; q < total;
       q += (long long)gridDim.x * kConsThreads) {
    const int y = (int)(q / per_row);
    const int x = (int)(q - (long long)y * per_row) * VEC;
    const long long off = (long long)y * w + x;
    float m[VEC], om[VEC];
    load_row_vec<VEC>(mp + off, m);
#pragma unroll
    for (int e = 0; e < VEC; ++e) om[e] = __fsub_rn(1.0f, m[e]);
    TeacherTaps<VEC> taps;
    if (LOWRES) {
      taps.ty = axis_tap(y, th, sy);
#pragma unroll
      for (int e = 0; e < VEC; ++e) taps.tx[e] = axis_tap(x + e, tw, sx);
    }
    // the confidence decision: the forward pass's byte per pixel, or (pass 1) from the largest mixed teacher logit
    // (sigmoid is monotone); then (pass 2) the gradient
    float conf[VEC];
    if (conf_in) {
      if (VEC == 4) {
        const uchar4 cb = __ldg(reinterpret_cast<const uchar4*>(conf_in + (long long)n * hw + off));
        conf[0] = cb.x ? 1.0f : 0.0f; conf[VEC > 1 ? 1 : 0] = cb.y ? 1.0f : 0.0f;
        conf[VEC > 2 ? 2 : 0] = cb.z ? 1.0f : 0.0f; conf[VEC > 3 ? 3 : 0] = cb.w ? 1.0f : 0.0f;
      } else {
        conf[0] = conf_in[(long long)n * hw + off] ? 1.0f : 0.0f;
      }
    } else {
      float tmax[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) tmax[e] = -INFINITY;
      for (int c = 0; c < C; ++c) {
        float tv[VEC];
        mixed_teacher<VEC, LOWRES>(ap, bp, c, t_plane, tw, off, taps, m, om, tv);
#pragma unroll
        for (int e = 0; e < VEC; ++e) tmax[e] = fmaxf(tmax[e], tv[e]);
      }
#pragma unroll
      for (int e = 0; e < VEC; ++e) conf[e] = sigmoidf_rn(tmax[e]) > thr ? 1.0f : 0.0f;
    }
    for (int c = 0; c < C; ++c) {
      float sv[VEC], tv[VEC], g[VEC];
      load_row_vec<VEC>(sp + (long long)c * hw + off, sv);
      mixed_teacher<VEC, LOWRES>(ap, bp, c, t_plane, tw, off, taps, m, om, tv);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const float s = sigmoidf_rn(sv[e]), t = sigmoidf_rn(tv[e]);
        const float d = __fsub_rn(s, t);
        const float up = __fmul_rn(__fmul_rn(scale, conf[e]), __fmul_rn(2.0f, d));
        g[e] = __fmul_rn(up, __fmul_rn(s, __fsub_rn(1.0f, s)));
      }
      if (VEC == 4) st_stream_f4(gp + (long long)c * hw + off, make_float4(g[0], g[VEC > 1 ? 1 : 0], g[VEC > 2 ? 2 : 0], g[VEC > 3 ? 3 : 0]));
      else gp[(long long)c * hw + off] = g[0];
    }
  }
}

static int cons_blocks(int n, long long hw) {
  long long bx = ((hw + 3) / 4 + kConsThreads - 1) / kConsThreads;
  long long cap = (long long)kNumSMs * 8 / (n > 0 ? n : 1);
  if (cap < 1) cap = 1;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  return (int)bx;
}

}  // namespace b200ssl

extern "C" {

size_t b200ssl_consistency_workspace_bytes(int n, int64_t hw) {
  if (n <= 0 || hw <= 0) return 0;
  return (size_t)n * b200ssl::cons_blocks(n, hw) * 2 * sizeof(double);
}

int b200ssl_consistency_forward(const float* student, const float* teacher, int n, int c, int64_t hw,
                                float threshold, float* stats_out, void* workspace, size_t workspace_bytes,
                                b200ssl_stream_t stream) {
  using namespace b200ssl;
  B200SSL_REQUIRE(n >= 1 && c >= 1 && hw >= 1, "consistency_forward: bad extents");
  B200SSL_REQUIRE(n <= 65535, "consistency_forward: too many images");
  B200SSL_REQUIRE(student && teacher && stats_out, "consistency_forward: null argument");
  const int bx = cons_blocks(n, hw);
  const size_t need = (size_t)n * bx * 2 * sizeof(double);
  if (!workspace || workspace_bytes < need) {
    set_error("consistency_forward: workspace too small (%zu < %zu)", workspace_bytes, need);
    return B200SSL_EWORKSPACE;
  }
  const bool vec = aligned16(student) && aligned16(teacher) && (hw % 4 == 0);
  cudaStream_t s = (cudaStream_t)stream;
  prof_begin("consistency_partial", s);
  consistency_partial_kernel<<<dim3((unsigned)bx, (unsigned)n), kConsThreads, 0, s>>>(
      student, teacher, c, hw, threshold, static_cast<double*>(workspace), vec);
  int rc = check_launch("consistency partial");
  if (rc) return rc;
  prof_begin("consistency_final", s);
  consistency_final_kernel<<<1, 32, 0, s>>>(static_cast<const double*>(workspace), n * bx, (double)n * (double)hw,
                                            stats_out);
  return check_launch("consistency final");
}

int b200ssl_consistency_backward(const float* student, const float* teacher, int n, int c, int64_t hw,
                                 float threshold, const float* stats, const float* grad_out,
                                 float* grad_student, b200ssl_stream_t stream) {
  using namespace b200ssl;
  B200SSL_REQUIRE(n >= 1 && c >= 1 && hw >= 1, "consistency_backward: bad extents");
  B200SSL_REQUIRE(n <= 65535, "consistency_backward: too many images");
  B200SSL_REQUIRE(student && teacher && stats && grad_out && grad_student, "consistency_backward: null argument");
  const bool vec = aligned16(student) && aligned16(teacher) && aligned16(grad_student) && (hw % 4 == 0);
  cudaStream_t s = (cudaStream_t)stream;
  prof_begin("consistency_grad", s);
  consistency_grad_kernel<<<dim3((unsigned)cons_blocks(n, hw), (unsigned)n), kConsThreads, 0, s>>>(
      student, teacher, c, hw, threshold, stats, grad_out, grad_student, vec);
  return check_launch("consistency grad");
}


size_t b200ssl_consistency_mixed_workspace_bytes(int n, int h, int w) {
  return b200ssl_consistency_workspace_bytes(n, (int64_t)h * w);
}

namespace {
struct MixedArgs {
  bool vec, lowres;
  int bx;
};
int mixed_args(const char* who, const float* student, const float* ta, const float* tb, const float* mask, int n, int c,
               int h, int w, int th, int tw, const float* extra, MixedArgs* out) {
  using namespace b200ssl;
  B200SSL_REQUIRE(n >= 1 && c >= 1 && h >= 1 && w >= 1 && th >= 1 && tw >= 1, "%s: bad extents", who);
  B200SSL_REQUIRE(n <= 65535, "%s: too many images", who);
  B200SSL_REQUIRE(th <= h && tw <= w, "%s: the teacher predictions are larger than the student's (%dx%d > %dx%d)", who, th, tw, h, w);
  B200SSL_REQUIRE(student && ta && tb && mask, "%s: null argument", who);
  out->lowres = !(th == h && tw == w);
  out->vec = (w % 4 == 0) && aligned16(student) && aligned16(mask) && (!extra || aligned16(extra)) &&
             (out->lowres || (aligned16(ta) && aligned16(tb)));
  out->bx = cons_blocks(n, (long long)h * w);
  return 0;
}
}  // namespace

int b200ssl_consistency_mixed_forward(const float* student, const float* teacher_a, const float* teacher_b,
                                      const float* mask, int n, int c, int h, int w, int th, int tw, float threshold,
                                      float* stats_out, unsigned char* conf_out, void* workspace,
                                      size_t workspace_bytes, b200ssl_stream_t stream) {
  using namespace b200ssl;
  MixedArgs a;
  int rc = mixed_args("consistency_mixed_forward", student, teacher_a, teacher_b, mask, n, c, h, w, th, tw, nullptr, &a);
  if (conf_out && (reinterpret_cast<uintptr_t>(conf_out) & 3u)) a.vec = false;
  if (rc) return rc;
  B200SSL_REQUIRE(stats_out, "consistency_mixed_forward: null argument");
  const size_t need = (size_t)n * a.bx * 2 * sizeof(double);
  if (!workspace || workspace_bytes < need) {
    set_error("consistency_mixed_forward: workspace too small (%zu < %zu)", workspace_bytes, need);
    return B200SSL_EWORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const dim3 grid((unsigned)a.bx, (unsigned)n);
  double* part = static_cast<double*>(workspace);
  prof_begin("consistency_mixed_partial", s);
#define LAUNCH(V, L) \
  consistency_mixed_partial_kernel<V, L><<<grid, kConsThreads, 0, s>>>(student, teacher_a, teacher_b, mask, c, h, w, th, tw, threshold, part, conf_out)
  if (a.vec) { if (a.lowres) LAUNCH(4, true); else LAUNCH(4, false); }
  else       { if (a.lowres) LAUNCH(1, true); else LAUNCH(1, false); }
#undef LAUNCH
  rc = check_launch("consistency mixed partial");
  if (rc) return rc;
  prof_begin("consistency_final", s);
  consistency_final_kernel<<<1, 32, 0, s>>>(part, n * a.bx, (double)n * (double)h * (double)w, stats_out);
  return check_launch("consistency final");
}

int b200ssl_consistency_mixed_backward(const float* student, const float* teacher_a, const float* teacher_b,
                                       const float* mask, int n, int c, int h, int w, int th, int tw, float threshold,
                                       const float* stats, const unsigned char* conf, const float* grad_out,
                                       float* grad_student, b200ssl_stream_t stream) {
  using namespace b200ssl;
  MixedArgs a;
  int rc = mixed_args("consistency_mixed_backward", student, teacher_a, teacher_b, mask, n, c, h, w, th, tw, grad_student, &a);
  if (conf && (reinterpret_cast<uintptr_t>(conf) & 3u)) a.vec = false;
  if (rc) return rc;
  B200SSL_REQUIRE(stats && grad_out && grad_student, "consistency_mixed_backward: null argument");
  cudaStream_t s = (cudaStream_t)stream;
  const dim3 grid((unsigned)a.bx, (unsigned)n);
  prof_begin("consistency_mixed_grad", s);
#define LAUNCH(V, L) \
  consistency_mixed_grad_kernel<V, L><<<grid, kConsThreads, 0, s>>>(student, teacher_a, teacher_b, mask, c, h, w, th, tw, threshold, stats, grad_out, grad_student, conf)
  if (a.vec) { if (a.lowres) LAUNCH(4, true); else LAUNCH(4, false); }
  else       { if (a.lowres) LAUNCH(1, true); else LAUNCH(1, false); }
#undef LAUNCH
  return check_launch("consistency mixed grad");
}

}  // extern "C"
