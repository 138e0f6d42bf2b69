// Row N4: multi-tensor gradient-norm clip + SGD step with the mean-teacher EMA fused into
// its epilogue -- the three statements that surround the loss path in the reference's step:
//   train.py:122  torch.nn.utils.clip_grad_norm_(model.module.parameters(), gradient_clip_value)
//   train.py:123  optimizer.step()        (torch.optim.SGD, momentum 0.9, weight decay 5e-4:
//                                          configs/default_config.py:151-154)
//   train.py:124  optimizer.zero_grad()
//   train.py:130  mean_teacher.update_ema_variables(model, ema_model, alpha)   (mean_teacher.py:10-11)
// ATen runs them as ~7 element-wise passes per tensor (or per foreach group) over hundreds of tiny
// tensors: launch-bound, 72 B/parameter.  Here: one reduction launch + one 1-block launch for the
// norm, then ONE launch that reads p, g, momentum, teacher once and writes p, momentum, teacher
// (28 B/parameter, +4 with zero_grad) walking the same kind of chunk table as the EMA kernel.
//
// Arithmetic (bit-exact against torch's CPU and CUDA paths, probed; every `add(alpha=)` is one fma):
//   g  = RN(g * coef)                       clip_grad_norm_: grads.mul_(clip_coef_clamped)
//   g  = fma(p, wd, g)                      grad.add(param, alpha=weight_decay)      (skipped if wd == 0)
//   b  = first step ? g : fma(g, 1-damp, RN(b * mu))         buf.mul_(mu).add_(grad, alpha=1-dampening)
//   d  = nesterov ? fma(b, mu, g) : b
//   p  = fma(d, -lr, p)                     param.add_(d, alpha=-lr)
//   e  = fma(p, 1-alpha, RN(e * alpha))     mean_teacher.py:10-11 on the UPDATED parameter
//   coef = min(RN(RN(1 / RN(norm + 1e-6)) * max_norm), 1)     (python `max_norm / tensor` is reciprocal * scalar)
// The norm itself is a sum of squares accumulated in fp64 in a fixed order (deterministic; within
// 1e-6 relative of torch's fp32 norm-of-norms, which depends on its reduction order).
#include "common.cuh"

namespace b200ssl {

constexpr int kOptThreads = 256;
constexpr int kOptVec = B200SSL_EMA_CHUNK / (kOptThreads * 4);  // float4 per thread per chunk (4)

struct SgdConsts {
  float neg_lr, mu, one_minus_damp, wd, ema_a, ema_b;
  int nesterov, first_step, zero_grad, has_wd, has_mu, has_ema;
};

__device__ __forceinline__ void sgd_one(float& p, float g, float& b, float& e, float coef, const SgdConsts& k) {
  g = __fmul_rn(g, coef);
  if (k.has_wd) g = __fmaf_rn(p, k.wd, g);
  float d = g;
  if (k.has_mu) {
    b = k.first_step ? g : __fmaf_rn(g, k.one_minus_damp, __fmul_rn(b, k.mu));
    d = k.nesterov ? __fmaf_rn(b, k.mu, g) : b;
  }
  p = __fmaf_rn(d, k.neg_lr, p);
  if (k.has_ema) e = __fmaf_rn(p, k.ema_b, __fmul_rn(e, k.ema_a));
}

__device__ __forceinline__ float4 ld_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p)
               : "memory");
  return r;
}

__global__ void __launch_bounds__(kOptThreads, 3)
sgd_ema_multi_kernel(const b200ssl_sgd_chunk* __restrict__ table, long long n_entries,
                     const float* __restrict__ coef_dev, const __grid_constant__ SgdConsts k) {
  const float coef = coef_dev ? *coef_dev : 1.0f;
  for (long long c = blockIdx.x; c < n_entries; c += gridDim.x) {
    const b200ssl_sgd_chunk ent = table[c];
    float* __restrict__ p = ent.param;
    float* __restrict__ g = ent.grad;
    float* __restrict__ b = ent.momentum;
    float* __restrict__ e = ent.ema;
    const int count = ent.count;
    const bool has_b = k.has_mu && b != nullptr, has_e = k.has_ema && e != nullptr;
    SgdConsts kk = k;
    kk.has_mu = has_b;
    kk.has_ema = has_e;
    const uintptr_t bits = reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                           (has_b ? reinterpret_cast<uintptr_t>(b) : 0) | (has_e ? reinterpret_cast<uintptr_t>(e) : 0);
    if ((bits & 15u) == 0) {
      const int nvec = count >> 2;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float4 pv[kOptVec / 2], gv[kOptVec / 2], bv[kOptVec / 2], ev[kOptVec / 2];
#pragma unroll
        for (int j = 0; j < kOptVec / 2; ++j) {
          const int v = threadIdx.x + (half * (kOptVec / 2) + j) * kOptThreads;
          if (v < nvec) {
            pv[j] = ld_f4(p + 4 * v);
            gv[j] = ld_f4(g + 4 * v);
            bv[j] = (has_b && !kk.first_step) ? ld_f4(b + 4 * v) : make_float4(0.f, 0.f, 0.f, 0.f);
            ev[j] = has_e ? ld_f4(e + 4 * v) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int j = 0; j < kOptVec / 2; ++j) {
          const int v = threadIdx.x + (half * (kOptVec / 2) + j) * kOptThreads;
          if (v < nvec) {
            sgd_one(pv[j].x, gv[j].x, bv[j].x, ev[j].x, coef, kk);
            sgd_one(pv[j].y, gv[j].y, bv[j].y, ev[j].y, coef, kk);
            sgd_one(pv[j].z, gv[j].z, bv[j].z, ev[j].z, coef, kk);
            sgd_one(pv[j].w, gv[j].w, bv[j].w, ev[j].w, coef, kk);
            st_stream_f4(p + 4 * v, pv[j]);
            if (has_b) st_stream_f4(b + 4 * v, bv[j]);
            if (has_e) st_stream_f4(e + 4 * v, ev[j]);
            if (kk.zero_grad) st_stream_f4(g + 4 * v, make_float4(0.f, 0.f, 0.f, 0.f));
          }
        }
      }
      const int tail = (nvec << 2) + threadIdx.x;
      if (tail < count) {
        float pp = p[tail], bb = (has_b && !kk.first_step) ? b[tail] : 0.f, ee = has_e ? e[tail] : 0.f;
        sgd_one(pp, g[tail], bb, ee, coef, kk);
        p[tail] = pp;
        if (has_b) b[tail] = bb;
        if (has_e) e[tail] = ee;
        if (kk.zero_grad) g[tail] = 0.f;
      }
    } else {
      for (int i = threadIdx.x; i < count; i += kOptThreads) {
        float pp = p[i], bb = (has_b && !kk.first_step) ? b[i] : 0.f, ee = has_e ? e[i] : 0.f;
        sgd_one(pp, g[i], bb, ee, coef, kk);
        p[i] = pp;
        if (has_b) b[i] = bb;
        if (has_e) e[i] = ee;
        if (kk.zero_grad) g[i] = 0.f;
      }
    }
  }
}

// fixed-order block reduction of doubles (warp shuffle tree, then warp 0 over the warp sums)
__device__ __forceinline__ double block_sum_f64(double v, double* smem) {
  v = warp_sum(v);
  if (lane_id() == 0) smem[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < 32) {
    t = (threadIdx.x < (blockDim.x >> 5)) ? smem[threadIdx.x] : 0.0;
    t = warp_sum(t);
  }
  __syncthreads();
  return t;  // valid in warp 0
}

__global__ void __launch_bounds__(kOptThreads, 4)
grad_sqnorm_multi_kernel(const b200ssl_sgd_chunk* __restrict__ table, long long n_entries,
                         double* __restrict__ partials) {
  __shared__ double smem[kOptThreads / 32];
  for (long long c = blockIdx.x; c < n_entries; c += gridDim.x) {
    const b200ssl_sgd_chunk ent = table[c];
    const float* __restrict__ g = ent.grad;
    const int count = ent.count;
    double acc = 0.0;
    if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
      const int nvec = count >> 2;
      float4 gv[kOptVec];
#pragma unroll
      for (int j = 0; j < kOptVec; ++j) {
        const int v = threadIdx.x + j * kOptThreads;
        gv[j] = (v < nvec) ? ld_stream_f4(g + 4 * v) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < kOptVec; ++j) {
        acc = fma((double)gv[j].x, (double)gv[j].x, acc);
        acc = fma((double)gv[j].y, (double)gv[j].y, acc);
        acc = fma((double)gv[j].z, (double)gv[j].z, acc);
        acc = fma((double)gv[j].w, (double)gv[j].w, acc);
      }
      const int tail = (nvec << 2) + threadIdx.x;
      if (tail < count) acc = fma((double)g[tail], (double)g[tail], acc);
    } else {
      for (int i = threadIdx.x; i < count; i += kOptThreads) acc = fma((double)g[i], (double)g[i], acc);
    }
    const double t = block_sum_f64(acc, smem);
    if (threadIdx.x == 0) partials[c] = t;
  }
}

__global__ void __launch_bounds__(1024)
grad_norm_final_kernel(const double* __restrict__ partials, long long n_entries, float max_norm,
                       float* __restrict__ out /* [norm, clip coefficient] */) {
  __shared__ double smem[32];
  double acc = 0.0;
  for (long long i = threadIdx.x; i < n_entries; i += blockDim.x) acc += partials[i];
  const double t = block_sum_f64(acc, smem);
  if (threadIdx.x == 0) {
    const float norm = (float)sqrt(t);
    // clip_coef = max_norm / (total_norm + 1e-6): python float / tensor = reciprocal(tensor) * float
    const float coef = __fmul_rn(__frcp_rn(__fadd_rn(norm, 1e-6f)), max_norm);
    out[0] = norm;
    // torch.clamp(clip_coef, max=1.0) propagates NaN (and so poisons every gradient, like torch); fminf would
    // return 1.0 for a NaN coefficient
    out[1] = (coef != coef) ? coef : fminf(coef, 1.0f);
  }
}

__global__ void __launch_bounds__(kOptThreads, 4)
grad_scale_multi_kernel(const b200ssl_sgd_chunk* __restrict__ table, long long n_entries,
                        const float* __restrict__ coef_dev) {
  const float coef = *coef_dev;
  for (long long c = blockIdx.x; c < n_entries; c += gridDim.x) {
    const b200ssl_sgd_chunk ent = table[c];
    float* __restrict__ g = ent.grad;
    const int count = ent.count;
    if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
      const int nvec = count >> 2;
#pragma unroll
      for (int j = 0; j < kOptVec; ++j) {
        const int v = threadIdx.x + j * kOptThreads;
        if (v < nvec) {
          float4 x = ld_f4(g + 4 * v);
          x.x = __fmul_rn(x.x, coef); x.y = __fmul_rn(x.y, coef); x.z = __fmul_rn(x.z, coef); x.w = __fmul_rn(x.w, coef);
          st_stream_f4(g + 4 * v, x);
        }
      }
      const int tail = (nvec << 2) + threadIdx.x;
      if (tail < count) g[tail] = __fmul_rn(g[tail], coef);
    } else {
      for (int i = threadIdx.x; i < count; i += kOptThreads) g[i] = __fmul_rn(g[i], coef);
    }
  }
}

static int opt_grid(long long n_entries, int per_sm) {
  const long long max_grid = (long long)kNumSMs * per_sm;
  return (int)(n_entries < max_grid ? n_entries : max_grid);
}

}  // namespace b200ssl

extern "C" {

int64_t b200ssl_sgd_build_table_host(void* const* param_ptrs_host, void* const* grad_ptrs_host,
                                     void* const* momentum_ptrs_host, void* const* ema_ptrs_host,
                                     const int64_t* numels_host, int n_tensors, b200ssl_sgd_chunk* table_host,
                                     int64_t table_capacity) {
  using namespace b200ssl;
  if (!param_ptrs_host || !grad_ptrs_host || !numels_host || !table_host) {
    set_error("sgd_build_table: null argument");
    return B200SSL_EINVAL;
  }
  int64_t n = 0;
  for (int i = 0; i < n_tensors; ++i) {
    const int64_t numel = numels_host[i];
    if (numel < 0) {
      set_error("sgd_build_table: tensor %d has negative numel", i);
      return B200SSL_EINVAL;
    }
    void* ptrs[4] = {param_ptrs_host[i], grad_ptrs_host[i], momentum_ptrs_host ? momentum_ptrs_host[i] : nullptr,
                     ema_ptrs_host ? ema_ptrs_host[i] : nullptr};
    if (numel > 0 && (!ptrs[0] || !ptrs[1])) {
      set_error("sgd_build_table: tensor %d has a null parameter or gradient pointer", i);
      return B200SSL_EINVAL;
    }
    for (void* q : ptrs)
      if (reinterpret_cast<uintptr_t>(q) & 3u) {
        set_error("sgd_build_table: tensor %d is not 4-byte aligned", i);
        return B200SSL_EINVAL;
      }
    for (int64_t off = 0; off < numel; off += B200SSL_EMA_CHUNK) {
      if (n >= table_capacity) {
        set_error("sgd_build_table: table too small (%lld entries)", (long long)table_capacity);
        return B200SSL_EWORKSPACE;
      }
      const int64_t cnt = (numel - off < B200SSL_EMA_CHUNK) ? (numel - off) : B200SSL_EMA_CHUNK;
      b200ssl_sgd_chunk& t = table_host[n];
      t.param = static_cast<float*>(ptrs[0]) + off;
      t.grad = static_cast<float*>(ptrs[1]) + off;
      t.momentum = ptrs[2] ? static_cast<float*>(ptrs[2]) + off : nullptr;
      t.ema = ptrs[3] ? static_cast<float*>(ptrs[3]) + off : nullptr;
      t.count = (int32_t)cnt;
      t.tensor = i;
      ++n;
    }
  }
  return n;
}

size_t b200ssl_grad_norm_workspace_bytes(int64_t n_entries) {
  return (size_t)(n_entries > 0 ? n_entries : 1) * sizeof(double);
}

int b200ssl_grad_norm_multi(const b200ssl_sgd_chunk* table_dev, int64_t n_entries, double max_norm,
                            float* norm_and_coef_out, void* workspace, size_t workspace_bytes,
                            b200ssl_stream_t stream) {
  using namespace b200ssl;
  B200SSL_REQUIRE(n_entries >= 0 && norm_and_coef_out, "grad_norm_multi: bad argument");
  B200SSL_REQUIRE(n_entries == 0 || table_dev, "grad_norm_multi: null table");
  if (!workspace || workspace_bytes < b200ssl_grad_norm_workspace_bytes(n_entries)) {
    set_error("grad_norm_multi: workspace too small");
    return B200SSL_EWORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  double* partials = static_cast<double*>(workspace);
  if (n_entries > 0) {
    prof_begin("grad_sqnorm_multi", s);
    grad_sqnorm_multi_kernel<<<opt_grid(n_entries, 8), kOptThreads, 0, s>>>(table_dev, n_entries, partials);
    int rc = check_launch("grad_sqnorm_multi");
    if (rc) return rc;
  }
  prof_begin("grad_norm_final", s);
  grad_norm_final_kernel<<<1, 1024, 0, s>>>(partials, n_entries, (float)max_norm, norm_and_coef_out);
  return check_launch("grad_norm_final");
}

int b200ssl_grad_scale_multi(const b200ssl_sgd_chunk* table_dev, int64_t n_entries, const float* coef_dev,
                             b200ssl_stream_t stream) {
  using namespace b200ssl;
  B200SSL_REQUIRE(n_entries >= 0 && coef_dev, "grad_scale_multi: bad argument");
  if (n_entries == 0) return 0;
  B200SSL_REQUIRE(table_dev != nullptr, "grad_scale_multi: null table");
  prof_begin("grad_scale_multi", (cudaStream_t)stream);
  grad_scale_multi_kernel<<<opt_grid(n_entries, 8), kOptThreads, 0, (cudaStream_t)stream>>>(table_dev, n_entries, coef_dev);
  return check_launch("grad_scale_multi");
}

int b200ssl_sgd_ema_multi(const b200ssl_sgd_chunk* table_dev, int64_t n_entries, const float* coef_dev,
                          const b200ssl_sgd_hyper* h, b200ssl_stream_t stream) {
  using namespace b200ssl;
  B200SSL_REQUIRE(n_entries >= 0 && h, "sgd_ema_multi: bad argument");
  if (n_entries == 0) return 0;
  B200SSL_REQUIRE(table_dev != nullptr, "sgd_ema_multi: null table");
  B200SSL_REQUIRE(!(h->nesterov && (h->momentum <= 0.0 || h->dampening != 0.0)),
                  "sgd_ema_multi: Nesterov momentum requires a momentum and zero dampening");
  SgdConsts k;
  k.neg_lr = (float)(-h->lr);                       // add_(d, alpha=-lr)
  k.mu = (float)h->momentum;
  k.one_minus_damp = (float)(1.0 - h->dampening);   // evaluated in double like the python expression
  k.wd = (float)h->weight_decay;
  k.ema_a = (float)h->ema_alpha;
  k.ema_b = (float)(1.0 - h->ema_alpha);
  k.nesterov = h->nesterov != 0;
  k.first_step = h->first_step != 0;
  k.zero_grad = h->zero_grad != 0;
  k.has_wd = h->weight_decay != 0.0;
  k.has_mu = h->momentum != 0.0;
  k.has_ema = h->ema_alpha >= 0.0;
  prof_begin("sgd_ema_multi", (cudaStream_t)stream);
  sgd_ema_multi_kernel<<<opt_grid(n_entries, 6), kOptThreads, 0, (cudaStream_t)stream>>>(table_dev, n_entries, coef_dev, k);
  return check_launch("sgd_ema_multi");
}

}  // extern "C"
