// Multi-tensor mean-teacher EMA: every parameter tensor of the model in one launch.
// Reference semantics: mean_teacher.py:10-11 (mul_ then add_(alpha=)), two launches per tensor
// there; here a persistent grid walks a chunk table so hundreds of tiny tensors cost one launch.
//
// Roofline: HBM-bound, 12 B per parameter (read e, read p, write e).
#include "common.cuh"

namespace b200ssl {

constexpr int kEmaThreads = 256;
constexpr int kEmaVecPerThread = B200SSL_EMA_CHUNK / (kEmaThreads * 4);  // float4 per thread/chunk
static_assert(kEmaVecPerThread * kEmaThreads * 4 == B200SSL_EMA_CHUNK, "chunk/threads mismatch");

__device__ __forceinline__ float ema_one(float e, float p, float a, float b) {
  // t = RN(e*a); out = fma(p, b, t)  -- the exact two-op form ATen executes
  return __fmaf_rn(p, b, __fmul_rn(e, a));
}

__device__ __forceinline__ float4 ld_f4_noalloc(const float* p) {
  float4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p)
               : "memory");
  return r;
}

__global__ void __launch_bounds__(kEmaThreads, 4)
ema_multi_kernel(const b200ssl_ema_chunk* __restrict__ table, long long n_entries, float a,
                 float b) {
  for (long long c = blockIdx.x; c < n_entries; c += gridDim.x) {
    // the entry is the same for the whole block: one broadcast 16+8 byte read
    const b200ssl_ema_chunk ent = table[c];
    float* __restrict__ e = ent.ema;
    const float* __restrict__ p = ent.param;
    const int count = ent.count;
    const bool vec = (((reinterpret_cast<uintptr_t>(e) | reinterpret_cast<uintptr_t>(p)) & 15u) == 0);
    if (vec) {
      const int nvec = count >> 2;
      float4 ev[kEmaVecPerThread], pv[kEmaVecPerThread];
#pragma unroll
      for (int k = 0; k < kEmaVecPerThread; ++k) {
        const int v = threadIdx.x + k * kEmaThreads;
        if (v < nvec) {
          ev[k] = ld_f4_noalloc(e + 4 * v);
          pv[k] = ld_stream_f4(p + 4 * v);
        }
      }
#pragma unroll
      for (int k = 0; k < kEmaVecPerThread; ++k) {
        const int v = threadIdx.x + k * kEmaThreads;
        if (v < nvec) {
          float4 o;
          o.x = ema_one(ev[k].x, pv[k].x, a, b);
          o.y = ema_one(ev[k].y, pv[k].y, a, b);
          o.z = ema_one(ev[k].z, pv[k].z, a, b);
          o.w = ema_one(ev[k].w, pv[k].w, a, b);
          st_stream_f4(e + 4 * v, o);
        }
      }
      const int tail = (nvec << 2) + threadIdx.x;
      if (tail < count) e[tail] = ema_one(e[tail], p[tail], a, b);
    } else {
      for (int i = threadIdx.x; i < count; i += kEmaThreads) e[i] = ema_one(e[i], p[i], a, b);
    }
  }
}

}  // namespace b200ssl

extern "C" {

int64_t b200ssl_ema_table_entries(const int64_t* numels_host, int n_tensors) {
  if (!numels_host || n_tensors < 0) return B200SSL_EINVAL;
  int64_t n = 0;
  for (int i = 0; i < n_tensors; ++i) {
    if (numels_host[i] < 0) return B200SSL_EINVAL;
    n += (numels_host[i] + B200SSL_EMA_CHUNK - 1) / B200SSL_EMA_CHUNK;
  }
  return n;
}

int64_t b200ssl_ema_build_table_host(void* const* ema_ptrs_host, void* const* param_ptrs_host,
                                     const int64_t* numels_host, int n_tensors,
                                     b200ssl_ema_chunk* table_host, int64_t table_capacity) {
  if (!ema_ptrs_host || !param_ptrs_host || !numels_host || !table_host) {
    b200ssl::set_error("ema_build_table: null argument");
    return B200SSL_EINVAL;
  }
  int64_t n = 0;
  for (int i = 0; i < n_tensors; ++i) {
    const int64_t numel = numels_host[i];
    if (numel < 0) {
      b200ssl::set_error("ema_build_table: tensor %d has negative numel", i);
      return B200SSL_EINVAL;
    }
    if (numel > 0 && (!ema_ptrs_host[i] || !param_ptrs_host[i])) {
      b200ssl::set_error("ema_build_table: tensor %d has a null data pointer", i);
      return B200SSL_EINVAL;
    }
    if ((reinterpret_cast<uintptr_t>(ema_ptrs_host[i]) | reinterpret_cast<uintptr_t>(param_ptrs_host[i])) & 3u) {
      b200ssl::set_error("ema_build_table: tensor %d is not 4-byte aligned", i);
      return B200SSL_EINVAL;
    }
    for (int64_t off = 0; off < numel; off += B200SSL_EMA_CHUNK) {
      if (n >= table_capacity) {
        b200ssl::set_error("ema_build_table: table too small (%lld entries)", (long long)table_capacity);
        return B200SSL_EWORKSPACE;
      }
      const int64_t cnt = (numel - off < B200SSL_EMA_CHUNK) ? (numel - off) : B200SSL_EMA_CHUNK;
      table_host[n].ema = static_cast<float*>(ema_ptrs_host[i]) + off;
      table_host[n].param = static_cast<const float*>(param_ptrs_host[i]) + off;
      table_host[n].count = (int32_t)cnt;
      table_host[n].pad_ = i;
      ++n;
    }
  }
  return n;
}

int b200ssl_ema_multi(const b200ssl_ema_chunk* table_dev, int64_t n_entries, double alpha,
                      b200ssl_stream_t stream) {
  using namespace b200ssl;
  B200SSL_REQUIRE(n_entries >= 0, "ema_multi: negative entry count");
  if (n_entries == 0) return 0;
  B200SSL_REQUIRE(table_dev != nullptr, "ema_multi: null table");
  const float a = (float)alpha;          // mul_(alpha): python float -> fp32 scalar
  const float b = (float)(1.0 - alpha);  // add_(alpha=1.-alpha): evaluated in double, then fp32
  const long long max_grid = (long long)kNumSMs * 8;
  const int grid = (int)(n_entries < max_grid ? n_entries : max_grid);
  prof_begin("ema_multi", (cudaStream_t)stream);
  ema_multi_kernel<<<grid, kEmaThreads, 0, (cudaStream_t)stream>>>(table_dev, n_entries, a, b);
  return check_launch("ema_multi");
}

}  // extern "C"
