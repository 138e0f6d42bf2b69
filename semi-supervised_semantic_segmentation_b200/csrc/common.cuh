// Shared helpers for the b200ssl kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "b200ssl.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "b200ssl kernels are written for sm_100a (B200) only"
#endif

namespace b200ssl {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
// optional per-kernel timing (b200ssl_prof_enable): prof_begin records a start event on the stream,
// check_launch / prof_end the matching stop event; both are a single flag test when profiling is off
void prof_begin(const char* name, cudaStream_t stream);
void prof_end();

inline int check_launch(const char* what) {
  prof_end();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  count_launch();
  return 0;
}

#define B200SSL_REQUIRE(cond, ...)      \
  do {                                  \
    if (!(cond)) {                      \
      ::b200ssl::set_error(__VA_ARGS__); \
      return B200SSL_EINVAL;            \
    }                                   \
  } while (0)

// per-device caches (function attributes, side streams) are indexed by this; devices beyond the table
// share the last slot, which callers never cache
constexpr int kMaxDevices = 65;
inline int current_device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices - 1) return kMaxDevices - 1;
  return dev;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- streaming 128-bit accesses: data touched once should not pollute L1 -----------------
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_f4(float* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float ld_stream_f1(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}

// torch.argmax's ordering: x beats the running best iff x > best, or x is NaN while best is not (the first maximum
// wins, NaN counts as the maximum).  Equivalent to !(x <= best) && best == best: two predicate instructions, no branch.
__device__ __forceinline__ bool argmax_beats(float x, float best) { return !(x <= best) & (best == best); }

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One warp-wide step: every lane contributes `bin` (or -1 for "nothing"); equal adjacent lanes merge.
__device__ __forceinline__ void warp_run_add(unsigned* hist, int bin) {
  const unsigned lane = lane_id();
  const int prev = __shfl_up_sync(0xffffffffu, bin, 1);
  const bool head = (lane == 0) || (bin != prev);
  const unsigned heads = __ballot_sync(0xffffffffu, head);
  if (head && bin >= 0) {
    const unsigned later = (lane == 31) ? 0u : (heads & (0xffffffffu << (lane + 1)));
    const int end = later ? (__ffs(later) - 1) : 32;
    atomicAdd(hist + bin, (unsigned)(end - (int)lane));
  }
}

// As warp_run_add, but every contributing lane stands for `weight` identical elements.
__device__ __forceinline__ void warp_run_add_weighted(unsigned* hist, int bin, unsigned weight) {
  const unsigned lane = lane_id();
  const int prev = __shfl_up_sync(0xffffffffu, bin, 1);
  const bool head = (lane == 0) || (bin != prev);
  const unsigned heads = __ballot_sync(0xffffffffu, head);
  if (head && bin >= 0) {
    const unsigned later = (lane == 31) ? 0u : (heads & (0xffffffffu << (lane + 1)));
    const int end = later ? (__ffs(later) - 1) : 32;
    atomicAdd(hist + bin, (unsigned)(end - (int)lane) * weight);
  }
}

// N consecutive elements per lane (labels are spatially coherent, so the N bins of a lane almost
// always agree): lanes whose N bins are equal go through ONE warp-merged atomic with weight N, the
// others fall back to one atomic per element.  Warp-collective: call with all 32 lanes.
template <int N>
__device__ __forceinline__ void lane_run_add(unsigned* hist, const int (&bin)[N]) {
  bool uniform = true;
#pragma unroll
  for (int e = 1; e < N; ++e) uniform = uniform && (bin[e] == bin[0]);
  warp_run_add_weighted(hist, uniform ? bin[0] : -1, (unsigned)N);
  if (!uniform) {
#pragma unroll
    for (int e = 0; e < N; ++e)
      if (bin[e] >= 0) atomicAdd(hist + bin[e], 1u);
  }
}
__device__ __forceinline__ void quad_run_add(unsigned* hist, const int (&bin)[4]) { lane_run_add<4>(hist, bin); }

// labels come as int64 (torch default), int32 or uint8; always widened to int64 for compares
template <int DT>
struct LabelT;
template <>
struct LabelT<B200SSL_I64> { using type = long long; };
template <>
struct LabelT<B200SSL_I32> { using type = int; };
template <>
struct LabelT<B200SSL_U8> { using type = unsigned char; };

}  // namespace b200ssl
