// CowMix mask generation: separable per-sample Gaussian smoothing of a noise field, per-sample
// mean / unbiased std, threshold.  Reference semantics: cowmix.py:27-37 (two depthwise conv2d
// calls with groups=N) and cowmix.py:56-68 (mean, std, erfinv threshold, compare).
//
// Design (B200):
//   * The two 1-D passes are the SAME kernel run twice: "convolve along the slow axis, coalesced
//     along the fast axis, write the result transposed".  Pass 1 maps noise[n][H][W] to
//     Vt[n][W][H]; pass 2 maps Vt to S[n][H][W].  The intermediate (P floats) lives in L2.
//   * This stage is fp32-FMA bound, not HBM bound: 2K FMA per pixel (K up to 193) against 8
//     compulsory bytes.  Each thread owns 2 adjacent fast-axis positions x R=16 slow-axis outputs
//     and uses the packed FFMA2 instruction (two fp32 FMAs per lane per issue, the only way to
//     reach the fp32 peak on sm_100).  Every loaded input pair feeds 16 FFMA2; the per-sample
//     tap (duplicated into both halves of a 64-bit word) is one broadcast shared-memory load.
//   * Accumulation order is fixed: taps ascending, one fma per tap, starting from +0 -- the
//     C oracle (oracle/oracle.c) does exactly the same and matches bit for bit.
//   * Statistics are accumulated in fp64 per thread and reduced in a fixed order (deterministic).
#include "common.cuh"

namespace b200ssl {

constexpr int kConvR = 16;         // slow-axis outputs per thread
constexpr int kConvThreads = 128;  // each thread covers 2 fast-axis positions

template <bool VEC2>
__device__ __forceinline__ float2 load_pair(const float* __restrict__ row, int col, int B) {
  float2 v = make_float2(0.f, 0.f);
  if (VEC2) {
    if (col + 1 < B) {
      v = __ldg(reinterpret_cast<const float2*>(row + col));
    } else if (col < B) {
      v.x = __ldg(row + col);
    }
  } else {
    if (col < B) v.x = __ldg(row + col);
    if (col + 1 < B) v.y = __ldg(row + col + 1);
  }
  return v;
}

// in  : [n][A][B]  (B contiguous)      out : [n][B][A]  (A contiguous)
// out[n][b][a] = sum_{i<K} taps[n][i] * in[n][a + i - K/2][b]    (zero padding along A)
template <bool VEC2, bool STATS>
__global__ void __launch_bounds__(kConvThreads)
conv_slow_axis_transposed(const float* __restrict__ in, float* __restrict__ out,
                          const float* __restrict__ taps, int K, int A, int B,
                          double* __restrict__ partials) {
  constexpr int R = kConvR;
  extern __shared__ float2 wdup[];  // wdup[t + R] = (w[t], w[t]) for 0<=t<K, zero elsewhere
  const int n = blockIdx.z;
  const int k = K >> 1;
  const int wlen = K + 3 * R;
  for (int i = threadIdx.x; i < wlen; i += kConvThreads) {
    const int t = i - R;
    const float w = (t >= 0 && t < K) ? __ldg(taps + (long long)n * K + t) : 0.f;
    wdup[i] = make_float2(w, w);
  }
  __syncthreads();

  const int a0 = blockIdx.y * R;
  const int col = (blockIdx.x * kConvThreads + threadIdx.x) * 2;
  const float* __restrict__ plane = in + (long long)n * A * B;

  // step t reads input row a0 + t - k; valid rows give t in [t_lo, t_hi)
  const int t_lo = max(0, k - a0);
  const int t_hi = min(R + K - 1, A - a0 + k);
  const int tb_lo = t_lo / R;
  const int tb_hi = (t_hi + R - 1) / R;

  float2 acc[R], wreg[R], vcur[R], vnext[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = make_float2(0.f, 0.f);
  // wreg[s] holds w[t] for the latest t with t % R == s; preload the taps before the first step
#pragma unroll
  for (int s = 0; s < R; ++s) wreg[s] = wdup[(tb_lo - 1) * R + s + R];

  auto load_group = [&](int tb, float2* v) {
#pragma unroll
    for (int s = 0; s < R; ++s) {
      const int a_in = a0 + tb * R + s - k;
      v[s] = (a_in >= 0 && a_in < A) ? load_pair<VEC2>(plane + (long long)a_in * B, col, B)
                                     : make_float2(0.f, 0.f);
    }
  };

  if (tb_lo < tb_hi) load_group(tb_lo, vcur);
  for (int tb = tb_lo; tb < tb_hi; ++tb) {
    if (tb + 1 < tb_hi) load_group(tb + 1, vnext);
#pragma unroll
    for (int s = 0; s < R; ++s) {
      wreg[s] = wdup[tb * R + s + R];
      const float2 v = vcur[s];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        // output r uses tap t - r, held in wreg[(s - r) mod R]
        acc[r] = __ffma2_rn(v, wreg[(s - r + R) % R], acc[r]);
      }
    }
#pragma unroll
    for (int s = 0; s < R; ++s) vcur[s] = vnext[s];
  }

  // transposed store: for a fixed fast-axis position b the R slow-axis outputs are contiguous
  float* __restrict__ oplane = out + (long long)n * A * B;
  double s1 = 0.0, s2 = 0.0;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int b = col + half;
    if (b < B) {
      float* dst = oplane + (long long)b * A + a0;
      float vals[R];
#pragma unroll
      for (int r = 0; r < R; ++r) vals[r] = half ? acc[r].y : acc[r].x;
      if (a0 + R <= A && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
#pragma unroll
        for (int q = 0; q < R / 4; ++q)
          *reinterpret_cast<float4*>(dst + 4 * q) =
              make_float4(vals[4 * q], vals[4 * q + 1], vals[4 * q + 2], vals[4 * q + 3]);
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (a0 + r < A) dst[r] = vals[r];
      }
      if (STATS) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (a0 + r < A) {
            const double x = (double)vals[r];
            s1 += x;
            s2 += x * x;
          }
        }
      }
    }
  }
  if (STATS) {
    __shared__ double red[2][kConvThreads / 32];
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane_id() == 0) {
      red[0][threadIdx.x >> 5] = s1;
      red[1][threadIdx.x >> 5] = s2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double t1 = 0.0, t2 = 0.0;
#pragma unroll
      for (int w = 0; w < kConvThreads / 32; ++w) {
        t1 += red[0][w];
        t2 += red[1][w];
      }
      const long long blk = (long long)blockIdx.y * gridDim.x + blockIdx.x;
      const long long per_sample = (long long)gridDim.x * gridDim.y;
      partials[(n * per_sample + blk) * 2 + 0] = t1;
      partials[(n * per_sample + blk) * 2 + 1] = t2;
    }
  }
}

// mask = (S > tau_n) with tau_n = RN(RN(factor_n * std_n) + mean_n)   (cowmix.py:60-68)
__global__ void __launch_bounds__(256)
cowmix_threshold_kernel(const float* __restrict__ S, const float* __restrict__ thr_factor,
                        const double* __restrict__ partials, int partials_per_sample,
                        long long plane, float* __restrict__ mask, bool vec) {
  __shared__ float tau_s;
  const int n = blockIdx.y;
  if (threadIdx.x < 32) {
    // fixed-order reduction of the per-block partial sums (deterministic)
    double s1 = 0.0, s2 = 0.0;
    for (int i = threadIdx.x; i < partials_per_sample; i += 32) {
      s1 += partials[((long long)n * partials_per_sample + i) * 2 + 0];
      s2 += partials[((long long)n * partials_per_sample + i) * 2 + 1];
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (threadIdx.x == 0) {
      const double M = (double)plane;
      const double mean = s1 / M;
      double var = (s2 - s1 * s1 / M) / (M - 1.0);  // unbiased; a 1-pixel plane gives NaN like torch
      if (var < 0.0) var = 0.0;
      const float stdf = (float)sqrt(var);
      tau_s = __fadd_rn(__fmul_rn(thr_factor[n], stdf), (float)mean);
    }
  }
  __syncthreads();
  const float tau = tau_s;
  const float* __restrict__ s = S + (long long)n * plane;
  float* __restrict__ m = mask + (long long)n * plane;
  if (vec) {
    const long long nv = plane >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv;
         i += (long long)gridDim.x * blockDim.x) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(s) + i);
      float4 o;
      o.x = v.x > tau ? 1.f : 0.f;
      o.y = v.y > tau ? 1.f : 0.f;
      o.z = v.z > tau ? 1.f : 0.f;
      o.w = v.w > tau ? 1.f : 0.f;
      st_stream_f4(m + 4 * i, o);
    }
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < plane;
         i += (long long)gridDim.x * blockDim.x)
      m[i] = s[i] > tau ? 1.f : 0.f;
  }
}

struct ConvGrid {
  dim3 grid;
  int partials_per_sample;
};
static ConvGrid conv_grid(int n, int A, int B) {
  ConvGrid g;
  g.grid = dim3((unsigned)((B + 2 * kConvThreads - 1) / (2 * kConvThreads)),
                (unsigned)((A + kConvR - 1) / kConvR), (unsigned)n);
  g.partials_per_sample = (int)(g.grid.x * g.grid.y);
  return g;
}

}  // namespace b200ssl

extern "C" {

size_t b200ssl_cowmix_workspace_bytes(int n, int h, int w) {
  using namespace b200ssl;
  if (n <= 0 || h <= 0 || w <= 0) return 0;
  const size_t plane = (size_t)h * w;
  const ConvGrid g2 = conv_grid(n, w, h);
  size_t bytes = 0;
  bytes += align_up((size_t)n * plane * sizeof(float), 256);  // Vt
  bytes += align_up((size_t)n * plane * sizeof(float), 256);  // S (when field_out is NULL)
  bytes += align_up((size_t)n * g2.partials_per_sample * 2 * sizeof(double), 256);
  return bytes;
}

int b200ssl_cowmix_mask(const float* noise, const float* taps, int K, const float* thr_factor,
                        int n, int h, int w, float* mask_out, float* field_out, void* workspace,
                        size_t workspace_bytes, b200ssl_stream_t stream) {
  using namespace b200ssl;
  B200SSL_REQUIRE(n >= 0 && h >= 0 && w >= 0, "cowmix_mask: negative extent");
  if (n == 0 || h == 0 || w == 0) return 0;
  B200SSL_REQUIRE(n <= 65535, "cowmix_mask: batch too large");
  B200SSL_REQUIRE(K >= 1 && (K & 1) == 1, "cowmix_mask: K must be odd and >= 1 (got %d)", K);
  B200SSL_REQUIRE(noise && taps && thr_factor && mask_out, "cowmix_mask: null argument");
  const size_t need = b200ssl_cowmix_workspace_bytes(n, h, w);
  if (!workspace || workspace_bytes < need) {
    set_error("cowmix_mask: workspace too small (%zu < %zu)", workspace_bytes, need);
    return B200SSL_EWORKSPACE;
  }
  const size_t plane = (size_t)h * w;
  char* ws = static_cast<char*>(workspace);
  float* Vt = reinterpret_cast<float*>(ws);
  ws += align_up((size_t)n * plane * sizeof(float), 256);
  float* S = field_out ? field_out : reinterpret_cast<float*>(ws);
  ws += align_up((size_t)n * plane * sizeof(float), 256);
  double* partials = reinterpret_cast<double*>(ws);

  cudaStream_t s = (cudaStream_t)stream;
  const size_t smem = (size_t)(K + 3 * kConvR) * sizeof(float2);
  B200SSL_REQUIRE(smem <= 48 * 1024, "cowmix_mask: K=%d too large", K);

  // pass 1: along H (slow axis of noise[n][H][W]) -> Vt[n][W][H]
  {
    const ConvGrid g = conv_grid(n, h, w);
    const bool vec2 = (w % 2 == 0) && ((reinterpret_cast<uintptr_t>(noise) & 7u) == 0);
    prof_begin("cowmix_conv_pass1", s);
    if (vec2)
      conv_slow_axis_transposed<true, false><<<g.grid, kConvThreads, smem, s>>>(noise, Vt, taps, K, h, w, nullptr);
    else
      conv_slow_axis_transposed<false, false><<<g.grid, kConvThreads, smem, s>>>(noise, Vt, taps, K, h, w, nullptr);
    int rc = check_launch("cowmix conv pass 1");
    if (rc) return rc;
  }
  // pass 2: along W (slow axis of Vt[n][W][H]) -> S[n][H][W], with per-block statistics
  const ConvGrid g2 = conv_grid(n, w, h);
  {
    const bool vec2 = (h % 2 == 0);
    prof_begin("cowmix_conv_pass2", s);
    if (vec2)
      conv_slow_axis_transposed<true, true><<<g2.grid, kConvThreads, smem, s>>>(Vt, S, taps, K, w, h, partials);
    else
      conv_slow_axis_transposed<false, true><<<g2.grid, kConvThreads, smem, s>>>(Vt, S, taps, K, w, h, partials);
    int rc = check_launch("cowmix conv pass 2");
    if (rc) return rc;
  }
  {
    const bool vec = (plane % 4 == 0) && aligned16(S) && aligned16(mask_out);
    long long bx = (long long)((vec ? plane / 4 : plane) + 255) / 256;
    long long cap = (long long)kNumSMs * 8 / n;
    if (cap < 1) cap = 1;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    prof_begin("cowmix_threshold", s);
    cowmix_threshold_kernel<<<dim3((unsigned)bx, (unsigned)n), 256, 0, s>>>(
        S, thr_factor, partials, g2.partials_per_sample, (long long)plane, mask_out, vec);
    return check_launch("cowmix threshold");
  }
}

}  // extern "C"
