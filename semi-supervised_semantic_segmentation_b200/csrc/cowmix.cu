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
//
// Fast path (conv_tma_tile_kernel): the input tile of a 64-row x 64-column output block -- (64 + K - 1) rows of
// 64 columns -- is brought into shared memory by TMA (cp.async.bulk.tensor, 32-row boxes, one
// mbarrier per box so the first rows can be consumed while the rest is in flight).  TMA zero-fills
// rows/columns outside the image, which IS the convolution's zero padding, so the inner loop has
// no bounds checks at all: per input row one 8-byte shared load of the data pair, one broadcast
// 8-byte load of the duplicated tap, 16 FFMA2.  The four warps of a block share the tile (read
// amplification (64+K-1)/64 instead of (16+K-1)/16 from L2; eight warps x 128 rows measured slower).  Needs a 16-byte aligned base and a
// fast-axis extent that is a multiple of 4 (tensor-map stride rule); other shapes take the generic
// kernel below, which produces bit-identical results.
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

#include "common.cuh"

namespace b200ssl {

constexpr int kConvR = 16;         // slow-axis outputs per thread
constexpr int kConvThreads = 128;  // each thread covers 2 fast-axis positions

// ---- TMA tile geometry ----
constexpr int kTileCols = 64;                       // fast-axis columns per block (32 lanes x 2)
constexpr int kTileWarps = 4;
constexpr int kTileRowsOut = kTileWarps * kConvR;   // 64 output rows per block
constexpr int kBoxRows = 32;                        // rows per TMA box / mbarrier
constexpr int kMaxBoxes = 10;                       // 320 staged rows: enough for K <= 193
constexpr int kTileHeadBytes = 128;                 // mbarriers in front of the taps and the tile
constexpr int kBoxBytes = kBoxRows * kTileCols * 4; // 8 KiB

__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2),
      "r"(smem_u32(bar)) : "memory");
}

// in  : [n][A][B] through the tensor map (dims {B, A, n}, box {64, 32, 1})
// out : [n][B][A];  out[n][b][a] = sum_{i<K} taps[n][i] * in[n][a + i - K/2][b]
template <bool STATS>
__global__ void __launch_bounds__(kTileWarps * 32, 3)
conv_tma_tile_kernel(const __grid_constant__ CUtensorMap in_map, float* __restrict__ out,
                     const float* __restrict__ taps, int K, int A, int B,
                     double* __restrict__ partials) {
  constexpr int R = kConvR;
  extern __shared__ __align__(128) unsigned char conv_smem[];
  // layout: [mbarriers, 128 B][duplicated taps, (K + 3R) float2 rounded up to 128 B][n_boxes x 8 KiB tile]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(conv_smem);
  float2* wdup = reinterpret_cast<float2*>(conv_smem + kTileHeadBytes);                 // [K + 3R]
  float* tile = reinterpret_cast<float*>(conv_smem + kTileHeadBytes + (((K + 3 * kConvR) * 8 + 127) / 128) * 128);

  const int n = blockIdx.z;
  const int k = K >> 1;
  const int a_blk = blockIdx.y * kTileRowsOut;
  const int b_blk = blockIdx.x * kTileCols;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // steps per warp: t = 0 .. R+K-2 reads staged row 16*warp + t; padded to whole groups of R
  const int groups = (R + K - 1 + R - 1) / R;
  const int rows_staged = (kTileWarps - 1) * R + groups * R;
  const int n_boxes = (rows_staged + kBoxRows - 1) / kBoxRows;  // <= kMaxBoxes (checked on the host)

  if (threadIdx.x == 0) {
    for (int i = 0; i < n_boxes; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < n_boxes; ++i) {
      mbar_expect_tx(&bars[i], kBoxBytes);
      tma_load_3d(tile + i * kBoxRows * kTileCols, &in_map, b_blk, a_blk - k + i * kBoxRows, n, &bars[i]);
    }
  }
  const int wlen = K + 3 * R;
  for (int i = threadIdx.x; i < wlen; i += kTileWarps * 32) {
    const int t = i - R;
    const float w = (t >= 0 && t < K) ? __ldg(taps + (long long)n * K + t) : 0.f;
    wdup[i] = make_float2(w, w);
  }
  __syncthreads();

  float2 acc[R], wreg[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = make_float2(0.f, 0.f);
#pragma unroll
  for (int s = 0; s < R; ++s) wreg[s] = wdup[s];  // taps t - R (all zero): slots of "previous" group
  const float2* __restrict__ col = reinterpret_cast<const float2*>(tile) + lane;  // row pitch 32 float2

  for (int g = 0; g < groups; ++g) {
    const int row0 = (warp + g) * R;  // staged row of step s = 0 in this group
    if (g == 0 || (row0 & (kBoxRows - 1)) == 0) mbar_wait(&bars[row0 / kBoxRows], 0);  // first box / entering a new box
#pragma unroll
    for (int s = 0; s < R; ++s) {
      wreg[s] = wdup[g * R + s + R];
      const float2 v = col[(row0 + s) * (kTileCols / 2)];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = __ffma2_rn(v, wreg[(s - r + R) % R], acc[r]);
    }
  }

  // transposed store: for a fixed column the R outputs of this thread are contiguous in `out`
  const int a0 = a_blk + warp * R;
  float* __restrict__ oplane = out + (long long)n * A * B;
  double s1 = 0.0, s2 = 0.0;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int b = b_blk + 2 * lane + half;
    if (b < B && a0 < A) {
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
    __shared__ double red[2][kTileWarps];
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
      red[0][warp] = s1;
      red[1][warp] = s2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double t1 = 0.0, t2 = 0.0;
#pragma unroll
      for (int w = 0; w < kTileWarps; ++w) {
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

// tau_n = RN(RN(factor_n * std_n) + mean_n) from the per-block partial sums (cowmix.py:60-66);
// call with the first warp of a block, the result is valid in lane 0
__device__ __forceinline__ float cowmix_tau_from_partials(const double* __restrict__ partials,
                                                          int partials_per_sample, int n, long long plane,
                                                          float factor) {
  // fixed-order reduction of the per-block partial sums (deterministic)
  double s1 = 0.0, s2 = 0.0;
  for (int i = threadIdx.x; i < partials_per_sample; i += 32) {
    s1 += partials[((long long)n * partials_per_sample + i) * 2 + 0];
    s2 += partials[((long long)n * partials_per_sample + i) * 2 + 1];
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  const double M = (double)plane;
  const double mean = s1 / M;
  double var = (s2 - s1 * s1 / M) / (M - 1.0);  // unbiased; a 1-pixel plane gives NaN like torch
  if (var < 0.0) var = 0.0;
  const float stdf = (float)sqrt(var);
  return __fadd_rn(__fmul_rn(factor, stdf), (float)mean);
}

// one warp per sample: tau[n] for the fused threshold+mix kernel
__global__ void cowmix_tau_kernel(const float* __restrict__ thr_factor, const double* __restrict__ partials,
                                  int partials_per_sample, long long plane, float* __restrict__ tau) {
  const int n = blockIdx.x;
  const float t = cowmix_tau_from_partials(partials, partials_per_sample, n, plane, thr_factor[n]);
  if (threadIdx.x == 0) tau[n] = t;
}

// mask = (S > tau_n) with tau_n = RN(RN(factor_n * std_n) + mean_n)   (cowmix.py:60-68)
__global__ void __launch_bounds__(256)
cowmix_threshold_kernel(const float* __restrict__ S, const float* __restrict__ thr_factor,
                        const double* __restrict__ partials, int partials_per_sample,
                        long long plane, float* __restrict__ mask, bool vec) {
  __shared__ float tau_s;
  const int n = blockIdx.y;
  if (threadIdx.x < 32) {
    const float t = cowmix_tau_from_partials(partials, partials_per_sample, n, plane, thr_factor[n]);
    if (threadIdx.x == 0) tau_s = t;
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// the TMA path needs: 16-byte aligned base, row pitch a multiple of 16 bytes, at least one full
// column tile, and K small enough for the staged rows to fit the 8 boxes
static bool conv_tma_ok(const float* in, int K, int A, int B) {
  const int groups = (kConvR + K - 1 + kConvR - 1) / kConvR;
  const int rows = (kTileWarps - 1) * kConvR + groups * kConvR;
  return aligned16(in) && (B % 4 == 0) && B >= kTileCols && A >= 1 && rows <= kMaxBoxes * kBoxRows &&
         tensor_map_encoder() != nullptr;
}

struct ConvGrid {
  dim3 grid;
  int partials_per_sample;
};
static ConvGrid conv_tma_grid(int n, int A, int B) {
  ConvGrid g;
  g.grid = dim3((unsigned)((B + kTileCols - 1) / kTileCols), (unsigned)((A + kTileRowsOut - 1) / kTileRowsOut),
                (unsigned)n);
  g.partials_per_sample = (int)(g.grid.x * g.grid.y);
  return g;
}

template <bool STATS>
static int launch_conv_tma(const float* in, float* out, const float* taps, int K, int n, int A, int B,
                           double* partials, cudaStream_t s, const char* name) {
  CUtensorMap map;
  const cuuint64_t dims[3] = {(cuuint64_t)B, (cuuint64_t)A, (cuuint64_t)n};
  const cuuint64_t strides[2] = {(cuuint64_t)B * 4, (cuuint64_t)A * B * 4};
  const cuuint32_t box[3] = {kTileCols, kBoxRows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = tensor_map_encoder()(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(in), dims,
                                          strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("%s: cuTensorMapEncodeTiled failed (%d)", name, (int)r);
    return B200SSL_EUNSUPPORTED;
  }
  const int groups = (kConvR + K - 1 + kConvR - 1) / kConvR;
  const int n_boxes = ((kTileWarps - 1) * kConvR + groups * kConvR + kBoxRows - 1) / kBoxRows;
  const size_t smem = (size_t)kTileHeadBytes + (((size_t)(K + 3 * kConvR) * 8 + 127) / 128) * 128 + (size_t)n_boxes * kBoxBytes;
  auto kern = conv_tma_tile_kernel<STATS>;
  // the opt-in shared-memory limit is a per-DEVICE attribute of the function
  static bool attr_done[kMaxDevices] = {};
  const int dev = current_device_slot();
  if (!attr_done[dev]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(kTileHeadBytes + (((193 + 3 * kConvR) * 8 + 127) / 128) * 128 + kMaxBoxes * kBoxBytes)) ==
        cudaSuccess)
      attr_done[dev] = dev != kMaxDevices - 1;
  }
  const ConvGrid g = conv_tma_grid(n, A, B);
  prof_begin(name, s);
  kern<<<g.grid, kTileWarps * 32, smem, s>>>(map, out, taps, K, A, B, partials);
  return check_launch(name);
}

static ConvGrid conv_grid(int n, int A, int B) {  // generic kernel
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
  const int pps = max(conv_grid(n, w, h).partials_per_sample, conv_tma_grid(n, w, h).partials_per_sample);
  size_t bytes = 0;
  bytes += align_up((size_t)n * plane * sizeof(float), 256);  // Vt
  bytes += align_up((size_t)n * plane * sizeof(float), 256);  // S (when field_out is NULL)
  bytes += align_up((size_t)n * pps * 2 * sizeof(double), 256);
  bytes += align_up((size_t)n * sizeof(float), 256);          // tau scratch (fused threshold+mix in the step)
  return bytes;
}

// conv passes: noise -> S (+ per-block statistics); shared by b200ssl_cowmix_mask / _field
static int cowmix_field_impl(const float* noise, const float* taps, int K, int n, int h, int w, float* field_out,
                             void* workspace, size_t workspace_bytes, cudaStream_t s, const char* who,
                             float** S_out, double** partials_out, int* pps_out, float** tau_ws_out) {
  using namespace b200ssl;
  B200SSL_REQUIRE(n <= 65535, "%s: batch too large", who);
  B200SSL_REQUIRE(K >= 1 && (K & 1) == 1, "%s: K must be odd and >= 1 (got %d)", who, K);
  const size_t need = b200ssl_cowmix_workspace_bytes(n, h, w);
  if (!workspace || workspace_bytes < need) {
    set_error("%s: workspace too small (%zu < %zu)", who, workspace_bytes, need);
    return B200SSL_EWORKSPACE;
  }
  const size_t plane = (size_t)h * w;
  char* ws = static_cast<char*>(workspace);
  float* Vt = reinterpret_cast<float*>(ws);
  ws += align_up((size_t)n * plane * sizeof(float), 256);
  float* S = field_out ? field_out : reinterpret_cast<float*>(ws);
  ws += align_up((size_t)n * plane * sizeof(float), 256);
  double* partials = reinterpret_cast<double*>(ws);
  const int pps_max = max(conv_grid(n, w, h).partials_per_sample, conv_tma_grid(n, w, h).partials_per_sample);
  ws += align_up((size_t)n * pps_max * 2 * sizeof(double), 256);
  *tau_ws_out = reinterpret_cast<float*>(ws);

  const size_t smem = (size_t)(K + 3 * kConvR) * sizeof(float2);
  B200SSL_REQUIRE(smem <= 48 * 1024, "%s: K=%d too large", who, K);

  // pass 1: along H (slow axis of noise[n][H][W]) -> Vt[n][W][H]
  if (conv_tma_ok(noise, K, h, w)) {
    int rc = launch_conv_tma<false>(noise, Vt, taps, K, n, h, w, nullptr, s, "cowmix_conv_pass1");
    if (rc) return rc;
  } else {
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
  const bool tma2 = conv_tma_ok(Vt, K, w, h);
  const ConvGrid g2 = tma2 ? conv_tma_grid(n, w, h) : conv_grid(n, w, h);
  if (tma2) {
    int rc = launch_conv_tma<true>(Vt, S, taps, K, n, w, h, partials, s, "cowmix_conv_pass2");
    if (rc) return rc;
  } else {
    const bool vec2 = (h % 2 == 0);
    prof_begin("cowmix_conv_pass2", s);
    if (vec2)
      conv_slow_axis_transposed<true, true><<<g2.grid, kConvThreads, smem, s>>>(Vt, S, taps, K, w, h, partials);
    else
      conv_slow_axis_transposed<false, true><<<g2.grid, kConvThreads, smem, s>>>(Vt, S, taps, K, w, h, partials);
    int rc = check_launch("cowmix conv pass 2");
    if (rc) return rc;
  }
  *S_out = S;
  *partials_out = partials;
  *pps_out = g2.partials_per_sample;
  return 0;
}

int b200ssl_cowmix_field(const float* noise, const float* taps, int K, const float* thr_factor, int n, int h,
                         int w, float* field_out, float* tau_out, void* workspace, size_t workspace_bytes,
                         b200ssl_stream_t stream) {
  using namespace b200ssl;
  B200SSL_REQUIRE(n >= 0 && h >= 0 && w >= 0, "cowmix_field: negative extent");
  if (n == 0 || h == 0 || w == 0) return 0;
  B200SSL_REQUIRE(noise && taps && thr_factor && field_out && tau_out, "cowmix_field: null argument");
  cudaStream_t s = (cudaStream_t)stream;
  float *S, *tau_ws;
  double* partials;
  int pps;
  int rc = cowmix_field_impl(noise, taps, K, n, h, w, field_out, workspace, workspace_bytes, s, "cowmix_field", &S,
                             &partials, &pps, &tau_ws);
  if (rc) return rc;
  prof_begin("cowmix_tau", s);
  cowmix_tau_kernel<<<n, 32, 0, s>>>(thr_factor, partials, pps, (long long)h * w, tau_out);
  return check_launch("cowmix tau");
}

int b200ssl_cowmix_mask(const float* noise, const float* taps, int K, const float* thr_factor,
                        int n, int h, int w, float* mask_out, float* field_out, void* workspace,
                        size_t workspace_bytes, b200ssl_stream_t stream) {
  using namespace b200ssl;
  B200SSL_REQUIRE(n >= 0 && h >= 0 && w >= 0, "cowmix_mask: negative extent");
  if (n == 0 || h == 0 || w == 0) return 0;
  B200SSL_REQUIRE(K >= 1 && (K & 1) == 1, "cowmix_mask: K must be odd and >= 1 (got %d)", K);
  B200SSL_REQUIRE(noise && taps && thr_factor && mask_out, "cowmix_mask: null argument");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t plane = (size_t)h * w;
  float *S, *tau_ws;
  double* partials;
  int pps;
  int rc = cowmix_field_impl(noise, taps, K, n, h, w, field_out, workspace, workspace_bytes, s, "cowmix_mask", &S,
                             &partials, &pps, &tau_ws);
  if (rc) return rc;
  {
    const bool vec = (plane % 4 == 0) && aligned16(S) && aligned16(mask_out);
    long long bx = (long long)((vec ? plane / 4 : plane) + 255) / 256;
    long long cap = (long long)kNumSMs * 8 / n;
    if (cap < 1) cap = 1;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    prof_begin("cowmix_threshold", s);
    cowmix_threshold_kernel<<<dim3((unsigned)bx, (unsigned)n), 256, 0, s>>>(
        S, thr_factor, partials, pps, (long long)plane, mask_out, vec);
    return check_launch("cowmix threshold");
  }
}

}  // extern "C"
