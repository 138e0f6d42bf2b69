// Small reductions around the loss path: channel argmax of the soft one-hot target
// (losses.py:240 `torch.argmax(target, dim=1)` + the per-sample `tgt.sum() > 0` of :246) and
// metrics.dice_metric (metrics.py:1-7).
#include "common.cuh"

namespace b200ssl {

// torch.argmax order: first maximal element wins, NaN counts as the maximum
__device__ __forceinline__ void argmax_step(float x, int c, float& best, int& arg) {
  const bool take = (arg < 0) || argmax_beats(x, best);
  if (take) { best = x; arg = c; }
}

template <typename OUT>
__global__ void __launch_bounds__(256)
argmax_channels_kernel(const float* __restrict__ x, int C, long long hw, OUT* __restrict__ out,
                       int* __restrict__ nonzero, bool vec) {
  const int n = blockIdx.y;
  const float* __restrict__ xp = x + (long long)n * C * hw;
  OUT* __restrict__ op = out + (long long)n * hw;
  int nz = 0;
  const long long quads = (hw + 3) / 4;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < quads;
       q += (long long)gridDim.x * blockDim.x) {
    const long long first = q * 4;
    float best[4];
    int arg[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) { best[e] = 0.f; arg[e] = -1; }
    if (vec && first + 4 <= hw) {
      for (int c0 = 0; c0 < C; c0 += 4) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (c0 + u < C) v[u] = ld_stream_f4(xp + (long long)(c0 + u) * hw + first);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (c0 + u < C) {
            argmax_step(v[u].x, c0 + u, best[0], arg[0]);
            argmax_step(v[u].y, c0 + u, best[1], arg[1]);
            argmax_step(v[u].z, c0 + u, best[2], arg[2]);
            argmax_step(v[u].w, c0 + u, best[3], arg[3]);
          }
        }
      }
    } else {
      for (int e = 0; e < 4; ++e)
        if (first + e < hw)
          for (int c = 0; c < C; ++c) argmax_step(xp[(long long)c * hw + first + e], c, best[e], arg[e]);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (first + e < hw) {
        op[first + e] = (OUT)arg[e];
        nz += (arg[e] != 0);
      }
    }
  }
  if (nonzero) {
    nz = warp_sum(nz);
    if (lane_id() == 0 && nz) atomicAdd(nonzero + n, nz);
  }
}

constexpr int kDiceThreads = 256;

__global__ void __launch_bounds__(kDiceThreads)
dice_partial_kernel(const float* __restrict__ x, const float* __restrict__ y, long long chw,
                    double* __restrict__ partials, bool vec) {
  const int n = blockIdx.y;
  const float* __restrict__ xp = x + (long long)n * chw;
  const float* __restrict__ yp = y + (long long)n * chw;
  double si = 0.0, sc = 0.0;
  if (vec) {
    const long long nv = chw >> 2;
    for (long long i = (long long)blockIdx.x * kDiceThreads + threadIdx.x; i < nv;
         i += (long long)gridDim.x * kDiceThreads) {
      const float4 a = ld_stream_f4(xp + 4 * i), b = ld_stream_f4(yp + 4 * i);
      si += (double)__fmul_rn(a.x, b.x) + (double)__fmul_rn(a.y, b.y) + (double)__fmul_rn(a.z, b.z) + (double)__fmul_rn(a.w, b.w);
      sc += (double)__fadd_rn(a.x, b.x) + (double)__fadd_rn(a.y, b.y) + (double)__fadd_rn(a.z, b.z) + (double)__fadd_rn(a.w, b.w);
    }
  } else {
    for (long long i = (long long)blockIdx.x * kDiceThreads + threadIdx.x; i < chw;
         i += (long long)gridDim.x * kDiceThreads) {
      si += (double)__fmul_rn(xp[i], yp[i]);
      sc += (double)__fadd_rn(xp[i], yp[i]);
    }
  }
  __shared__ double red[2][kDiceThreads / 32];
  si = warp_sum(si);
  sc = warp_sum(sc);
  if (lane_id() == 0) { red[0][threadIdx.x >> 5] = si; red[1][threadIdx.x >> 5] = sc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < kDiceThreads / 32; ++w) { a += red[0][w]; b += red[1][w]; }
    partials[((long long)n * gridDim.x + blockIdx.x) * 2 + 0] = a;
    partials[((long long)n * gridDim.x + blockIdx.x) * 2 + 1] = b;
  }
}

__global__ void dice_final_kernel(const double* __restrict__ partials, int per_sample, int n,
                                  float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a = 0.0, b = 0.0;
  for (int k = 0; k < per_sample; ++k) {
    a += partials[((long long)i * per_sample + k) * 2 + 0];
    b += partials[((long long)i * per_sample + k) * 2 + 1];
  }
  const float inter = (float)a, card = (float)b;
  out[i] = __fdiv_rn(__fadd_rn(__fmul_rn(2.0f, inter), 1.0f), __fadd_rn(card, 1.0f));
}

static int dice_blocks(int n, long long chw) {
  long long bx = (chw / 4 + kDiceThreads * 4 - 1) / (kDiceThreads * 4);
  long long cap = (long long)kNumSMs * 8 / (n > 0 ? n : 1);
  if (cap < 1) cap = 1;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  return (int)bx;
}

// ------------------------------------------------------------------------------------------
// Validation metric in one pass (train.py:171-175): argmax over the channels of the LOW-RESOLUTION logits,
// one_hot, nearest-neighbour resize to the mask size, mask > 0.5, Dice of channel `fg` -- six ATen kernels and
// three full-resolution temporaries in the reference.  Here every mask pixel looks up the argmax of its
// nearest low-resolution source pixel (ATen `nearest`: src = min(floor(dst * in/out), in - 1) with in/out
// evaluated in fp32) and the per-image 2x2 confusion matrix {TN, FP, FN, TP} of (mask > thr, argmax == fg)
// is accumulated as integers; b200ssl_dice_from_cm turns it into metrics.py:1-7's value.
// grid = (blocks, n); mask channel `fg` of mask [n, mask_channels, H, W].
// ------------------------------------------------------------------------------------------
// One warp per mask row (no per-pixel division), 4 consecutive pixels per lane (one 128-bit mask load when the
// rows are 16-byte aligned); the argmax of a source pixel is reused by the neighbours that map to it.
__global__ void __launch_bounds__(256)
validation_cm_kernel(const float* __restrict__ logits, int C, int h, int w, const float* __restrict__ mask,
                     int mask_channels, int H, int W, float thr, int fg, bool vec,
                     unsigned long long* __restrict__ cm) {
  const int n = blockIdx.y;
  const float* __restrict__ lp = logits + (long long)n * C * h * w;
  const float* __restrict__ mp = mask + ((long long)n * mask_channels + fg) * (long long)H * W;
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  const long long hw_low = (long long)h * w;
  const int lane = (int)lane_id();
  unsigned n_px = 0, n_m = 0, n_p = 0, n_mp = 0;          // pixels, mask fg, predicted fg, both
  for (int y = blockIdx.x * 8 + (int)(threadIdx.x >> 5); y < H; y += (int)gridDim.x * 8) {
    const int ys = min((int)floorf((float)y * sy), h - 1);
    const float* __restrict__ lrow = lp + (long long)ys * w;
    const float* __restrict__ mrow = mp + (long long)y * W;
    for (int x0 = lane * 4; x0 < W; x0 += 128) {
      float m[4] = {0.f, 0.f, 0.f, 0.f};
      if (vec) {
        const float4 v = ld_stream_f4(mrow + x0);
        m[0] = v.x; m[1] = v.y; m[2] = v.z; m[3] = v.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (x0 + e < W) m[e] = ld_stream_f1(mrow + x0 + e);
      }
      int last_xs = -1, p_fg = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (x0 + e < W) {
          const int xs = min((int)floorf((float)(x0 + e) * sx), w - 1);
          if (xs != last_xs) {
            const float* __restrict__ q = lrow + xs;
            float best = __ldg(q);
            int arg = 0;
            for (int c = 1; c < C; ++c) {   // torch.argmax: first maximum wins, NaN counts as the maximum
              const float v = __ldg(q + c * hw_low);
              if (argmax_beats(v, best)) { best = v; arg = c; }
            }
            p_fg = arg == fg ? 1 : 0;
            last_xs = xs;
          }
          const unsigned m_fg = m[e] > thr ? 1u : 0u;
          n_px += 1u;
          n_m += m_fg;
          n_p += (unsigned)p_fg;
          n_mp += m_fg & (unsigned)p_fg;
        }
      }
    }
  }
  const unsigned cnt[4] = {n_px - n_m - n_p + n_mp, n_p - n_mp, n_m - n_mp, n_mp};   // TN, FP, FN, TP
  __shared__ unsigned red[4][8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const unsigned t = warp_sum(cnt[k]);
    if (lane == 0) red[k][threadIdx.x >> 5] = t;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    unsigned long long t = 0;
    for (int wv = 0; wv < 8; ++wv) t += red[threadIdx.x][wv];
    if (t) atomicAdd(cm + (long long)n * 4 + threadIdx.x, t);
  }
}

}  // namespace b200ssl

extern "C" {

int b200ssl_argmax_channels(const float* x, int n_images, int n_channels, int64_t hw,
                            void* labels_out, int out_dtype, int32_t* nonzero_out,
                            b200ssl_stream_t stream) {
  using namespace b200ssl;
  B200SSL_REQUIRE(n_images >= 0 && hw >= 0 && n_channels >= 1, "argmax_channels: bad extents");
  B200SSL_REQUIRE(n_images <= 65535, "argmax_channels: too many images");
  B200SSL_REQUIRE(out_dtype == B200SSL_I64 || out_dtype == B200SSL_U8, "argmax_channels: out dtype must be int64 or uint8");
  B200SSL_REQUIRE(out_dtype != B200SSL_U8 || n_channels <= 256, "argmax_channels: uint8 output needs C <= 256");
  if (n_images == 0 || hw == 0) return 0;
  B200SSL_REQUIRE(x && labels_out, "argmax_channels: null argument");
  const bool vec = aligned16(x) && (hw % 4 == 0);
  long long bx = ((hw + 3) / 4 + 255) / 256;
  long long cap = (long long)kNumSMs * 8 / n_images;
  if (cap < 1) cap = 1;
  if (bx > cap) bx = cap;
  const dim3 grid((unsigned)bx, (unsigned)n_images);
  cudaStream_t s = (cudaStream_t)stream;
  prof_begin("argmax_channels", s);
  if (out_dtype == B200SSL_I64)
    argmax_channels_kernel<long long><<<grid, 256, 0, s>>>(x, n_channels, hw, static_cast<long long*>(labels_out), nonzero_out, vec);
  else
    argmax_channels_kernel<unsigned char><<<grid, 256, 0, s>>>(x, n_channels, hw, static_cast<unsigned char*>(labels_out), nonzero_out, vec);
  return check_launch("argmax_channels");
}

size_t b200ssl_dice_workspace_bytes(int n, int64_t chw) {
  if (n <= 0 || chw <= 0) return 0;
  return (size_t)n * b200ssl::dice_blocks(n, chw) * 2 * sizeof(double);
}

int b200ssl_dice_metric(const float* input, const float* target, int n, int64_t chw,
                        float* dice_out, void* workspace, size_t workspace_bytes,
                        b200ssl_stream_t stream) {
  using namespace b200ssl;
  B200SSL_REQUIRE(n >= 0 && chw >= 0, "dice_metric: bad extents");
  B200SSL_REQUIRE(n <= 65535, "dice_metric: too many samples");
  if (n == 0) return 0;
  B200SSL_REQUIRE(input && target && dice_out, "dice_metric: null argument");
  const int bx = dice_blocks(n, chw);
  const size_t need = (size_t)n * bx * 2 * sizeof(double);
  if (!workspace || workspace_bytes < need) {
    set_error("dice_metric: workspace too small (%zu < %zu)", workspace_bytes, need);
    return B200SSL_EWORKSPACE;
  }
  const bool vec = aligned16(input) && aligned16(target) && (chw % 4 == 0);
  cudaStream_t s = (cudaStream_t)stream;
  prof_begin("dice_partial", s);
  dice_partial_kernel<<<dim3((unsigned)bx, (unsigned)n), kDiceThreads, 0, s>>>(
      input, target, chw, static_cast<double*>(workspace), vec);
  int rc = check_launch("dice partial");
  if (rc) return rc;
  prof_begin("dice_final", s);
  dice_final_kernel<<<(n + 127) / 128, 128, 0, s>>>(static_cast<const double*>(workspace), bx, n, dice_out);
  return check_launch("dice final");
}

int b200ssl_validation_cm(const float* logits, int n, int n_channels, int h, int w, const float* mask,
                          int mask_channels, int H, int W, float threshold, int fg_class, long long* cm_per_image,
                          b200ssl_stream_t stream) {
  using namespace b200ssl;
  B200SSL_REQUIRE(n >= 0 && n_channels >= 1 && h >= 1 && w >= 1 && H >= 0 && W >= 0 && mask_channels >= 1,
                  "validation_cm: bad extents");
  B200SSL_REQUIRE(fg_class >= 0 && fg_class < mask_channels, "validation_cm: class %d not in the mask", fg_class);
  B200SSL_REQUIRE(n <= 65535, "validation_cm: too many samples");
  if (n == 0 || H == 0 || W == 0) return 0;
  B200SSL_REQUIRE(logits && mask && cm_per_image, "validation_cm: null argument");
  long long bx = (H + 7) / 8;                              // 8 rows (warps) per block and trip
  long long cap = (long long)kNumSMs * 8 / n;
  if (cap < 1) cap = 1;
  if (bx > cap) bx = cap;
  const bool vec = aligned16(mask) && (W % 4 == 0);
  prof_begin("validation_cm", (cudaStream_t)stream);
  validation_cm_kernel<<<dim3((unsigned)bx, (unsigned)n), 256, 0, (cudaStream_t)stream>>>(
      logits, n_channels, h, w, mask, mask_channels, H, W, threshold, fg_class, vec,
      reinterpret_cast<unsigned long long*>(cm_per_image));
  return check_launch("validation_cm");
}

}  // extern "C"
