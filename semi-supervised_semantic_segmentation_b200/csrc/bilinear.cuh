// Bilinear interpolation taps shared by the fused up-sampling consumers (mix.cu: teacher predictions inside the
// mix; lovasz.cu: student logits inside the Lovasz front end, row N2 of SURVEY 8f).  Arithmetic = ATen's
// upsample_bilinear2d (align_corners=False), bit for bit on up-sampling and same-size shapes:
//   s  = max(fma(in/out, dst + 0.5, -0.5), 0);  i0 = min(floor(s), in-1);  l = clamp(s - i0, 0, 1)
//   v  = fma(1-ly, fma(1-lx, v00, RN(lx*v01)), RN(ly * fma(1-lx, v10, RN(lx*v11))))
#pragma once
#include "common.cuh"

namespace b200ssl {

struct AxisTap {
  int i0, i1;
  float w0, w1;
};
__device__ __forceinline__ AxisTap axis_tap(int dst, int in, float scale) {
  float s = __fmaf_rn(scale, (float)dst + 0.5f, -0.5f);
  s = s < 0.f ? 0.f : s;
  int i0 = (int)floorf(s);
  i0 = i0 > in - 1 ? in - 1 : i0;
  float l = __fsub_rn(s, (float)i0);
  l = l < 0.f ? 0.f : (l > 1.f ? 1.f : l);
  AxisTap t;
  t.i0 = i0;
  t.i1 = i0 + (i0 < in - 1 ? 1 : 0);
  t.w0 = __fsub_rn(1.0f, l);
  t.w1 = l;
  return t;
}
__device__ __forceinline__ float bilerp(const float* __restrict__ row0, const float* __restrict__ row1,
                                        const AxisTap& tx, float wy0, float wy1) {
  const float t0 = __fmaf_rn(tx.w0, __ldg(row0 + tx.i0), __fmul_rn(tx.w1, __ldg(row0 + tx.i1)));
  const float t1 = __fmaf_rn(tx.w0, __ldg(row1 + tx.i0), __fmul_rn(tx.w1, __ldg(row1 + tx.i1)));
  return __fmaf_rn(wy0, t0, __fmul_rn(wy1, t1));
}

}  // namespace b200ssl
