/*
 * oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, single-threaded restatement of the arithmetic of the reference's semi-supervised
 * loss-and-mixing path, used ONLY as the checker in tests/, __graft_entry__.smoke() and the
 * cpu_baseline leg of bench.py.  Nothing under semi-supervised_semantic_segmentation_b200/ may
 * link, load or call this file.
 *
 * Every function cites the reference lines (under /root/reference) whose arithmetic it follows.
 * Where the reference delegates to ATen (torch 2.11.0 is the effective pin) the
 * published behaviour of the ATen op is restated: IEEE fp32 ops in the order the Python
 * expression issues them.  Built with -ffp-contract=off so that only the explicit fmaf() calls
 * fuse.
 *
 * Pinning: the reference ships no tests or golden vectors ("parity unpinned" by the reference
 * itself); this oracle is pinned against outputs of the reference's own functions executed in
 * the build container (tests/golden/make_golden.py -> tests/golden/*.npz, checked by
 * tests/test_oracle_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * CowMix separable smoothing, cowmix.py:27-37 (two depthwise conv2d, zero padding K/2):
 *   V[n,y,x] = sum_i taps[n,i] * noise[n, y+i-k, x];   S[n,y,x] = sum_j taps[n,j] * V[n,y,x+j-k]
 * Accumulation order is the one the CUDA kernel documents: taps ascending, one fused
 * multiply-add per tap, accumulator starting at +0; out-of-image taps are skipped (adding
 * w*0 to a finite accumulator does not change it).
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_cowmix_field(const float* noise, const float* taps, int K, int n, int h, int w,
                              float* V, float* S) {
  const int k = K / 2;
  for (int s = 0; s < n; ++s) {
    const float* wt = taps + (size_t)s * K;
    const float* in = noise + (size_t)s * h * w;
    float* v = V + (size_t)s * h * w;
    float* o = S + (size_t)s * h * w;
    for (int y = 0; y < h; ++y) {
      float* row = v + (size_t)y * w;
      for (int x = 0; x < w; ++x) row[x] = 0.0f;
      for (int i = 0; i < K; ++i) {
        const int yy = y + i - k;
        if (yy < 0 || yy >= h) continue;
        const float* src = in + (size_t)yy * w;
        const float wi = wt[i];
        for (int x = 0; x < w; ++x) row[x] = fmaf(src[x], wi, row[x]);
      }
    }
    for (int y = 0; y < h; ++y) {
      const float* src = v + (size_t)y * w;
      float* row = o + (size_t)y * w;
      for (int x = 0; x < w; ++x) {
        float acc = 0.0f;
        for (int j = 0; j < K; ++j) {
          const int xx = x + j - k;
          if (xx < 0 || xx >= w) continue;
          acc = fmaf(src[xx], wt[j], acc);
        }
        row[x] = acc;
      }
    }
  }
}

/* cowmix.py:60-66: per-sample mean and UNBIASED std over (1,2,3), tau = factor*std + mean.
 * Statistics in double (ATen's CPU mean/std accumulate in double for float inputs via Welford /
 * cascade sums; the result is rounded to fp32 once), threshold arithmetic in fp32. */
ORC_API void orc_cowmix_tau(const float* S, const float* factor, int n, long long plane,
                            float* tau, float* mean_out, float* std_out) {
  for (int s = 0; s < n; ++s) {
    const float* p = S + (size_t)s * plane;
    double s1 = 0.0;
    for (long long i = 0; i < plane; ++i) s1 += (double)p[i];
    const double mean = s1 / (double)plane;
    double s2 = 0.0;
    for (long long i = 0; i < plane; ++i) {
      const double d = (double)p[i] - mean;
      s2 += d * d;
    }
    const double var = s2 / ((double)plane - 1.0);
    const float stdf = (float)sqrt(var);
    const float meanf = (float)mean;
    if (mean_out) mean_out[s] = meanf;
    if (std_out) std_out[s] = stdf;
    tau[s] = factor[s] * stdf + meanf; /* fp32 mul then fp32 add (cowmix.py:66) */
  }
}

/* cowmix.py:68: mask = (S > tau).to(fp32) */
ORC_API void orc_cowmix_threshold(const float* S, const float* tau, int n, long long plane,
                                  float* mask) {
  for (int s = 0; s < n; ++s)
    for (long long i = 0; i < plane; ++i)
      mask[(size_t)s * plane + i] = S[(size_t)s * plane + i] > tau[s] ? 1.0f : 0.0f;
}

/* cowmix.py:72-73: tensor_a * mask + tensor_b * (1. - mask); four separately rounded fp32 ops.
 * mask is [n,1,hw] (mask_channels == 1) or [n,c,hw]. */
ORC_API void orc_mix(const float* a, const float* b, const float* mask, int mask_channels,
                     long long n, int c, long long hw, float* out) {
  for (long long s = 0; s < n; ++s)
    for (int ch = 0; ch < c; ++ch)
      for (long long i = 0; i < hw; ++i) {
        const size_t idx = ((size_t)s * c + ch) * hw + i;
        const float m = mask_channels == 1 ? mask[(size_t)s * hw + i] : mask[idx];
        const float om = 1.0f - m;
        const float t0 = a[idx] * m;
        const float t1 = b[idx] * om;
        out[idx] = t0 + t1;
      }
}

/* ------------------------------------------------------------------------------------------
 * Lovasz: one segment (one class of one image group), lovasz.py:187-200 with lovasz_grad
 * lovasz.py:19-31.
 *   pred[i], fg[i] (0/1) for the L VALID pixels of the segment (ignored pixels already dropped,
 *   lovasz.py:217-219).  errors = |fg - pred|; stable descending sort (ties by ascending index --
 *   torch.sort leaves tie order unspecified); jaccard deltas from integer counts with
 *   one IEEE divide, `1 - q`, and the adjacent difference, all fp32.
 * Outputs: loss (the dot product, accumulated in double), grad[i] = dLoss_seg/dpred[i]
 *   = delta[rank(i)] * sign(pred - fg)   (sign(0) = 0, torch's abs backward), and optionally
 *   delta_sorted / order for diagnostics.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  float err;
  int32_t idx;
} orc_key;

static int orc_key_cmp(const void* pa, const void* pb) {
  const orc_key* a = (const orc_key*)pa;
  const orc_key* b = (const orc_key*)pb;
  if (a->err > b->err) return -1;
  if (a->err < b->err) return 1;
  return (a->idx > b->idx) - (a->idx < b->idx);
}

ORC_API double orc_lovasz_segment(const float* pred, const uint8_t* fg, long long L, float* grad,
                                  float* delta_sorted, int32_t* order) {
  if (L <= 0) return 0.0;
  orc_key* keys = (orc_key*)malloc((size_t)L * sizeof(orc_key));
  long long G = 0;
  for (long long i = 0; i < L; ++i) {
    const float f = fg[i] ? 1.0f : 0.0f;
    keys[i].err = fabsf(f - pred[i]);
    keys[i].idx = (int32_t)i;
    G += fg[i] ? 1 : 0;
  }
  qsort(keys, (size_t)L, sizeof(orc_key), orc_key_cmp);
  const float gts = (float)G;
  long long cfg = 0, cbg = 0;
  float jprev = 0.0f;
  double loss = 0.0;
  for (long long k = 0; k < L; ++k) {
    const int32_t i = keys[k].idx;
    if (fg[i]) ++cfg; else ++cbg;
    const float inter = gts - (float)cfg;   /* gts - cumsum(gt)      lovasz.py:26 */
    const float uni = gts + (float)cbg;     /* gts + cumsum(1 - gt)  lovasz.py:27 */
    const float q = inter / uni;
    const float j = 1.0f - q;               /* lovasz.py:28 */
    const float d = (k == 0) ? j : (j - jprev); /* lovasz.py:29-30 */
    jprev = j;
    loss += (double)keys[k].err * (double)d;    /* lovasz.py:200 */
    const float diff = (fg[i] ? 1.0f : 0.0f) - pred[i];
    float g = 0.0f;
    if (diff > 0.0f) g = -d;
    else if (diff < 0.0f) g = d;
    if (grad) grad[i] = g;
    if (delta_sorted) delta_sorted[k] = d;
    if (order) order[k] = i;
  }
  free(keys);
  return loss;
}

/* ------------------------------------------------------------------------------------------
 * lovasz_hinge_flat, lovasz.py:96-111, on the L valid pixels of one image (or of the batch):
 *   signs = 2*label - 1; errors = 1 - logits*signs (the product is exact, ONE rounding);
 *   descending sort of the SIGNED errors (all of them, as the reference does; ties by ascending index);
 *   loss = dot(relu(errors_sorted), lovasz_grad(gt_sorted)).
 * grad[i] = dLoss/dlogit[i] = -sign_i * delta[rank(i)] where error_i > 0, else 0 (relu's backward).
 * ------------------------------------------------------------------------------------------ */
ORC_API double orc_lovasz_hinge_segment(const float* logits, const uint8_t* fg, long long L, float* grad) {
  if (L <= 0) return 0.0;
  orc_key* keys = (orc_key*)malloc((size_t)L * sizeof(orc_key));
  long long G = 0;
  for (long long i = 0; i < L; ++i) {
    const float sx = fg[i] ? logits[i] : -logits[i];
    keys[i].err = 1.0f - sx;
    keys[i].idx = (int32_t)i;
    G += fg[i] ? 1 : 0;
  }
  qsort(keys, (size_t)L, sizeof(orc_key), orc_key_cmp);
  const float gts = (float)G;
  long long cfg = 0, cbg = 0;
  float jprev = 0.0f;
  double loss = 0.0;
  for (long long k = 0; k < L; ++k) {
    const int32_t i = keys[k].idx;
    if (fg[i]) ++cfg; else ++cbg;
    const float j = 1.0f - (gts - (float)cfg) / (gts + (float)cbg);
    const float d = (k == 0) ? j : (j - jprev);
    jprev = j;
    const float e = keys[k].err;
    if (e > 0.0f) {
      loss += (double)e * (double)d;
      if (grad) grad[i] = fg[i] ? -d : d;
    } else if (grad) {
      grad[i] = 0.0f;
    }
  }
  free(keys);
  return loss;
}

/* lovasz.py:19-31 on an already sorted 0/1 vector (known-answer tests) */
ORC_API void orc_lovasz_grad(const uint8_t* gt_sorted, long long p, float* out) {
  long long G = 0;
  for (long long i = 0; i < p; ++i) G += gt_sorted[i] ? 1 : 0;
  const float gts = (float)G;
  long long cfg = 0, cbg = 0;
  float jprev = 0.0f;
  for (long long k = 0; k < p; ++k) {
    if (gt_sorted[k]) ++cfg; else ++cbg;
    const float j = 1.0f - (gts - (float)cfg) / (gts + (float)cbg);
    out[k] = (k == 0) ? j : (j - jprev);
    jprev = j;
  }
}

/* ------------------------------------------------------------------------------------------
 * mean_teacher.py:10-11: ema.mul_(alpha).add_(param, alpha=1.-alpha)
 *   t = RN(e * f32(alpha));  e' = fmaf(p, f32(1.0 - alpha evaluated in double), t)
 * (ATen's add with alpha is `a + alpha*b` evaluated as one fused multiply-add on both the
 * vectorised CPU path and the CUDA path; SURVEY 3.5 [probed].)
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_ema(float* ema, const float* param, long long n, double alpha) {
  const float a = (float)alpha;
  const float b = (float)(1.0 - alpha);
  for (long long i = 0; i < n; ++i) {
    const float t = ema[i] * a;
    ema[i] = fmaf(param[i], b, t);
  }
}

/* ------------------------------------------------------------------------------------------
 * Row N2: F.interpolate(x, size, mode='bilinear', align_corners=False) (train.py:72-75,93-94,
 * losses.py:18-19) = ATen upsample_bilinear2d:
 *   s = max(fmaf(in/out, dst+0.5, -0.5), 0); i0 = min(floor(s), in-1); l = clamp(s-i0, 0, 1)
 *   v = fmaf(1-ly, fmaf(1-lx, v00, lx*v01), ly*fmaf(1-lx, v10, lx*v11))
 * [probed: bit-identical to torch 2.11 CPU for out >= in, tests/golden/upsample.npz; down-sampling
 * takes another ATen code path and agrees to 1 ulp only].
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_upsample_bilinear(const float* in, long long planes, int h, int w, float* out, int H, int W) {
  const float ry = (float)h / (float)H, rx = (float)w / (float)W;
  for (long long p = 0; p < planes; ++p) {
    const float* src = in + (size_t)p * h * w;
    float* dst = out + (size_t)p * H * W;
    for (int y = 0; y < H; ++y) {
      float s = fmaf(ry, (float)y + 0.5f, -0.5f);
      if (s < 0.f) s = 0.f;
      int y0 = (int)floorf(s);
      if (y0 > h - 1) y0 = h - 1;
      float ly = s - (float)y0;
      ly = ly < 0.f ? 0.f : (ly > 1.f ? 1.f : ly);
      const float hy = 1.0f - ly;
      const int y1 = y0 + (y0 < h - 1 ? 1 : 0);
      for (int x = 0; x < W; ++x) {
        float t = fmaf(rx, (float)x + 0.5f, -0.5f);
        if (t < 0.f) t = 0.f;
        int x0 = (int)floorf(t);
        if (x0 > w - 1) x0 = w - 1;
        float lx = t - (float)x0;
        lx = lx < 0.f ? 0.f : (lx > 1.f ? 1.f : lx);
        const float hx = 1.0f - lx;
        const int x1 = x0 + (x0 < w - 1 ? 1 : 0);
        const float t0 = fmaf(hx, src[(size_t)y0 * w + x0], lx * src[(size_t)y0 * w + x1]);
        const float t1 = fmaf(hx, src[(size_t)y1 * w + x0], lx * src[(size_t)y1 * w + x1]);
        dst[(size_t)y * W + x] = fmaf(hy, t0, ly * t1);
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * Row N2, student side: the backward of F.interpolate(.., 'bilinear', align_corners=False) (autograd's
 * upsample_bilinear2d_backward behind losses.py:18-19) as the deterministic gather the CUDA kernel
 * computes: for a low-resolution pixel (iy, ix) the output pixels whose taps touch it are visited in
 * ascending (y, x) order, acc = fma(RN(wy * wx), g, acc), wy = [i0(y) == iy] * (1 - ly) + [i1(y) == iy] * ly.
 * ATen adds the same products (CPU: in output order per tap; CUDA: atomically), so this agrees with
 * autograd to rounding (<= 1e-5 relative in L2, tests/golden/lowres_lovasz.npz), not bit for bit.
 * ------------------------------------------------------------------------------------------ */
static void orc_axis_tap(int dst, int in, float scale, int* i0, int* i1, float* w0, float* w1) {
  float s = fmaf(scale, (float)dst + 0.5f, -0.5f);
  if (s < 0.f) s = 0.f;
  int a = (int)floorf(s);
  if (a > in - 1) a = in - 1;
  float l = s - (float)a;
  l = l < 0.f ? 0.f : (l > 1.f ? 1.f : l);
  *i0 = a;
  *i1 = a + (a < in - 1 ? 1 : 0);
  *w0 = 1.0f - l;
  *w1 = l;
}

ORC_API void orc_upsample_bilinear_backward(const float* gfull, long long planes, int H, int W, float* glow, int h,
                                            int w) {
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  for (long long p = 0; p < planes; ++p) {
    const float* g = gfull + (size_t)p * H * W;
    float* o = glow + (size_t)p * h * w;
    for (int iy = 0; iy < h; ++iy)
      for (int ix = 0; ix < w; ++ix) {
        /* generous integer windows: every pixel outside the true support has weight 0 and is skipped */
        long long y_lo = (long long)(iy - 1) * H / h - 2, y_hi = (long long)(iy + 2) * H / h + 2;
        long long x_lo = (long long)(ix - 1) * W / w - 2, x_hi = (long long)(ix + 2) * W / w + 2;
        if (y_lo < 0) y_lo = 0;
        if (x_lo < 0) x_lo = 0;
        if (y_hi > H - 1) y_hi = H - 1;
        if (x_hi > W - 1) x_hi = W - 1;
        float acc = 0.f;
        for (long long y = y_lo; y <= y_hi; ++y) {
          int i0, i1;
          float w0, w1;
          orc_axis_tap((int)y, h, sy, &i0, &i1, &w0, &w1);
          const float wy = (i0 == iy ? w0 : 0.f) + (i1 == iy ? w1 : 0.f);
          if (wy == 0.f) continue;
          for (long long x = x_lo; x <= x_hi; ++x) {
            orc_axis_tap((int)x, w, sx, &i0, &i1, &w0, &w1);
            const float wx = (i0 == ix ? w0 : 0.f) + (i1 == ix ? w1 : 0.f);
            if (wx == 0.f) continue;
            acc = fmaf(wy * wx, g[(size_t)y * W + x], acc);
          }
        }
        o[(size_t)iy * w + ix] = acc;
      }
  }
}

/* ------------------------------------------------------------------------------------------
 * Row N4: train.py:122-124,130 -- clip_grad_norm_ -> torch.optim.SGD.step -> EMA, per element:
 *   g = RN(g*coef) (only when clipping); g = fmaf(p, wd, g) (wd != 0);
 *   b = first ? g : fmaf(g, 1-damp, RN(b*mu)); d = nesterov ? fmaf(b, mu, g) : b  (mu != 0);
 *   p = fmaf(d, -lr, p);  e = fmaf(p, 1-alpha, RN(e*alpha)) (alpha >= 0).
 * torch/optim/sgd.py _single_tensor_sgd / _multi_tensor_sgd; every add(alpha=) is one fma [probed
 * bit for bit against torch.optim.SGD, foreach on and off, tests/golden/sgd.npz].
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_sgd_ema(float* p, const float* g, float* buf, float* ema, long long n, int use_coef,
                         float coef, double lr, double mu, double damp, double wd, double alpha,
                         int nesterov, int first) {
  const float neg_lr = (float)(-lr), muf = (float)mu, omd = (float)(1.0 - damp), wdf = (float)wd;
  const float a = (float)alpha, b1 = (float)(1.0 - alpha);
  for (long long i = 0; i < n; ++i) {
    float gi = use_coef ? g[i] * coef : g[i];
    if (wd != 0.0) gi = fmaf(p[i], wdf, gi);
    float d = gi;
    if (mu != 0.0) {
      float b = first ? gi : fmaf(gi, omd, buf[i] * muf);
      buf[i] = b;
      d = nesterov ? fmaf(b, muf, gi) : b;
    }
    p[i] = fmaf(d, neg_lr, p[i]);
    if (alpha >= 0.0 && ema) {
      const float t = ema[i] * a;
      ema[i] = fmaf(p[i], b1, t);
    }
  }
}

/* sum of squares in double, sequential (the "exact" value the norm kernels are compared with) */
ORC_API double orc_sqnorm(const float* g, long long n) {
  double acc = 0.0;
  for (long long i = 0; i < n; ++i) acc += (double)g[i] * (double)g[i];
  return acc;
}

/* torch.nn.utils.clip_grad_norm_: clip_coef = max_norm / (total_norm + 1e-6) with a tensor total_norm,
 * i.e. reciprocal(total_norm + 1e-6) * max_norm in fp32, clamped to <= 1 [probed]. */
ORC_API float orc_clip_coef(float total_norm, double max_norm) {
  const float t = total_norm + 1e-6f;
  const float r = 1.0f / t;
  const float c = r * (float)max_norm;
  return c < 1.0f ? c : (c != c ? c : 1.0f);
}

/* ------------------------------------------------------------------------------------------
 * Confusion matrix (restated oracle): bincount(label*D + pred) over pixels whose
 * label != ignore.  other_bucket: D = C+1 and out-of-range values go to row/column C; otherwise
 * D = C and pixels with an out-of-range label or prediction are dropped (and counted).
 * ------------------------------------------------------------------------------------------ */
ORC_API long long orc_confusion(const int64_t* labels, const int64_t* preds, long long n, int C,
                                int other_bucket, int has_ignore, int64_t ignore, int64_t* cm) {
  const int D = C + (other_bucket ? 1 : 0);
  long long dropped = 0;
  for (long long i = 0; i < n; ++i) {
    int64_t l = labels[i], p = preds[i];
    if (has_ignore && l == ignore) continue;
    const int lo = (l < 0 || l >= C), po = (p < 0 || p >= C);
    if (lo || po) {
      if (!other_bucket) { ++dropped; continue; }
      if (lo) l = C;
      if (po) p = C;
    }
    cm[l * D + p] += 1;
  }
  return dropped;
}

/* torch.argmax(x, dim=1) for x [n, C, hw]: first maximum wins, NaN is the maximum (losses.py:240) */
ORC_API void orc_argmax_channels(const float* x, long long n, int C, long long hw, int64_t* out) {
  for (long long s = 0; s < n; ++s)
    for (long long i = 0; i < hw; ++i) {
      int arg = 0;
      float best = x[((size_t)s * C) * hw + i];
      for (int c = 1; c < C; ++c) {
        const float v = x[((size_t)s * C + c) * hw + i];
        if (v > best || (v != v && best == best)) { best = v; arg = c; }
      }
      out[(size_t)s * hw + i] = arg;
    }
}

/* metrics.py:1-7: (2*sum(x*y) + 1) / (sum(x+y) + 1) per sample; sums in double then fp32 */
ORC_API void orc_dice(const float* x, const float* y, int n, long long chw, float* out) {
  for (int s = 0; s < n; ++s) {
    double si = 0.0, sc = 0.0;
    for (long long i = 0; i < chw; ++i) {
      si += (double)(x[(size_t)s * chw + i] * y[(size_t)s * chw + i]);
      sc += (double)(x[(size_t)s * chw + i] + y[(size_t)s * chw + i]);
    }
    const float inter = (float)si, card = (float)sc;
    out[s] = (2.0f * inter + 1.0f) / (card + 1.0f);
  }
}
