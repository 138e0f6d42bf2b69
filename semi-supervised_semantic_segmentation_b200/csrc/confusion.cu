// Confusion matrix cm[label*C + pred] as a shared-memory-privatised histogram, plus the fused
// argmax-from-logits variant and Dice from per-image 2x2 matrices.
//
// No reference function computes a confusion matrix; the quantities derived from it
// are metrics.dice_metric (metrics.py:1-7) and lovasz.iou (lovasz.py:54-73).  The restated oracle
// is bincount(label*C + pred, minlength=C*C) over the non-ignored pixels.
//
// Segmentation labels are spatially coherent, so most of a warp hits the same bin.  A per-pixel
// shared-memory atomic would serialise (32 cycles per warp on one bank); instead each warp finds
// the runs of equal bins among adjacent lanes with one shuffle + one ballot and issues ONE atomic
// per run.  Each warp owns a private histogram copy when C*C is small enough, so there is no
// inter-warp contention either.  Counts are exact integers: results are bit-identical to the oracle.
//
// Roofline: HBM-bound at 16 B/pixel with int64 label+pred (85 % of the measured peak).  At 2 B/pixel (uint8) the
// histogram itself is the bound: confusion_u8_kernel below stitches runs across the lanes of a warp and issues one
// shared-memory atomic per run (57 % of the HBM peak on segmentation-like masks); the generic kernel serves uint8
// only for unaligned inputs and for matrices beyond the shared-memory budget.
#include "common.cuh"

namespace b200ssl {

constexpr int kCmThreads = 256;
constexpr int kCmWarps = kCmThreads / 32;
constexpr int kCmSmemBudget = 64 * 1024;

template <typename T>
__device__ __forceinline__ int make_bin(T l, T p, int C, int D, bool has_ignore, long long ignore,
                                        unsigned& dropped) {
  long long ll = (long long)l, pp = (long long)p;
  if (has_ignore && ll == ignore) return -1;
  const bool lo = (ll < 0 || ll >= C), po = (pp < 0 || pp >= C);
  if (lo || po) {
    if (D == C) {  // no "other" bucket: drop and count
      ++dropped;
      return -1;
    }
    if (lo) ll = C;
    if (po) pp = C;
  }
  return (int)ll * D + (int)pp;
}

// same for labels that fit 32 bits (uint8 / int32 inputs): no 64-bit compares on the hot path.
// ignore32 is the ignore index if it is representable, else a value no label can take.
__device__ __forceinline__ int make_bin32(int l, int p, int C, int D, bool has_ignore, int ignore32,
                                          unsigned& dropped) {
  if (has_ignore && l == ignore32) return -1;
  const bool lo = ((unsigned)l >= (unsigned)C), po = ((unsigned)p >= (unsigned)C);
  if (lo || po) {
    if (D == C) {
      ++dropped;
      return -1;
    }
    if (lo) l = C;
    if (po) p = C;
  }
  return l * D + p;
}

// flush a block's private copies into the global int64 matrix
__device__ __forceinline__ void flush_hist(const unsigned* hist, int copies, int bins,
                                           unsigned long long* cm) {
  for (int b = threadIdx.x; b < bins; b += blockDim.x) {
    unsigned long long s = 0;
    for (int c = 0; c < copies; ++c) s += hist[c * bins + b];
    if (s) atomicAdd(cm + b, s);
  }
}

// T = label/pred element type, ELEMS = elements per 16-byte vector (1 => scalar loads)
template <typename T, int ELEMS>
__global__ void __launch_bounds__(kCmThreads)
confusion_kernel(const T* __restrict__ labels, const T* __restrict__ preds, long long plane_pixels,
                 int C, int D, bool has_ignore, long long ignore, int copies,
                 unsigned long long* __restrict__ cm, long long cm_plane_stride,
                 unsigned long long* __restrict__ dropped_out) {
  extern __shared__ unsigned smem_hist[];
  const int bins = D * D;
  const bool use_smem = copies > 0;
  labels += (long long)blockIdx.y * plane_pixels;
  preds += (long long)blockIdx.y * plane_pixels;
  cm += (long long)blockIdx.y * cm_plane_stride;
  if (use_smem) {
    for (int i = threadIdx.x; i < copies * bins; i += blockDim.x) smem_hist[i] = 0;
    __syncthreads();
  }
  const int warp = threadIdx.x >> 5;
  // a warp's histogram: its private copy, a shared copy, or the global matrix itself (32-bit view
  // is not possible there, so the global fallback uses 64-bit atomics below)
  unsigned* my_hist = use_smem ? smem_hist + (warp % copies) * bins : nullptr;

  union Vec { uint4 u; T e[ELEMS > 1 ? ELEMS : 1]; };
  constexpr int kPerGroup = 32 * ELEMS;
  const long long n_groups = (plane_pixels + kPerGroup - 1) / kPerGroup;
  // contiguous range of warp-groups per block
  const long long per_block = (n_groups + gridDim.x - 1) / gridDim.x;
  const long long g_begin = (long long)blockIdx.x * per_block;
  const long long g_end = min(n_groups, g_begin + per_block);
  unsigned dropped = 0;
  constexpr int kUnroll = 4;
  for (long long g0 = g_begin + warp; g0 < g_end; g0 += (long long)kCmWarps * kUnroll) {
    Vec lv[kUnroll], pv[kUnroll];
    bool full[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long g = g0 + (long long)u * kCmWarps;
      const long long first = (g * 32 + lane_id()) * ELEMS;
      full[u] = (g < g_end) && (first + ELEMS <= plane_pixels);
      if (full[u]) {
        if (ELEMS > 1) {
          lv[u].u = __ldg(reinterpret_cast<const uint4*>(labels + first));
          pv[u].u = __ldg(reinterpret_cast<const uint4*>(preds + first));
        } else {
          lv[u].e[0] = __ldg(labels + first);
          pv[u].e[0] = __ldg(preds + first);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long g = g0 + (long long)u * kCmWarps;
      if (g >= g_end) break;  // warp-uniform
      const long long first = (g * 32 + lane_id()) * ELEMS;
      // Packed labels (4 x int32 / 16 x uint8 per lane): a lane whose ELEMS labels and ELEMS
      // predictions are all equal -- the normal case inside a segment -- computes ONE bin and adds it
      // with weight ELEMS through the warp-merged path; only lanes that straddle a boundary (or the
      // ragged tail) fall back to one bin and one atomic per element.
      if (ELEMS >= 4 && use_smem) {
        bool lane_uniform = false;
        if (full[u]) {
          unsigned l0 = lv[u].u.x, p0 = pv[u].u.x;
          if (sizeof(T) == 1) { l0 = (l0 & 0xffu) * 0x01010101u; p0 = (p0 & 0xffu) * 0x01010101u; }
          lane_uniform = lv[u].u.x == l0 && lv[u].u.y == l0 && lv[u].u.z == l0 && lv[u].u.w == l0 &&
                         pv[u].u.x == p0 && pv[u].u.y == p0 && pv[u].u.z == p0 && pv[u].u.w == p0;
        }
        int b = -1;
        if (lane_uniform) {
          unsigned d1 = 0;
          b = make_bin<T>(lv[u].e[0], pv[u].e[0], C, D, has_ignore, ignore, d1);
          dropped += d1 * ELEMS;
        }
        warp_run_add_weighted(my_hist, b, (unsigned)ELEMS);
        if (!lane_uniform) {
          // boundary lane (or ragged tail): run-length merge inside the lane, 32-bit arithmetic
          const bool ign_fits = ignore >= -2147483647ll && ignore <= 2147483647ll;
          const int ignore32 = ign_fits ? (int)ignore : (int)0x80000000;
          const bool has_ign32 = has_ignore && ign_fits;
          int cur = -1;
          unsigned cnt = 0;
          auto merge = [&](int be, unsigned n) {
            if (be == cur) {
              cnt += n;
            } else {
              if (cur >= 0) atomicAdd(my_hist + cur, cnt);
              cur = be;
              cnt = n;
            }
          };
          if (full[u]) {
            // word by word: a uint8 word holds 4 pixels that mostly agree even in a boundary lane
            const unsigned lw[4] = {lv[u].u.x, lv[u].u.y, lv[u].u.z, lv[u].u.w};
            const unsigned pw[4] = {pv[u].u.x, pv[u].u.y, pv[u].u.z, pv[u].u.w};
#pragma unroll
            for (int w4 = 0; w4 < 4; ++w4) {
              if (sizeof(T) == 1) {
                const unsigned l0 = lw[w4] & 0xffu, p0 = pw[w4] & 0xffu;
                if (lw[w4] == l0 * 0x01010101u && pw[w4] == p0 * 0x01010101u) {
                  unsigned d1 = 0;
                  const int be = make_bin32((int)l0, (int)p0, C, D, has_ign32, ignore32, d1);
                  dropped += 4u * d1;
                  merge(be, 4u);
                } else {
#pragma unroll
                  for (int e = 0; e < 4; ++e)
                    merge(make_bin32((int)((lw[w4] >> (8 * e)) & 0xffu), (int)((pw[w4] >> (8 * e)) & 0xffu), C, D,
                                     has_ign32, ignore32, dropped), 1u);
                }
              } else {
                merge(make_bin32((int)lw[w4], (int)pw[w4], C, D, has_ign32, ignore32, dropped), 1u);
              }
            }
          } else {
            for (int e = 0; e < ELEMS; ++e)
              if (first + e < plane_pixels)
                merge(make_bin32((int)labels[first + e], (int)preds[first + e], C, D, has_ign32, ignore32, dropped), 1u);
          }
          if (cur >= 0) atomicAdd(my_hist + cur, cnt);
        }
        continue;
      }
      int bin[ELEMS];
#pragma unroll
      for (int e = 0; e < ELEMS; ++e) {
        bin[e] = -1;
        if (full[u]) {
          bin[e] = make_bin<T>(lv[u].e[e], pv[u].e[e], C, D, has_ignore, ignore, dropped);
        } else if (first + e < plane_pixels) {  // ragged tail of the plane: scalar loads
          bin[e] = make_bin<T>(labels[first + e], preds[first + e], C, D, has_ignore, ignore, dropped);
        }
      }
      if (use_smem) {
        // the ELEMS pixels of a lane almost always share a bin: one warp-merged atomic per lane group
        lane_run_add<ELEMS>(my_hist, bin);
      } else {
        // very large C: the matrix does not fit in shared memory, merge runs then go to L2
#pragma unroll
        for (int e = 0; e < ELEMS; ++e) {
          const unsigned lane = lane_id();
          const int prev = __shfl_up_sync(0xffffffffu, bin[e], 1);
          const bool head = (lane == 0) || (bin[e] != prev);
          const unsigned heads = __ballot_sync(0xffffffffu, head);
          if (head && bin[e] >= 0) {
            const unsigned later = (lane == 31) ? 0u : (heads & (0xffffffffu << (lane + 1)));
            const int end = later ? (__ffs(later) - 1) : 32;
            atomicAdd(cm + bin[e], (unsigned long long)(end - (int)lane));
          }
        }
      }
    }
  }
  if (dropped_out) {
    const unsigned d = warp_sum(dropped);
    if (lane_id() == 0 && d) atomicAdd(dropped_out, (unsigned long long)d);
  }
  if (use_smem) {
    __syncthreads();
    flush_hist(smem_hist, copies, bins, cm);
  }
}

// ---- uint8 labels: 16 pixels per lane, runs merged ACROSS the lanes of a warp ------------------------
// A warp step covers 512 consecutive pixels; on segmentation masks that is one to a few dozen runs of equal
// (label, pred) pairs, and the kernel is bound by the instructions it issues per pixel (ncu r2_u: 71 % issue
// utilisation with the first version of this kernel), then by shared-memory atomics (about 2 LSU cycles per
// active lane), not by HBM.  So: every lane finds the boundaries inside its 16 pixels with byte arithmetic
// (pixel i differs from pixel i-1; one dot-product instruction per word packs the flags into a 16-bit mask),
// which gives its first run (key kf, length cf) and its last run (kt, ct); a lane without a boundary is one run
// of 16.  Runs are then stitched across lanes with two ballots and one shuffle: the lane in which a run STARTS
// adds the whole run -- its own tail, 16 for every following boundary-free lane that continues it, and the head
// of the lane the run ends in -- with ONE atomic.  Runs that start and end inside a lane (two or more
// boundaries in 16 pixels) are added by that lane directly.  Exact integer counts, bit-identical to the generic
// kernel.  A key is the byte quadruple [label, pred, pred, pred] (one PRMT to build, two to compare/decode).
__device__ __forceinline__ unsigned nonzero_bytes_x128(unsigned d) {   // byte i = (byte i of d != 0) ? 0x80 : 0
  return (((d & 0x7f7f7f7fu) + 0x7f7f7f7fu) | d) & 0x80808080u;
}

__device__ __forceinline__ unsigned byte_of(const unsigned (&w)[4], int pixel) {
  const unsigned lo = pixel < 4 ? w[0] : w[1], hi = pixel < 12 ? w[2] : w[3];
  return ((pixel < 8 ? lo : hi) >> ((pixel & 3) * 8)) & 0xffu;
}

__device__ __forceinline__ void smem_red_add(unsigned addr, unsigned v) {
  asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// SIMPLE: no other bucket, no dropped counter and the void label (if any) is not a class index, so a pixel is
// counted iff max(label, pred) < C.  n_vec = plane_pixels / 16 (< 2^31, checked by the launcher).
template <int UNROLL, bool SIMPLE>
__global__ void __launch_bounds__(kCmThreads)
confusion_u8_kernel(const unsigned char* __restrict__ labels, const unsigned char* __restrict__ preds,
                    long long plane_pixels, int n_vec, int C, int D, bool has_ignore, int ignore32, int copies,
                    unsigned long long* __restrict__ cm, long long cm_plane_stride,
                    unsigned long long* __restrict__ dropped_out) {
  extern __shared__ unsigned smem_hist[];
  const int bins = D * D;
  labels += (long long)blockIdx.y * plane_pixels;
  preds += (long long)blockIdx.y * plane_pixels;
  cm += (long long)blockIdx.y * cm_plane_stride;
  for (int i = threadIdx.x; i < copies * bins; i += blockDim.x) smem_hist[i] = 0;
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  const int lane = (int)lane_id();
  const unsigned hist_addr = (unsigned)__cvta_generic_to_shared(smem_hist + (warp % copies) * bins);
  unsigned dropped = 0;
  auto add = [&](unsigned key, unsigned count) {
    const unsigned l = key & 0xffu, p = key >> 24;
    if (SIMPLE) {
      if (max(l, p) < (unsigned)C) smem_red_add(hist_addr + (l * D + p) * 4u, count);
    } else {
      unsigned d1 = 0;
      const int b = make_bin32((int)l, (int)p, C, D, has_ignore, ignore32, d1);
      dropped += d1 * count;
      if (b >= 0) smem_red_add(hist_addr + (unsigned)b * 4u, count);
    }
  };

  const int n_groups = (int)((plane_pixels + 511) >> 9);
  const int per_block = (n_groups + (int)gridDim.x - 1) / (int)gridDim.x;
  const int g_begin = (int)min((long long)blockIdx.x * per_block, (long long)n_groups);
  const int g_end = min(n_groups, g_begin + per_block);
  const uint4* __restrict__ lvec = reinterpret_cast<const uint4*>(labels);
  const uint4* __restrict__ pvec = reinterpret_cast<const uint4*>(preds);
  for (int g0 = g_begin + warp; g0 < g_end; g0 += kCmWarps * UNROLL) {
    uint4 lv[UNROLL], pv[UNROLL];
    bool full[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int g = g0 + u * kCmWarps;
      const int v = g * 32 + lane;
      full[u] = (g < g_end) && (v < n_vec);
      if (full[u]) {
        lv[u] = ld_stream_u4(lvec + v);
        pv[u] = ld_stream_u4(pvec + v);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int g = g0 + u * kCmWarps;
      if (g >= g_end) break;  // warp-uniform
      // a lane without a full vector: keys that link to nothing, and it is no continuation of anything
      unsigned kf = 0x00ff0100u, kt = 0x00ff0200u, cf = 0, ct = 0;
      bool uni = false;
      if (full[u]) {
        const unsigned lw[4] = {lv[u].x, lv[u].y, lv[u].z, lv[u].w};
        const unsigned pw[4] = {pv[u].x, pv[u].y, pv[u].z, pv[u].w};
        // byte i of bnd[w] != 0  <=>  pixel 4w+i differs from pixel 4w+i-1 (pixel 0: never)
        unsigned bnd[4];
        bnd[0] = (lw[0] ^ __byte_perm(lw[0], 0u, 0x2100)) | (pw[0] ^ __byte_perm(pw[0], 0u, 0x2100));
#pragma unroll
        for (int w = 1; w < 4; ++w)
          bnd[w] = (lw[w] ^ __byte_perm(lw[w - 1], lw[w], 0x6543)) | (pw[w] ^ __byte_perm(pw[w - 1], pw[w], 0x6543));
        const unsigned lo = __dp4a(nonzero_bytes_x128(bnd[0]), 0x08040201u,
                                   __dp4a(nonzero_bytes_x128(bnd[1]), 0x80402010u, 0u));
        const unsigned hi = __dp4a(nonzero_bytes_x128(bnd[2]), 0x08040201u,
                                   __dp4a(nonzero_bytes_x128(bnd[3]), 0x80402010u, 0u));
        const unsigned b16 = (lo + (hi << 8)) >> 7;             // bit i: a run starts at pixel i (i >= 1)
        kf = __byte_perm(lw[0], pw[0], 0x4440);
        kt = __byte_perm(lw[3], pw[3], 0x7773);
        uni = b16 == 0u;
        cf = (unsigned)(__ffs(b16 | 0x10000u) - 1);             // length of the first run (16: no boundary)
        ct = (unsigned)(__clz(b16 | 1u) - 15);                  // 16 - position of the last boundary
        unsigned inner = b16 & (b16 - 1u);
        int pos = (int)cf;
        while (inner) {                                         // runs that begin and end inside this lane
          const int next = __ffs(inner) - 1;
          add(byte_of(lw, pos) | (byte_of(pw, pos) << 24), (unsigned)(next - pos));
          pos = next;
          inner &= inner - 1u;
        }
      }
      const unsigned prev_t = __shfl_up_sync(0xffffffffu, kt, 1);
      const bool linked = lane > 0 && kf == prev_t;                  // my first run continues the previous lane's last
      const unsigned U = __ballot_sync(0xffffffffu, uni), L = __ballot_sync(0xffffffffu, linked);
      const unsigned above = lane == 31 ? 0u : (0xffffffffu << (lane + 1));
      const unsigned stop = ~(L & U) & above;                        // first later lane that is not a pure continuation
      const int e = stop ? (__ffs(stop) - 1) : 32;
      const unsigned cf_e = __shfl_sync(0xffffffffu, cf, e & 31);
      if (full[u]) {
        if (!uni || !linked) {                                       // my last run starts in this lane
          unsigned total = ct + 16u * (unsigned)(e - lane - 1);
          if (e < 32 && ((L >> e) & 1u)) total += cf_e;              // ... and ends inside lane e
          add(kt, total);
        }
        if (!uni && !linked) add(kf, cf);                            // a first run nobody else owns
      } else if (g * 32 + lane == n_vec) {                           // ragged tail of the plane
        const long long first = (long long)n_vec * 16;
        for (int i = 0; first + i < plane_pixels; ++i)
          add((unsigned)labels[first + i] | ((unsigned)preds[first + i] << 24), 1u);
      }
    }
  }
  if (!SIMPLE && dropped_out) {
    const unsigned d = warp_sum(dropped);
    if (lane == 0 && d) atomicAdd(dropped_out, (unsigned long long)d);
  }
  __syncthreads();
  flush_hist(smem_hist, copies, bins, cm);
}

// ---- fused argmax + confusion matrix from logits -----------------------------------------------
// logits [n_images, C, hw] fp32; each lane owns 4 consecutive pixels (float4 per channel).
template <typename T>
__global__ void __launch_bounds__(kCmThreads)
confusion_logits_kernel(const float* __restrict__ logits, const T* __restrict__ labels,
                        long long hw, int C, int D, bool has_ignore, long long ignore, int copies,
                        bool vec_ok, unsigned long long* __restrict__ cm,
                        long long cm_plane_stride, unsigned long long* __restrict__ dropped_out) {
  extern __shared__ unsigned smem_hist[];
  const int bins = D * D;
  const bool use_smem = copies > 0;
  logits += (long long)blockIdx.y * C * hw;
  labels += (long long)blockIdx.y * hw;
  cm += (long long)blockIdx.y * cm_plane_stride;
  if (use_smem) {
    for (int i = threadIdx.x; i < copies * bins; i += blockDim.x) smem_hist[i] = 0;
    __syncthreads();
  }
  const int warp = threadIdx.x >> 5;
  unsigned* my_hist = use_smem ? smem_hist + (warp % copies) * bins : nullptr;
  constexpr int kPerGroup = 32 * 4;
  const long long n_groups = (hw + kPerGroup - 1) / kPerGroup;
  const long long per_block = (n_groups + gridDim.x - 1) / gridDim.x;
  const long long g_begin = (long long)blockIdx.x * per_block;
  const long long g_end = min(n_groups, g_begin + per_block);
  unsigned dropped = 0;
  for (long long g = g_begin + warp; g < g_end; g += kCmWarps) {
    const long long first = (g * 32 + lane_id()) * 4;
    float best[4];
    int arg[4];
    const bool full = vec_ok && (first + 4 <= hw);
#pragma unroll
    for (int e = 0; e < 4; ++e) { best[e] = 0.f; arg[e] = -1; }
    if (full) {
      // torch.argmax: first maximal element wins; NaN counts as the maximum
      for (int c0 = 0; c0 < C; c0 += 4) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (c0 + u < C) v[u] = ld_stream_f4(logits + (long long)(c0 + u) * hw + first);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (c0 + u < C) {
            const float x[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const bool take = (arg[e] < 0) || argmax_beats(x[e], best[e]);
              if (take) { best[e] = x[e]; arg[e] = c0 + u; }
            }
          }
        }
      }
    } else {
      for (int e = 0; e < 4; ++e) {
        if (first + e < hw) {
          for (int c = 0; c < C; ++c) {
            const float x = logits[(long long)c * hw + first + e];
            const bool take = (arg[e] < 0) || argmax_beats(x, best[e]);
            if (take) { best[e] = x; arg[e] = c; }
          }
        }
      }
    }
    int bin[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      bin[e] = -1;
      if (first + e < hw) {
        bin[e] = make_bin<long long>((long long)labels[first + e], (long long)arg[e], C, D, has_ignore,
                                     ignore, dropped);
      }
      if (!use_smem && bin[e] >= 0) atomicAdd(cm + bin[e], 1ull);
    }
    if (use_smem) lane_run_add<4>(my_hist, bin);
  }
  if (dropped_out) {
    const unsigned d = warp_sum(dropped);
    if (lane_id() == 0 && d) atomicAdd(dropped_out, (unsigned long long)d);
  }
  if (use_smem) {
    __syncthreads();
    flush_hist(smem_hist, copies, bins, cm);
  }
}

// metrics.py:1-7 on {0,1} inputs: intersection = TP, cardinality = 2TP + FP + FN (fp32 sums of
// integers); dice = (2*I + 1) / (card + 1), each op rounded to fp32 like the ATen chain.
__global__ void dice_from_cm_kernel(const long long* __restrict__ cm, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long fp = cm[4 * i + 1], fn = cm[4 * i + 2], tp = cm[4 * i + 3];
  const float inter = __ll2float_rn(tp);
  const float card = __ll2float_rn(2 * tp + fp + fn);
  out[i] = __fdiv_rn(__fadd_rn(__fmul_rn(2.0f, inter), 1.0f), __fadd_rn(card, 1.0f));
}

static int cm_copies(int D) {
  const long long bytes = (long long)D * D * 4;
  long long copies = kCmSmemBudget / bytes;
  if (copies > kCmWarps) copies = kCmWarps;
  return (int)copies;  // 0 => global atomics
}

template <typename T, int ELEMS>
static int launch_confusion(const void* labels, const void* preds, long long plane_pixels,
                            int planes, int C, int D, bool has_ignore, long long ignore,
                            unsigned long long* cm, long long cm_stride,
                            unsigned long long* dropped, cudaStream_t s) {
  const int copies = cm_copies(D);
  const size_t smem = (size_t)copies * D * D * 4;
  const long long groups = (plane_pixels + 32 * ELEMS - 1) / (32 * ELEMS);
  // enough work per block to amortise the histogram flush; never more than 8 CTAs per SM in total
  long long bx = (groups + kCmWarps * 4 - 1) / (kCmWarps * 4);
  long long cap = (long long)kNumSMs * 8 / planes;
  if (cap < 1) cap = 1;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  auto kern = confusion_kernel<T, ELEMS>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  prof_begin("confusion_matrix", s);
  kern<<<dim3((unsigned)bx, (unsigned)planes), kCmThreads, smem, s>>>(
      static_cast<const T*>(labels), static_cast<const T*>(preds), plane_pixels, C, D, has_ignore,
      ignore, copies, cm, cm_stride, dropped);
  return check_launch("confusion_matrix");
}

constexpr int kCm8Unroll = 2;                     // 1: 2-5 % slower (r2_u), 4 in the first version: no gain
constexpr long long kCm8MaxPlane = 1ll << 34;    // 32-bit vector indices inside confusion_u8_kernel

static int launch_confusion_u8(const void* labels, const void* preds, long long plane_pixels, int planes, int C,
                               int D, bool has_ignore, long long ignore, int copies, unsigned long long* cm,
                               long long cm_stride, unsigned long long* dropped, cudaStream_t s) {
  const size_t smem = (size_t)copies * D * D * 4;
  const bool ign_fits = ignore >= 0 && ignore <= 255;               // no uint8 label equals any other value
  const bool ign = has_ignore && ign_fits;
  const bool simple = D == C && dropped == nullptr && (!ign || ignore >= C);
  auto kern = simple ? confusion_u8_kernel<kCm8Unroll, true> : confusion_u8_kernel<kCm8Unroll, false>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  // one wave: as many CTAs as fit at once, each with one contiguous range of the plane
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kCmThreads, smem) != cudaSuccess || per_sm < 1)
    per_sm = 1;
  const long long groups = (plane_pixels + 511) / 512;
  long long bx = (groups + kCmWarps * kCm8Unroll - 1) / (kCmWarps * kCm8Unroll);
  long long cap = (long long)kNumSMs * per_sm / planes;
  if (cap < 1) cap = 1;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  prof_begin("confusion_matrix", s);
  kern<<<dim3((unsigned)bx, (unsigned)planes), kCmThreads, smem, s>>>(
      static_cast<const unsigned char*>(labels), static_cast<const unsigned char*>(preds), plane_pixels,
      (int)(plane_pixels >> 4), C, D, ign, ign ? (int)ignore : -1, copies, cm, cm_stride, dropped);
  return check_launch("confusion_matrix");
}

}  // namespace b200ssl

extern "C" {

int b200ssl_confusion_matrix(const void* labels, const void* preds, int64_t n_pixels,
                             int num_classes, int other_bucket, int has_ignore,
                             int64_t ignore_index, int label_dtype, int per_image, int64_t hw,
                             long long* cm, long long* dropped, b200ssl_stream_t stream) {
  using namespace b200ssl;
  B200SSL_REQUIRE(n_pixels >= 0, "confusion_matrix: negative pixel count");
  B200SSL_REQUIRE(num_classes >= 1 && num_classes <= 4096, "confusion_matrix: num_classes out of range");
  B200SSL_REQUIRE(cm != nullptr, "confusion_matrix: null cm");
  if (n_pixels == 0) return 0;
  B200SSL_REQUIRE(labels && preds, "confusion_matrix: null input");
  const int D = num_classes + (other_bucket ? 1 : 0);
  long long plane = n_pixels;
  int planes = 1;
  long long stride = 0;
  if (per_image) {
    B200SSL_REQUIRE(hw > 0 && n_pixels % hw == 0, "confusion_matrix: per_image needs n_pixels %% hw == 0");
    B200SSL_REQUIRE(n_pixels / hw <= 65535, "confusion_matrix: too many images for per_image mode");
    plane = hw;
    planes = (int)(n_pixels / hw);
    stride = (long long)D * D;
  }
  cudaStream_t s = (cudaStream_t)stream;
  unsigned long long* ucm = reinterpret_cast<unsigned long long*>(cm);
  unsigned long long* udrop = reinterpret_cast<unsigned long long*>(dropped);
  const bool has_ign = has_ignore != 0;
  size_t esz = label_dtype == B200SSL_I64 ? 8 : label_dtype == B200SSL_I32 ? 4 : 1;
  const bool vec = aligned16(labels) && aligned16(preds) && (planes == 1 || (plane * esz) % 16 == 0);
  switch (label_dtype) {
    case B200SSL_I64:
      return vec ? launch_confusion<long long, 2>(labels, preds, plane, planes, num_classes, D, has_ign, ignore_index, ucm, stride, udrop, s)
                 : launch_confusion<long long, 1>(labels, preds, plane, planes, num_classes, D, has_ign, ignore_index, ucm, stride, udrop, s);
    case B200SSL_I32:
      return vec ? launch_confusion<int, 4>(labels, preds, plane, planes, num_classes, D, has_ign, ignore_index, ucm, stride, udrop, s)
                 : launch_confusion<int, 1>(labels, preds, plane, planes, num_classes, D, has_ign, ignore_index, ucm, stride, udrop, s);
    case B200SSL_U8:
      if (vec && cm_copies(D) > 0 && plane <= kCm8MaxPlane)
        return launch_confusion_u8(labels, preds, plane, planes, num_classes, D, has_ign, ignore_index, cm_copies(D), ucm, stride, udrop, s);
      return vec ? launch_confusion<unsigned char, 16>(labels, preds, plane, planes, num_classes, D, has_ign, ignore_index, ucm, stride, udrop, s)
                 : launch_confusion<unsigned char, 1>(labels, preds, plane, planes, num_classes, D, has_ign, ignore_index, ucm, stride, udrop, s);
    default:
      set_error("confusion_matrix: unknown label dtype %d", label_dtype);
      return B200SSL_EINVAL;
  }
}

int b200ssl_confusion_from_logits(const float* logits, const void* labels, int n_images,
                                  int num_classes, int64_t hw, int has_ignore, int64_t ignore_index,
                                  int label_dtype, int per_image, long long* cm, long long* dropped,
                                  b200ssl_stream_t stream) {
  using namespace b200ssl;
  B200SSL_REQUIRE(n_images >= 0 && hw >= 0, "confusion_from_logits: negative extent");
  B200SSL_REQUIRE(num_classes >= 1 && num_classes <= 4096, "confusion_from_logits: num_classes out of range");
  B200SSL_REQUIRE(n_images <= 65535, "confusion_from_logits: too many images");
  B200SSL_REQUIRE(cm != nullptr, "confusion_from_logits: null cm");
  if (n_images == 0 || hw == 0) return 0;
  B200SSL_REQUIRE(logits && labels, "confusion_from_logits: null input");
  const int copies = cm_copies(num_classes);
  const int D = num_classes;
  const size_t smem = (size_t)copies * num_classes * num_classes * 4;
  const long long groups = (hw + 127) / 128;
  long long bx = (groups + kCmWarps * 4 - 1) / (kCmWarps * 4);
  long long cap = (long long)kNumSMs * 8 / n_images;
  if (cap < 1) cap = 1;
  if (bx > cap) bx = cap;
  const bool vec_ok = aligned16(logits) && (hw % 4 == 0);
  const long long stride = per_image ? (long long)num_classes * num_classes : 0;
  cudaStream_t s = (cudaStream_t)stream;
  unsigned long long* ucm = reinterpret_cast<unsigned long long*>(cm);
  unsigned long long* udrop = reinterpret_cast<unsigned long long*>(dropped);
#define LAUNCH_LOGITS(T)                                                                       \
  do {                                                                                         \
    auto kern = confusion_logits_kernel<T>;                                                    \
    if (smem > 48 * 1024)                                                                      \
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
    kern<<<dim3((unsigned)bx, (unsigned)n_images), kCmThreads, smem, s>>>(                     \
        logits, static_cast<const T*>(labels), hw, num_classes, D, has_ignore != 0, ignore_index, \
        copies, vec_ok, ucm, stride, udrop);                                                   \
  } while (0)
  prof_begin("confusion_from_logits", s);
  switch (label_dtype) {
    case B200SSL_I64: LAUNCH_LOGITS(long long); break;
    case B200SSL_I32: LAUNCH_LOGITS(int); break;
    case B200SSL_U8: LAUNCH_LOGITS(unsigned char); break;
    default:
      set_error("confusion_from_logits: unknown label dtype %d", label_dtype);
      return B200SSL_EINVAL;
  }
#undef LAUNCH_LOGITS
  return check_launch("confusion_from_logits");
}

int b200ssl_dice_from_cm(const long long* cm_per_image, int n_images, float* dice_out,
                         b200ssl_stream_t stream) {
  using namespace b200ssl;
  B200SSL_REQUIRE(n_images >= 0, "dice_from_cm: negative image count");
  if (n_images == 0) return 0;
  B200SSL_REQUIRE(cm_per_image && dice_out, "dice_from_cm: null argument");
  prof_begin("dice_from_cm", (cudaStream_t)stream);
  dice_from_cm_kernel<<<(n_images + 127) / 128, 128, 0, (cudaStream_t)stream>>>(cm_per_image, n_images, dice_out);
  return check_launch("dice_from_cm");
}

}  // extern "C"
