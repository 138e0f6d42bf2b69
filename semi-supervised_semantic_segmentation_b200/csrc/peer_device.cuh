// Device side of the NVLink peer exchange (see peer.cu): shared with the kernels that post their own
// results (the Lovasz finalising block posts [confusion matrix || loss] and collects the previous step).
#pragma once
#include "common.cuh"

namespace b200ssl {

constexpr int kPeerMaxRanks = B200SSL_PEER_MAX_RANKS;
constexpr int kPeerDepth = 4;
constexpr int kPeerMaxWords = B200SSL_PEER_MAX_WORDS;
constexpr int kPeerMaxFloats = B200SSL_PEER_MAX_FLOATS;
constexpr size_t kAckOffset = (size_t)kPeerDepth * kPeerMaxRanks * kPeerMaxWords;  // in 8-byte words
constexpr size_t kStatusOffset = kAckOffset + kPeerMaxRanks;   // sticky time-out flag
constexpr size_t kSeqOffset = kStatusOffset + 1;               // last step this rank posted     (local only)
constexpr size_t kCollectedOffset = kStatusOffset + 2;         // last step this rank collected  (local only)
constexpr size_t kMailWords = kStatusOffset + 16;

// The sequence numbers live in DEVICE memory (the local mailbox), not in kernel arguments: a launch
// carries no per-step state, so the whole step -- post and collect included -- can be captured once in a
// CUDA graph and replayed, and a failed host call can never leave host and device counters out of step.
struct PeerDev {  // by-value kernel parameter
  unsigned long long* mail[kPeerMaxRanks];
  int rank, world;
  unsigned long long timeout_ns;
};

struct PeerFloats {
  const float* p[kPeerMaxFloats];
};

__host__ __device__ inline size_t ll_index(int slot, int src, int w) {
  return ((size_t)slot * kPeerMaxRanks + src) * kPeerMaxWords + w;
}

__device__ __forceinline__ unsigned long long ld_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Block-cooperative (every thread of the single block calls it): publishes the next step.  Returns its
// sequence number to all threads.
__device__ __forceinline__ unsigned peer_post_block(const PeerDev& c, const long long* __restrict__ ints, int n_ints,
                                                    const PeerFloats& f, int n_floats) {
  __shared__ unsigned s_seq;
  unsigned long long* me = c.mail[c.rank];
  if (threadIdx.x == 0) s_seq = (unsigned)ld_sys(me + kSeqOffset) + 1u;
  __syncthreads();
  const unsigned seq = s_seq;
  // flow control: the slot of step seq was last used by step seq-depth
  if ((int)threadIdx.x < c.world && seq > (unsigned)kPeerDepth) {
    const unsigned need = seq - (unsigned)kPeerDepth;
    const unsigned long long* a = me + kAckOffset + threadIdx.x;
    const unsigned long long t0 = global_ns();
    while ((int)((unsigned)ld_sys(a) - need) < 0) {
      if (global_ns() - t0 > c.timeout_ns) {
        atomicExch(me + kStatusOffset, 1ull);
        break;
      }
      __nanosleep(200);
    }
  }
  __syncthreads();
  const int slot = (int)(seq % (unsigned)kPeerDepth);
  const int nw = 2 * n_ints + n_floats;
  for (int w = threadIdx.x; w < nw; w += blockDim.x) {
    unsigned data;
    if (w < 2 * n_ints) {
      const unsigned long long v = (unsigned long long)ints[w >> 1];
      data = (w & 1) ? (unsigned)(v >> 32) : (unsigned)v;
    } else {
      data = __float_as_uint(*f.p[w - 2 * n_ints]);
    }
    const unsigned long long word = ((unsigned long long)seq << 32) | data;
    const size_t at = ll_index(slot, c.rank, w);
    for (int r = 0; r < c.world; ++r) st_sys(c.mail[(c.rank + r) % c.world] + at, word);
  }
  if (threadIdx.x == 0) st_sys(me + kSeqOffset, (unsigned long long)seq);
  return seq;
}

__device__ __forceinline__ unsigned peer_poll(const PeerDev& c, unsigned seq, const unsigned long long* p,
                                              unsigned long long* status, bool* dead) {
  unsigned long long v = ld_sys(p);
  if ((unsigned)(v >> 32) == seq) return (unsigned)v;
  if (*dead) return 0u;
  const unsigned long long t0 = global_ns();
  for (;;) {
    v = ld_sys(p);
    if ((unsigned)(v >> 32) == seq) return (unsigned)v;
    if (global_ns() - t0 > c.timeout_ns) {
      atomicExch(status, 2ull);
      *dead = true;
      return 0u;
    }
    __nanosleep(100);
  }
}

// Block-cooperative: sums the world rows of step `seq` in the LOCAL mailbox in rank order (int64 adds for
// counts, fp64 adds for scalars: identical bits on every rank) unless that step has been collected
// already, then acknowledges the slot to every peer.  seq == 0: nothing has been posted yet.
// Every thread polls up to kPollUnroll (word, rank) pairs per round and ISSUES all of its sys-scope loads before it
// looks at the first tag, so blockDim.x * kPollUnroll loads are in flight together (a thread walking all ranks of
// an item one poll after the other costs 2*world dependent ~1 us round trips: 16 at 8 GPUs; one pair per thread
// still cost 28 rounds for the 883 words x 8 ranks of a 21-class matrix: +57 us per step at N = 8).  The sums then
// read the staged words from shared memory in rank order.
constexpr int kPollUnroll = 8;
__device__ __forceinline__ void peer_collect_block(const PeerDev& c, unsigned seq, int n_ints, int n_floats,
                                                   long long* __restrict__ ints_out, double* __restrict__ floats_out) {
  __shared__ unsigned s_done;
  __shared__ unsigned s_words[256 * kPollUnroll];
  unsigned long long* me = c.mail[c.rank];
  unsigned long long* status = me + kStatusOffset;
  if (threadIdx.x == 0) s_done = (unsigned)ld_sys(me + kCollectedOffset);
  __syncthreads();
  if (seq == 0u || (int)(s_done - seq) >= 0) return;   // block-uniform
  const int slot = (int)(seq % (unsigned)kPeerDepth);
  const int world = c.world;
  const int nthreads = min((int)blockDim.x, 256);
  const int nw = 2 * n_ints + n_floats;                 // words per rank
  const int per_chunk = ((nthreads * kPollUnroll) / world) & ~1;   // words per chunk (even: an int64 never straddles)
  bool dead = false;
  const int t = (int)threadIdx.x;
  for (int w0 = 0; w0 < nw; w0 += per_chunk) {
    const int words_here = min(per_chunk, nw - w0);
    const int pairs_here = words_here * world;           // staged as [word][rank]
    if (t < nthreads) {
      unsigned long long v[kPollUnroll];
#pragma unroll
      for (int u = 0; u < kPollUnroll; ++u) {
        const int idx = t + u * nthreads;
        if (idx < pairs_here) v[u] = ld_sys(me + ll_index(slot, idx % world, w0 + idx / world));
      }
#pragma unroll
      for (int u = 0; u < kPollUnroll; ++u) {
        const int idx = t + u * nthreads;
        if (idx < pairs_here)
          s_words[idx] = ((unsigned)(v[u] >> 32) == seq)
                             ? (unsigned)v[u]
                             : peer_poll(c, seq, me + ll_index(slot, idx % world, w0 + idx / world), status, &dead);
      }
    }
    __syncthreads();
    // one thread per item of this chunk: rank order, identical result on every rank
    for (int k = t; k < words_here; k += (int)blockDim.x) {
      const int w = w0 + k;
      if (w < 2 * n_ints) {
        if ((w & 1) == 0) {
          long long acc = 0;
          for (int r = 0; r < world; ++r)
            acc += (long long)(((unsigned long long)s_words[(k + 1) * world + r] << 32) | s_words[k * world + r]);
          ints_out[w >> 1] = acc;
        }
      } else {
        double acc = 0.0;
        for (int r = 0; r < world; ++r) acc += (double)__uint_as_float(s_words[k * world + r]);
        floats_out[w - 2 * n_ints] = acc;
      }
    }
    __syncthreads();
  }
  // every word of this slot has been consumed: tell the peers they may reuse it
  if ((int)threadIdx.x < world) st_sys(c.mail[threadIdx.x] + kAckOffset + c.rank, (unsigned long long)seq);
  if (threadIdx.x == 0) st_sys(me + kCollectedOffset, (unsigned long long)seq);
}

// What a producing kernel needs to run the exchange by itself (filled by peer_tail): it posts this step and,
// when prev_* are given, collects the PREVIOUS step in the same block -- by then every peer has long posted
// it, so the step costs no extra launch, event or stream for the collective and never waits unless a peer
// is a whole step behind.
struct PeerTail {
  PeerDev dev;
  const long long* ints;
  int n_ints;
  int enabled;
  long long* prev_ints_out;
  double* prev_floats_out;
};

int peer_tail(::b200ssl_peer_comm* c, int n_ints, int n_floats, PeerTail* out);               // peer.cu
int peer_post_impl(::b200ssl_peer_comm* c, const long long* ints, int n_ints, const float* const* floats_host,
                   int n_floats, long long* prev_ints_out, double* prev_floats_out, cudaStream_t s);  // peer.cu
int binary_lovasz_fused_impl(const float* scores, const float* target, int n_images, int n_channels, int64_t hw,
                             int cls, const float* grad_out, unsigned char* labels_out, int32_t* nonzero,
                             float* loss_out, float* denom_out, float* seg_loss, int32_t* seg_fg,
                             int32_t* seg_valid, float* grad, long long* cm, int cm_has_ignore,
                             int64_t cm_ignore_index, void* workspace, size_t workspace_bytes,
                             b200ssl_stream_t stream, const PeerTail* tail);                   // lovasz.cu

int lovasz_forward_backward_tail(const b200ssl_lovasz_desc* d, const float* probas, const void* labels,
                                 const float* grad_out, const int32_t* binary_nonzero, float* loss_out,
                                 float* denom_out, float* seg_loss, int32_t* seg_fg, int32_t* seg_valid,
                                 float* grad_probas, void* workspace, size_t workspace_bytes, b200ssl_stream_t stream,
                                 const PeerTail* tail);                                        // lovasz.cu

}  // namespace b200ssl
