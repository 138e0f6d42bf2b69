// Device side of the NVLink peer exchange (see peer.cu): shared with the kernels that post their own
// results (the last Lovasz pass posts [confusion matrix || loss] from its finalising block).
#pragma once
#include "common.cuh"

namespace b200ssl {

constexpr int kPeerMaxRanks = B200SSL_PEER_MAX_RANKS;
constexpr int kPeerDepth = 4;
constexpr int kPeerMaxWords = B200SSL_PEER_MAX_WORDS;
constexpr int kPeerMaxFloats = B200SSL_PEER_MAX_FLOATS;
constexpr size_t kAckOffset = (size_t)kPeerDepth * kPeerMaxRanks * kPeerMaxWords;  // in 8-byte words
constexpr size_t kStatusOffset = kAckOffset + kPeerMaxRanks;
constexpr size_t kMailWords = kStatusOffset + 16;

struct PeerDev {  // by-value kernel parameter
  unsigned long long* mail[kPeerMaxRanks];
  int rank, world;
  unsigned seq;
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

// Block-cooperative: every thread of the (single) block calls it.
__device__ __forceinline__ void peer_post_block(const PeerDev& c, const long long* __restrict__ ints, int n_ints,
                                const PeerFloats& f, int n_floats) {
  unsigned long long* me = c.mail[c.rank];
  // flow control: the slot of step seq was last used by step seq-depth
  if ((int)threadIdx.x < c.world && c.seq > (unsigned)kPeerDepth) {
    const unsigned need = c.seq - (unsigned)kPeerDepth;
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
  const int slot = (int)(c.seq % (unsigned)kPeerDepth);
  const int nw = 2 * n_ints + n_floats;
  for (int w = threadIdx.x; w < nw; w += blockDim.x) {
    unsigned data;
    if (w < 2 * n_ints) {
      const unsigned long long v = (unsigned long long)ints[w >> 1];
      data = (w & 1) ? (unsigned)(v >> 32) : (unsigned)v;
    } else {
      data = __float_as_uint(*f.p[w - 2 * n_ints]);
    }
    const unsigned long long word = ((unsigned long long)c.seq << 32) | data;
    const size_t at = ll_index(slot, c.rank, w);
    for (int r = 0; r < c.world; ++r) st_sys(c.mail[(c.rank + r) % c.world] + at, word);
  }
}


// what a producing kernel needs to post on behalf of b200ssl_peer_post (filled by peer_begin_post)
struct PeerTail {
  PeerDev dev;
  const long long* ints;
  int n_ints;
  int enabled;
};

int peer_begin_post(::b200ssl_peer_comm* c, int n_ints, int n_floats, PeerDev* out);          // peer.cu
int peer_post_impl(::b200ssl_peer_comm* c, const long long* ints, int n_ints, const float* const* floats_host,
                   int n_floats, cudaStream_t s);                                             // peer.cu
int binary_lovasz_fused_impl(const float* scores, const float* target, int n_images, int n_channels, int64_t hw,
                             int cls, const float* grad_out, unsigned char* labels_out, int32_t* nonzero,
                             float* loss_out, float* denom_out, float* seg_loss, int32_t* seg_fg,
                             int32_t* seg_valid, float* grad, long long* cm, int cm_has_ignore,
                             int64_t cm_ignore_index, void* workspace, size_t workspace_bytes,
                             b200ssl_stream_t stream, const PeerTail* tail);                   // lovasz.cu

}  // namespace b200ssl
