// Per-step all-reduce of the confusion matrix and the loss scalars over NVLink peer memory.
//
//   reference: utils/utils.py:43-54 reduce_tensor (dist.reduce per scalar, 4-6 times per step, each
//   followed by .item(): train.py:53,58,109,113,178,181); SURVEY 8(e): ONE all-reduce per step over
//   [C*C int64 || scalars], latency-bound, issued after the last kernel and consumed lazily.
//
// The payload is a few dozen bytes (40 B for configs[1], 3.5 KB for 21 classes), so the cost of a
// library collective is all launch + protocol latency and rank skew.  This file replaces it with a
// one-shot exchange through peer-mapped mailboxes (every GPU of an HGX box reaches every other one
// through NVSwitch):
//
//   post     (any stream, no waiting in steady state): every rank stores its words into ITS row of
//            EVERY rank's mailbox.  A word is 64 bit = (sequence number << 32 | 32 payload bits), so a
//            single 8-byte store publishes data and "ready" flag together -- no fence, no second
//            round trip (the LL idea).  int64 counts travel as two words, fp32 scalars as one.
//   collect  (the communicator's own stream, or the caller's): polls the world rows of the LOCAL
//            mailbox until every word carries this step's sequence number, adds them in rank order
//            (int64 adds for counts, fp64 adds for scalars: identical bits on every rank, exact
//            for counts), then acknowledges the slot to every peer.
//
// Mailbox rows are kPeerDepth deep (slot = seq % depth); a post for step s only reuses the slot of
// step s-depth after every peer has acknowledged collecting that step, so a rank may run up to
// depth-1 steps ahead of the slowest consumer without ever stalling its main stream.  Every spin is
// bounded by a timeout (default 20 s, B200SSL_PEER_TIMEOUT_MS) that raises a sticky status flag
// instead of hanging the GPU.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "peer_device.cuh"

namespace b200ssl {

__global__ void __launch_bounds__(256) peer_post_kernel(const __grid_constant__ PeerDev c,
                                                        const long long* __restrict__ ints, int n_ints,
                                                        const __grid_constant__ PeerFloats f, int n_floats) {
  peer_post_block(c, ints, n_ints, f, n_floats);
}

__device__ __forceinline__ unsigned peer_poll(const PeerDev& c, const unsigned long long* p, unsigned long long* status,
                                              bool* dead) {
  unsigned long long v = ld_sys(p);
  if ((unsigned)(v >> 32) == c.seq) return (unsigned)v;
  if (*dead) return 0u;
  const unsigned long long t0 = global_ns();
  for (;;) {
    v = ld_sys(p);
    if ((unsigned)(v >> 32) == c.seq) return (unsigned)v;
    if (global_ns() - t0 > c.timeout_ns) {
      atomicExch(status, 2ull);
      *dead = true;
      return 0u;
    }
    __nanosleep(100);
  }
}

__global__ void __launch_bounds__(256) peer_collect_kernel(const __grid_constant__ PeerDev c, int n_ints, int n_floats,
                                                           long long* __restrict__ ints_out,
                                                           double* __restrict__ floats_out) {
  unsigned long long* me = c.mail[c.rank];
  unsigned long long* status = me + kStatusOffset;
  const int slot = (int)(c.seq % (unsigned)kPeerDepth);
  bool dead = false;
  for (int item = threadIdx.x; item < n_ints + n_floats; item += blockDim.x) {
    if (item < n_ints) {
      long long acc = 0;
      for (int r = 0; r < c.world; ++r) {  // rank order: identical result on every rank
        const unsigned lo = peer_poll(c, me + ll_index(slot, r, 2 * item), status, &dead);
        const unsigned hi = peer_poll(c, me + ll_index(slot, r, 2 * item + 1), status, &dead);
        acc += (long long)(((unsigned long long)hi << 32) | lo);
      }
      ints_out[item] = acc;
    } else {
      double acc = 0.0;
      for (int r = 0; r < c.world; ++r)
        acc += (double)__uint_as_float(peer_poll(c, me + ll_index(slot, r, 2 * n_ints + (item - n_ints)), status, &dead));
      floats_out[item - n_ints] = acc;
    }
  }
  __syncthreads();
  // every word of this slot has been consumed: tell the peers they may reuse it
  if ((int)threadIdx.x < c.world) st_sys(c.mail[threadIdx.x] + kAckOffset + c.rank, (unsigned long long)c.seq);
}

}  // namespace b200ssl

struct b200ssl_peer_comm {
  int rank = 0, world = 1, dev = 0;
  unsigned long long* local = nullptr;
  unsigned long long* mail[b200ssl::kPeerMaxRanks] = {};
  bool opened[b200ssl::kPeerMaxRanks] = {};
  bool connected = false;
  unsigned seq = 0;          // last posted step
  unsigned collected = 0;    // last step a collect was issued for
  int n_ints = 0, n_floats = 0;
  unsigned long long timeout_ns = 20ull * 1000 * 1000 * 1000;
  cudaStream_t stream = nullptr;  // lazy collects run here
  cudaEvent_t posted = nullptr, done = nullptr;
};

namespace b200ssl {

static PeerDev device_view(const b200ssl_peer_comm* c, unsigned seq) {
  PeerDev d;
  for (int r = 0; r < kPeerMaxRanks; ++r) d.mail[r] = c->mail[r];
  d.rank = c->rank;
  d.world = c->world;
  d.seq = seq;
  d.timeout_ns = c->timeout_ns;
  return d;
}

// used by step.cu when the producing kernel posts by itself: validates, advances the sequence number and
// hands out the device view; the caller must make sure exactly one kernel calls peer_post_block with it
int peer_begin_post(b200ssl_peer_comm* c, int n_ints, int n_floats, PeerDev* out) {
  B200SSL_REQUIRE(c && c->connected && out, "peer_begin_post: communicator not connected");
  B200SSL_REQUIRE(n_ints >= 0 && n_floats >= 0 && n_floats <= kPeerMaxFloats && 2 * n_ints + n_floats <= kPeerMaxWords &&
                      n_ints + n_floats > 0, "peer_begin_post: payload does not fit");
  B200SSL_REQUIRE(c->collected == c->seq, "peer_begin_post: the previous post has not been collected yet");
  c->seq += 1;
  c->n_ints = n_ints;
  c->n_floats = n_floats;
  *out = device_view(c, c->seq);
  return 0;
}

// used by step.cu: the same post as b200ssl_peer_post (kept here so that a later fusion into the
// producing kernel only has to call peer_post_block)
int peer_post_impl(b200ssl_peer_comm* c, const long long* ints, int n_ints, const float* const* floats_host,
                   int n_floats, cudaStream_t s) {
  B200SSL_REQUIRE(c && c->connected, "peer_post: communicator not connected");
  B200SSL_REQUIRE(n_ints >= 0 && n_floats >= 0 && n_floats <= kPeerMaxFloats && 2 * n_ints + n_floats <= kPeerMaxWords &&
                      n_ints + n_floats > 0,
                  "peer_post: payload of %d counts + %d scalars does not fit (%d words, %d scalars max)", n_ints,
                  n_floats, kPeerMaxWords, kPeerMaxFloats);
  B200SSL_REQUIRE(n_ints == 0 || ints, "peer_post: null counts");
  B200SSL_REQUIRE(c->collected == c->seq, "peer_post: the previous post has not been collected yet");
  PeerFloats f = {};
  for (int i = 0; i < n_floats; ++i) {
    B200SSL_REQUIRE(floats_host && floats_host[i], "peer_post: null scalar pointer");
    f.p[i] = floats_host[i];
  }
  c->seq += 1;
  c->n_ints = n_ints;
  c->n_floats = n_floats;
  prof_begin("peer_post", s);
  peer_post_kernel<<<1, 256, 0, s>>>(device_view(c, c->seq), ints, n_ints, f, n_floats);
  return check_launch("peer_post");
}

}  // namespace b200ssl

extern "C" {

int b200ssl_peer_create(int rank, int world, b200ssl_peer_comm** comm_out, unsigned char* handle_out) {
  using namespace b200ssl;
  B200SSL_REQUIRE(comm_out != nullptr, "peer_create: null output");
  B200SSL_REQUIRE(world >= 1 && world <= kPeerMaxRanks && rank >= 0 && rank < world, "peer_create: bad rank/world %d/%d",
                  rank, world);
  static_assert(sizeof(cudaIpcMemHandle_t) <= B200SSL_PEER_HANDLE_BYTES, "handle size");
  b200ssl_peer_comm* c = new b200ssl_peer_comm();
  c->rank = rank;
  c->world = world;
  if (const char* t = getenv("B200SSL_PEER_TIMEOUT_MS")) {
    const long long ms = atoll(t);
    if (ms > 0) c->timeout_ns = (unsigned long long)ms * 1000000ull;
  }
  cudaError_t e = cudaGetDevice(&c->dev);
  // the mailbox must be exportable with cudaIpcGetMemHandle, which a caller-owned sub-allocation of a
  // caching allocator is not: this is the one device allocation the library makes itself (2 MB)
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&c->local), kMailWords * 8);
  if (e == cudaSuccess) e = cudaMemset(c->local, 0, kMailWords * 8);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  int lo = 0, hi = 0;
  if (e == cudaSuccess) e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
  if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, hi);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->posted, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming);
  if (e == cudaSuccess && handle_out) {
    cudaIpcMemHandle_t h;
    memset(handle_out, 0, B200SSL_PEER_HANDLE_BYTES);
    e = cudaIpcGetMemHandle(&h, c->local);
    if (e == cudaSuccess) memcpy(handle_out, &h, sizeof(h));
  }
  if (e != cudaSuccess) {
    set_error("peer_create: %s", cudaGetErrorString(e));
    cudaGetLastError();
    if (c->local) cudaFree(c->local);
    delete c;
    return (int)e;
  }
  c->mail[rank] = c->local;
  *comm_out = c;
  return 0;
}

void* b200ssl_peer_mailbox(b200ssl_peer_comm* c) { return c ? c->local : nullptr; }

int b200ssl_peer_connect(b200ssl_peer_comm* c, const unsigned char* handles) {
  using namespace b200ssl;
  B200SSL_REQUIRE(c && handles, "peer_connect: null argument");
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * B200SSL_PEER_HANDLE_BYTES, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      set_error("peer_connect: cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
      cudaGetLastError();
      return (int)e;
    }
    c->mail[r] = static_cast<unsigned long long*>(p);
    c->opened[r] = true;
  }
  c->connected = true;
  return 0;
}

int b200ssl_peer_connect_ptrs(b200ssl_peer_comm* c, void* const* mailboxes_host) {
  using namespace b200ssl;
  B200SSL_REQUIRE(c && mailboxes_host, "peer_connect_ptrs: null argument");
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) continue;
    B200SSL_REQUIRE(mailboxes_host[r] != nullptr, "peer_connect_ptrs: null mailbox for rank %d", r);
    c->mail[r] = static_cast<unsigned long long*>(mailboxes_host[r]);
  }
  c->connected = true;
  return 0;
}

int b200ssl_peer_post(b200ssl_peer_comm* c, const long long* ints, int n_ints, const float* const* floats_host,
                      int n_floats, b200ssl_stream_t stream) {
  return b200ssl::peer_post_impl(c, ints, n_ints, floats_host, n_floats, (cudaStream_t)stream);
}

int b200ssl_peer_collect(b200ssl_peer_comm* c, long long* ints_out, double* floats_out, b200ssl_stream_t post_stream,
                         b200ssl_stream_t stream) {
  using namespace b200ssl;
  B200SSL_REQUIRE(c && c->connected, "peer_collect: communicator not connected");
  B200SSL_REQUIRE(c->collected + 1 == c->seq, "peer_collect: nothing posted");
  B200SSL_REQUIRE((c->n_ints == 0 || ints_out) && (c->n_floats == 0 || floats_out), "peer_collect: null output");
  cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
  if (!stream) {
    // lazy: order the communicator's stream after the post, run there, never touch the caller's stream
    if (cudaEventRecord(c->posted, (cudaStream_t)post_stream) != cudaSuccess ||
        cudaStreamWaitEvent(c->stream, c->posted, 0) != cudaSuccess) {
      set_error("peer_collect: %s", cudaGetErrorString(cudaGetLastError()));
      return (int)cudaErrorUnknown;
    }
  }
  c->collected = c->seq;
  prof_begin("peer_collect", s);
  peer_collect_kernel<<<1, 256, 0, s>>>(device_view(c, c->seq), c->n_ints, c->n_floats, ints_out, floats_out);
  int rc = check_launch("peer_collect");
  if (rc == 0 && !stream && cudaEventRecord(c->done, c->stream) != cudaSuccess) {
    set_error("peer_collect: %s", cudaGetErrorString(cudaGetLastError()));
    return (int)cudaErrorUnknown;
  }
  return rc;
}

int b200ssl_peer_join(b200ssl_peer_comm* c, b200ssl_stream_t stream) {
  using namespace b200ssl;
  B200SSL_REQUIRE(c != nullptr, "peer_join: null communicator");
  if (c->collected == 0) return 0;
  if (cudaStreamWaitEvent((cudaStream_t)stream, c->done, 0) != cudaSuccess) {
    set_error("peer_join: %s", cudaGetErrorString(cudaGetLastError()));
    return (int)cudaErrorUnknown;
  }
  return 0;
}

int b200ssl_peer_allreduce(b200ssl_peer_comm* c, const long long* ints, int n_ints, const float* const* floats_host,
                           int n_floats, long long* ints_out, double* floats_out, b200ssl_stream_t stream) {
  int rc = b200ssl::peer_post_impl(c, ints, n_ints, floats_host, n_floats, (cudaStream_t)stream);
  if (rc) return rc;
  return b200ssl_peer_collect(c, ints_out, floats_out, stream, stream);
}

int b200ssl_peer_status(b200ssl_peer_comm* c) {
  using namespace b200ssl;
  B200SSL_REQUIRE(c && c->local, "peer_status: null communicator");
  unsigned long long st = 0;
  cudaError_t e = cudaMemcpy(&st, c->local + kStatusOffset, 8, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) {
    set_error("peer_status: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return (int)e;
  }
  if (st) set_error("peer: a %s timed out waiting for a peer (status %llu)", st == 1 ? "post" : "collect", st);
  return st ? B200SSL_ETIMEOUT : 0;
}

int b200ssl_peer_destroy(b200ssl_peer_comm* c) {
  if (!c) return 0;
  cudaStreamSynchronize(c->stream);
  for (int r = 0; r < c->world; ++r)
    if (c->opened[r]) cudaIpcCloseMemHandle(c->mail[r]);
  if (c->local) cudaFree(c->local);
  if (c->posted) cudaEventDestroy(c->posted);
  if (c->done) cudaEventDestroy(c->done);
  if (c->stream) cudaStreamDestroy(c->stream);
  cudaGetLastError();
  delete c;
  return 0;
}

}  // extern "C"
