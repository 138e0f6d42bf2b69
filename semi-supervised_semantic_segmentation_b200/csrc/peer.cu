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
//   post     (no waiting in steady state): every rank stores its words into ITS row of EVERY rank's
//            mailbox.  A word is 64 bit = (sequence number << 32 | 32 payload bits), so a single 8-byte
//            store publishes data and "ready" flag together -- no fence, no second round trip (the LL
//            idea).  int64 counts travel as two words, fp32 scalars as one.
//   collect  polls the world rows of the LOCAL mailbox until every word carries the step's sequence
//            number, adds them in rank order (int64 adds for counts, fp64 adds for scalars: identical
//            bits on every rank, exact for counts), then acknowledges the slot to every peer.
//
// Round 2: the sequence numbers are device-resident (peer_device.cuh) and the collect of step s-1 is
// folded into the block that posts step s (`prev_*_out`): a step with the exchange attached issues
// exactly the launches of a single-GPU step -- no collect kernel, no events, no communicator stream --
// and is CUDA-graph capturable.  Only the LAST step of a run needs an explicit b200ssl_peer_collect
// (idempotent: it does nothing if the step has been collected).
//
// Mailbox rows are kPeerDepth deep (slot = seq % depth); a post for step s only reuses the slot of
// step s-depth after every peer has acknowledged collecting that step.  Every spin is bounded by a
// timeout (default 20 s, B200SSL_PEER_TIMEOUT_MS) that raises a sticky status flag instead of hanging
// the GPU.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "peer_device.cuh"

namespace b200ssl {

__global__ void __launch_bounds__(256) peer_post_kernel(const __grid_constant__ PeerDev c,
                                                        const long long* __restrict__ ints, int n_ints,
                                                        const __grid_constant__ PeerFloats f, int n_floats,
                                                        long long* __restrict__ prev_ints_out,
                                                        double* __restrict__ prev_floats_out, int collect_this,
                                                        long long* __restrict__ ints_out, double* __restrict__ floats_out) {
  const unsigned seq = peer_post_block(c, ints, n_ints, f, n_floats);
  if (prev_ints_out || prev_floats_out) peer_collect_block(c, seq - 1u, n_ints, n_floats, prev_ints_out, prev_floats_out);
  if (collect_this) peer_collect_block(c, seq, n_ints, n_floats, ints_out, floats_out);   // all-reduce in one launch
}

// collects the latest posted step unless it has been collected already (idempotent flush)
__global__ void __launch_bounds__(256) peer_collect_kernel(const __grid_constant__ PeerDev c, int n_ints, int n_floats,
                                                           long long* __restrict__ ints_out,
                                                           double* __restrict__ floats_out) {
  __shared__ unsigned s_seq;
  if (threadIdx.x == 0) s_seq = (unsigned)ld_sys(c.mail[c.rank] + kSeqOffset);
  __syncthreads();
  peer_collect_block(c, s_seq, n_ints, n_floats, ints_out, floats_out);
}

}  // namespace b200ssl

struct b200ssl_peer_comm {
  int rank = 0, world = 1, dev = 0;
  unsigned long long* local = nullptr;
  unsigned long long* mail[b200ssl::kPeerMaxRanks] = {};
  bool opened[b200ssl::kPeerMaxRanks] = {};
  bool connected = false;
  int n_ints = -1, n_floats = -1;   // payload shape, fixed by the first post
  unsigned long long timeout_ns = 20ull * 1000 * 1000 * 1000;
};

namespace b200ssl {

static PeerDev device_view(const b200ssl_peer_comm* c) {
  PeerDev d;
  for (int r = 0; r < kPeerMaxRanks; ++r) d.mail[r] = c->mail[r];
  d.rank = c->rank;
  d.world = c->world;
  d.timeout_ns = c->timeout_ns;
  return d;
}

static int check_payload(b200ssl_peer_comm* c, int n_ints, int n_floats, const char* who) {
  B200SSL_REQUIRE(c && c->connected, "%s: communicator not connected", who);
  B200SSL_REQUIRE(n_ints >= 0 && n_floats >= 0 && n_floats <= kPeerMaxFloats && 2 * n_ints + n_floats <= kPeerMaxWords &&
                      n_ints + n_floats > 0,
                  "%s: payload of %d counts + %d scalars does not fit (%d words, %d scalars max)", who, n_ints,
                  n_floats, kPeerMaxWords, kPeerMaxFloats);
  // a collect sums whatever shape was posted: one shape per communicator keeps post and collect in agreement
  B200SSL_REQUIRE(c->n_ints < 0 || (c->n_ints == n_ints && c->n_floats == n_floats),
                  "%s: this communicator exchanges %d counts + %d scalars, not %d + %d", who, c->n_ints, c->n_floats,
                  n_ints, n_floats);
  c->n_ints = n_ints;
  c->n_floats = n_floats;
  return 0;
}

// used by step.cu when the producing kernel runs the exchange by itself: validates and hands out the device
// view.  No host-side state advances, so a caller that fails before its kernel launches leaves nothing behind.
int peer_tail(b200ssl_peer_comm* c, int n_ints, int n_floats, PeerTail* out) {
  int rc = check_payload(c, n_ints, n_floats, "peer_tail");
  if (rc) return rc;
  out->dev = device_view(c);
  out->enabled = 1;
  return 0;
}

int peer_post_impl(b200ssl_peer_comm* c, const long long* ints, int n_ints, const float* const* floats_host,
                   int n_floats, long long* prev_ints_out, double* prev_floats_out, cudaStream_t s) {
  int rc = check_payload(c, n_ints, n_floats, "peer_post");
  if (rc) return rc;
  B200SSL_REQUIRE(n_ints == 0 || ints, "peer_post: null counts");
  PeerFloats f = {};
  for (int i = 0; i < n_floats; ++i) {
    B200SSL_REQUIRE(floats_host && floats_host[i], "peer_post: null scalar pointer");
    f.p[i] = floats_host[i];
  }
  prof_begin("peer_post", s);
  peer_post_kernel<<<1, 256, 0, s>>>(device_view(c), ints, n_ints, f, n_floats, prev_ints_out, prev_floats_out, 0,
                                     nullptr, nullptr);
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
                      int n_floats, long long* prev_ints_out, double* prev_floats_out, b200ssl_stream_t stream) {
  return b200ssl::peer_post_impl(c, ints, n_ints, floats_host, n_floats, prev_ints_out, prev_floats_out,
                                 (cudaStream_t)stream);
}

int b200ssl_peer_collect(b200ssl_peer_comm* c, long long* ints_out, double* floats_out, b200ssl_stream_t stream) {
  using namespace b200ssl;
  B200SSL_REQUIRE(c && c->connected, "peer_collect: communicator not connected");
  if (c->n_ints < 0) return 0;   // nothing has ever been posted
  B200SSL_REQUIRE((c->n_ints == 0 || ints_out) && (c->n_floats == 0 || floats_out), "peer_collect: null output");
  prof_begin("peer_collect", (cudaStream_t)stream);
  peer_collect_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(device_view(c), c->n_ints, c->n_floats, ints_out, floats_out);
  return check_launch("peer_collect");
}

int b200ssl_peer_allreduce(b200ssl_peer_comm* c, const long long* ints, int n_ints, const float* const* floats_host,
                           int n_floats, long long* ints_out, double* floats_out, b200ssl_stream_t stream) {
  using namespace b200ssl;
  int rc = check_payload(c, n_ints, n_floats, "peer_allreduce");
  if (rc) return rc;
  B200SSL_REQUIRE((n_ints == 0 || (ints && ints_out)) && (n_floats == 0 || floats_out), "peer_allreduce: null argument");
  PeerFloats f = {};
  for (int i = 0; i < n_floats; ++i) {
    B200SSL_REQUIRE(floats_host && floats_host[i], "peer_allreduce: null scalar pointer");
    f.p[i] = floats_host[i];
  }
  prof_begin("peer_allreduce", (cudaStream_t)stream);
  peer_post_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(device_view(c), ints, n_ints, f, n_floats, nullptr, nullptr, 1,
                                                        ints_out, floats_out);
  return check_launch("peer_allreduce");
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
  cudaDeviceSynchronize();
  for (int r = 0; r < c->world; ++r)
    if (c->opened[r]) cudaIpcCloseMemHandle(c->mail[r]);
  if (c->local) cudaFree(c->local);
  cudaGetLastError();
  delete c;
  return 0;
}

}  // extern "C"
