"""The one communication helper on the hot path, plus the packed per-step all-reduce.

    reference utils/utils.py:43-54  reduce_tensor(inp): dist.reduce(dst=0) in place on the tensor
    (new) StepReducer: ONE all-reduce per step for [C*C int64 confusion matrix || loss scalars]
          instead of the reference's 4-6 scalar reduces with an .item() sync each (train.py:53-114).
"""
import torch
import torch.distributed as dist


def reduce_tensor(inp):
    """
    Reduce the loss from all processes so that
    process with rank 0 has the averaged results.
    (Same contract as the reference: returns the SAME tensor, summed in place on rank 0; the caller
    divides by world_size.)
    """
    if not dist.is_available() or not dist.is_initialized():
        return inp
    world_size = dist.get_world_size()
    if world_size < 2:
        return inp
    with torch.no_grad():
        reduced_inp = inp
        dist.reduce(reduced_inp, dst=0)
    return reduced_inp


class StepReducer:
    """Packs the integer confusion matrix and k loss scalars into ONE all-reduce per step.

    The packed buffer is fp64: pixel counts below 2**53 are represented exactly and their sums are
    exact, so the reduced matrix equals the single-process matrix of the concatenated batch bit for
    bit (configs[4], 10 000 masks of 1024x2048, totals 2**34.3 pixels); the fp32 scalars ride along
    in the same message.  Nothing is read back on the host: the caller consumes `cm` / `scalars`
    lazily, so no step ends in a device sync.
    """

    def __init__(self, num_classes, n_scalars, device, group=None, backend="auto"):
        """backend: "peer" = NVLink peer-memory exchange (`PeerAllReduce`, CUDA, one box), "dist" =
        torch.distributed all_reduce (NCCL / gloo), "auto" = peer when every rank can set it up."""
        if backend not in ("auto", "peer", "dist"):
            raise ValueError("backend must be 'auto', 'peer' or 'dist'")
        self.num_classes = num_classes
        self.n_scalars = n_scalars
        self.group = group
        self._cc = num_classes * num_classes
        self.buf = torch.zeros(self._cc + max(n_scalars, 1), dtype=torch.float64, device=device)
        self.peer = None
        if backend != "dist" and torch.device(device).type == "cuda" and self.world_size() > 1:
            self.peer = make_peer_all_reduce(self._cc, n_scalars, device, group, required=(backend == "peer"))

    def world_size(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def all_reduce(self, cm, scalars, async_op=False):
        """cm: int64 [C,C]; scalars: sequence of 0-dim tensors.  Returns (cm_sum int64 [C,C],
        scalar_sum fp64 [k]) -- sums over ranks; divide the scalars by world_size() for the
        reference's averages.  With async_op=True also returns the work handle (wait before use)."""
        cc = self._cc
        if self.peer is not None:
            ints, floats = self.peer.all_reduce(cm.reshape(-1), [s.detach().reshape(()) for s in scalars],
                                                lazy=async_op)
            if async_op:
                return (self, _PeerHandle(self.peer))
            return ints.view(self.num_classes, self.num_classes), floats
        self.buf[:cc].copy_(cm.reshape(-1))
        for i, s in enumerate(scalars):
            self.buf[cc + i].copy_(s.detach().reshape(()))
        handle = None
        if self.world_size() > 1:
            handle = dist.all_reduce(self.buf, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
        out = (self.buf[:cc].to(torch.int64).view(self.num_classes, self.num_classes),
               self.buf[cc:cc + self.n_scalars])
        if async_op:
            # the int64 view above was taken before the reduction completed: re-derive it lazily
            return (self, handle)
        return out

    def result(self):
        """(cm_sum int64 [C,C], scalar_sum fp64 [k]) of the last all_reduce (after its handle completed)."""
        if self.peer is not None:
            ints, floats = self.peer.result()
            return ints.view(self.num_classes, self.num_classes), floats
        cc = self._cc
        return (self.buf[:cc].to(torch.int64).view(self.num_classes, self.num_classes),
                self.buf[cc:cc + self.n_scalars])


class _PeerHandle:
    """Work-handle look-alike for StepReducer.all_reduce(async_op=True) on the peer backend."""

    def __init__(self, peer):
        self._peer = peer

    def wait(self):
        self._peer.result()
        return True


def make_peer_all_reduce(n_ints, n_floats, device, group=None, required=False):
    """Collectively set up a `PeerAllReduce` and prove it with one exchange.  Returns it when EVERY rank
    succeeded, otherwise None on every rank (the decision is itself all-reduced, so ranks cannot
    disagree about which transport the following steps use).  required=True raises instead."""
    import warnings
    peer, err = None, None
    try:
        peer = PeerAllReduce(n_ints, n_floats, device, group)
        rank, world = peer.rank, peer.world
        ints = torch.arange(n_ints, dtype=torch.int64, device=device) * (rank + 1) - 3 if n_ints else None
        floats = [torch.tensor(0.5 * (rank + 1) + i, dtype=torch.float32, device=device) for i in range(n_floats)]
        peer.all_reduce(ints, floats, lazy=True)
        got_i, got_f = peer.result()
        tri = world * (world + 1) // 2
        ok = True
        if n_ints:
            ok = ok and torch.equal(got_i.cpu(), torch.arange(n_ints, dtype=torch.int64) * tri - 3 * world)
        if n_floats:
            want = torch.tensor([0.5 * tri + i * world for i in range(n_floats)], dtype=torch.float64)
            ok = ok and torch.equal(got_f.cpu(), want)
        peer.status()
        if not ok:
            err = "self-test sums are wrong"
    except Exception as e:   # noqa: BLE001  (any set-up failure means: use the library collective)
        err = f"{type(e).__name__}: {e}"
    flag = torch.tensor([0 if err else 1], dtype=torch.int32,
                        device=device if dist.get_backend(group) == "nccl" else "cpu")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag) == 1:
        return peer
    if peer is not None:
        try:
            peer.close(barrier=False)
        except Exception:   # noqa: BLE001
            pass
    msg = f"b200ssl: NVLink peer all-reduce unavailable on this job ({err or 'another rank failed'})"
    if required:
        raise RuntimeError(msg)
    warnings.warn(msg + "; falling back to torch.distributed all_reduce")
    return None


class PeerAllReduce:
    """The per-step all-reduce of [int64 counts || fp32 scalars] over NVLink peer memory
    (csrc/peer.cu, `b200ssl_peer_*`): no NCCL call on the data path.

    torch.distributed is used ONCE, at construction, to exchange the 64-byte CUDA IPC handles of the
    mailboxes (and for the barriers around set-up and tear-down).  Per step every rank stores its
    words directly into every rank's mailbox; the rows of step s-1 are added in rank order by the very
    block that posts step s (no separate collect launch in steady state, sequence numbers on the device:
    the step is CUDA-graph capturable); `result()` flushes the last step on the current stream.

    Results are sums over ranks: counts int64 (exact), scalars fp64 (fixed rank order, bit-identical
    on every rank).  Replaces the reference's `reduce_tensor` calls (utils/utils.py:43-54).
    """

    def __init__(self, n_ints, n_floats, device, group=None, _inprocess=None):
        import ctypes as C
        from . import _lib
        self._lib, self._C = _lib, C
        lib = _lib.lib
        if 2 * n_ints + n_floats > _lib.PEER_MAX_WORDS or n_floats > _lib.PEER_MAX_FLOATS or n_ints + n_floats < 1:
            raise ValueError(f"PeerAllReduce: {n_ints} counts + {n_floats} scalars do not fit one exchange")
        self.n_ints, self.n_floats, self.device, self.group = n_ints, n_floats, torch.device(device), group
        if _inprocess is not None:                      # (rank, world): several communicators in one process (tests)
            self.rank, self.world = _inprocess
        elif dist.is_available() and dist.is_initialized():
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        else:
            self.rank, self.world = 0, 1
        if self.world > _lib.PEER_MAX_RANKS:
            raise ValueError(f"PeerAllReduce: world size {self.world} > {_lib.PEER_MAX_RANKS}")
        self._comm = C.c_void_p()
        handle = (C.c_ubyte * _lib.PEER_HANDLE_BYTES)()
        # Every rank walks through the same collectives (all_gather of the handles, barrier) whether or
        # not its own set-up succeeded; a failure is raised only afterwards, so a rank that cannot map
        # a peer never leaves the others waiting.
        err = None
        with torch.cuda.device(self.device):
            try:
                _lib.check(lib.b200ssl_peer_create(self.rank, self.world, C.byref(self._comm), handle), "peer_create")
            except Exception as e:   # noqa: BLE001
                err = e
            if _inprocess is None:
                if self.world > 1:
                    handles = [None] * self.world
                    dist.all_gather_object(handles, b"" if err else bytes(handle), group=group)
                    if err is None and any(len(h) != _lib.PEER_HANDLE_BYTES for h in handles):
                        err = RuntimeError("PeerAllReduce: another rank could not create its mailbox")
                    blob = b"".join(handles)
                else:
                    blob = bytes(handle)
                if err is None:
                    try:
                        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
                        _lib.check(lib.b200ssl_peer_connect(self._comm, buf), "peer_connect")
                    except Exception as e:   # noqa: BLE001
                        err = e
                if self.world > 1:
                    dist.barrier(group=group)
        if err is not None:
            raise err
        # results live in a ring as deep as the mailboxes, so a consumer may lag a few steps behind
        self._ring = [(torch.zeros(max(n_ints, 1), dtype=torch.int64, device=self.device),
                       torch.zeros(max(n_floats, 1), dtype=torch.float64, device=self.device))
                      for _ in range(_lib.PEER_DEPTH)]
        self._step = 0
        self._pending = None          # (counts, scalars) of the last posted step
        self._closed = False

    @staticmethod
    def connect_inprocess(comms):
        """Wire communicators that live in one process (same device or peer-enabled devices)."""
        import ctypes as C
        from . import _lib
        ptrs = (C.c_void_p * len(comms))(*[_lib.lib.b200ssl_peer_mailbox(c._comm) for c in comms])
        for c in comms:
            with torch.cuda.device(c.device):
                _lib.check(_lib.lib.b200ssl_peer_connect_ptrs(c._comm, ptrs), "peer_connect_ptrs")

    @property
    def handle(self):
        return self._comm

    def begin_step(self, out=None):
        """Bookkeeping for one posted step (used by LossPathStep and all_reduce): `out` = (counts int64,
        scalars fp64) tensors that will receive this step's sums (default: the next pair of the internal
        ring).  Returns (out, prev) where `prev` is the pair of the previous posted step -- the caller hands
        it to the post as the fold-collect destination -- or None when there is no previous step."""
        if out is None:
            self._step += 1
            out = self._ring[self._step % len(self._ring)]
        prev, self._pending = self._pending, out
        return out, prev

    def all_reduce(self, ints, floats, lazy=True):
        """ints: int64 tensor with n_ints elements (or None); floats: sequence of n_floats fp32 0-dim
        tensors.  lazy=True: posts this step and, in the same launch, completes the previous step's
        exchange; the returned (counts, scalars) tensors are valid on the current stream after `result()`
        (or after the next lazy all_reduce).  lazy=False: post + collect of this step in one launch; the
        current stream then waits for the slowest rank's post."""
        C, lib = self._C, self._lib.lib
        if len(floats) != self.n_floats:
            raise ValueError(f"PeerAllReduce: expected {self.n_floats} scalars, got {len(floats)}")
        if self.n_ints:
            self._lib.require_cuda(ints, "ints", torch.int64)
            if ints.numel() != self.n_ints or not ints.is_contiguous():
                raise ValueError("PeerAllReduce: counts must be a contiguous int64 tensor of n_ints elements")
        fl = [self._lib.require_cuda(f, "scalar", torch.float32) for f in floats]
        fptrs = (C.c_void_p * max(len(fl), 1))(*[f.data_ptr() for f in fl])
        (out_i, out_f), prev = self.begin_step()
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        iptr = ints.data_ptr() if self.n_ints else None
        with torch.cuda.device(self.device):
            if lazy:
                self._lib.check(lib.b200ssl_peer_post(self._comm, iptr, self.n_ints, fptrs, self.n_floats,
                                                      prev[0].data_ptr() if prev else None,
                                                      prev[1].data_ptr() if prev else None, stream), "peer_post")
            else:
                if prev is not None:      # an earlier lazy step must not stay uncollected behind this one
                    self._lib.check(lib.b200ssl_peer_collect(self._comm, prev[0].data_ptr(), prev[1].data_ptr(),
                                                             stream), "peer_collect")
                self._lib.check(lib.b200ssl_peer_allreduce(self._comm, iptr, self.n_ints, fptrs, self.n_floats,
                                                           out_i.data_ptr(), out_f.data_ptr(), stream),
                                "peer_allreduce")
        return out_i[:self.n_ints], out_f[:self.n_floats]

    def result(self):
        """Complete the exchange of the last posted step on the current stream (a no-op on the device if
        it has already been collected) and return its (counts, scalars)."""
        if self._pending is None:
            raise RuntimeError("PeerAllReduce.result: nothing has been posted")
        out_i, out_f = self._pending
        stream = self._C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        with torch.cuda.device(self.device):
            self._lib.check(self._lib.lib.b200ssl_peer_collect(self._comm, out_i.data_ptr(), out_f.data_ptr(),
                                                               stream), "peer_collect")
        return out_i[:self.n_ints], out_f[:self.n_floats]

    def status(self):
        """Synchronous health check: raises if any wait timed out."""
        with torch.cuda.device(self.device):
            self._lib.check(self._lib.lib.b200ssl_peer_status(self._comm), "peer_status")

    def close(self, barrier=True):
        if self._closed:
            return
        self._closed = True
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            if barrier and self.world > 1 and dist.is_available() and dist.is_initialized():
                dist.barrier(group=self.group)        # nobody may still be storing into a mailbox that is freed
            self._lib.lib.b200ssl_peer_destroy(self._comm)
        self._comm = None
