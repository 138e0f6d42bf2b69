"""The NVLink peer-memory all-reduce (csrc/peer.cu) against plain integer / fp64 sums.

On a one-GPU box the protocol (sequence-tagged words, slot ring, acknowledgements, rank-order sums)
is exercised with several communicators living in ONE process on cuda:0, wired with plain pointers
(b200ssl_peer_connect_ptrs); the CUDA-IPC wiring between processes runs when the box has >= 2 GPUs.
Oracle = the sum the reference's `dist.reduce` computes (utils/utils.py:43-54): exact for the int64
counts, and -- because the ranks are added in a fixed order in fp64 -- bit-identical on every rank
for the scalars."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def utils():
    os.environ.setdefault("B200SSL_PEER_TIMEOUT_MS", "5000")
    import b200ssl
    return b200ssl.utils


def _make(utils, world, n_ints, n_floats, dev):
    comms = [utils.PeerAllReduce(n_ints, n_floats, dev, _inprocess=(r, world)) for r in range(world)]
    utils.PeerAllReduce.connect_inprocess(comms)
    return comms


@pytest.mark.parametrize("world,n_ints,n_floats,lazy", [(2, 4, 1, True), (2, 4, 1, False), (3, 441, 3, True),
                                                        (8, 361, 2, True), (4, 0, 5, True), (2, 2044, 8, True)])
def test_inprocess_ranks_sum_exactly(utils, world, n_ints, n_floats, lazy):
    dev = torch.device("cuda:0")
    comms = _make(utils, world, n_ints, n_floats, dev)
    streams = [torch.cuda.Stream(dev) for _ in range(world)]
    gen = torch.Generator().manual_seed(world * 1000 + n_ints)
    steps = 11                                            # > 2 trips around the 4-deep slot ring
    pending = None
    for step in range(steps):
        ints = torch.randint(-2**40, 2**40, (world, max(n_ints, 1)), generator=gen, dtype=torch.int64)
        floats = torch.randn(world, max(n_floats, 1), generator=gen) * 10.0
        d_ints, d_floats = ints.to(dev), floats.to(dev)
        torch.cuda.synchronize(dev)
        outs = [None] * world
        if lazy:
            # lazy: the post of step s also completes step s-1 (checked below through `pending`); result()
            # flushes step s on the rank's stream
            for r in range(world):
                with torch.cuda.stream(streams[r]):
                    comms[r].all_reduce(d_ints[r, :n_ints] if n_ints else None,
                                        [d_floats[r, i] for i in range(n_floats)], lazy=True)
            if step % 3 != 2:                             # every third step is left to the NEXT post's fold-collect
                for r in range(world):
                    with torch.cuda.stream(streams[r]):
                        outs[r] = tuple(t.clone() for t in comms[r].result())
            else:
                pending = (ints, floats, [c._pending for c in comms])
                continue
        else:
            # non-lazy collects spin on their own stream until every rank has posted: issue all the
            # posts first (one rank's blocking collect must never sit in front of another's post)
            import ctypes as C
            from b200ssl import _lib
            res = []
            for r in range(world):
                fl = [d_floats[r, i] for i in range(n_floats)]
                fptrs = (C.c_void_p * max(n_floats, 1))(*[f.data_ptr() for f in fl])
                _lib.check(_lib.lib.b200ssl_peer_post(comms[r].handle, d_ints[r].data_ptr() if n_ints else None, n_ints,
                                                      fptrs, n_floats, None, None, C.c_void_p(streams[r].cuda_stream)))
            for r in range(world):
                oi = torch.zeros(max(n_ints, 1), dtype=torch.int64, device=dev)
                of = torch.zeros(max(n_floats, 1), dtype=torch.float64, device=dev)
                s = C.c_void_p(streams[r].cuda_stream)
                _lib.check(_lib.lib.b200ssl_peer_collect(comms[r].handle, oi.data_ptr(), of.data_ptr(), s))
                _lib.check(_lib.lib.b200ssl_peer_collect(comms[r].handle, oi.data_ptr(), of.data_ptr(), s))  # idempotent
                res.append((oi[:n_ints], of[:n_floats]))
            outs = res
        torch.cuda.synchronize(dev)

        def expect(ints, floats):
            want_i = ints[:, :n_ints].sum(0)
            want_f = torch.zeros(n_floats, dtype=torch.float64)
            for r in range(world):                        # rank order, fp64, like the kernel
                want_f = want_f + floats[r, :n_floats].double()
            return want_i, want_f

        def same(got, want):
            return torch.equal(got[0].cpu(), want[0]) and \
                np.array_equal(got[1].cpu().numpy().view(np.uint64), want[1].numpy().view(np.uint64))

        for r in range(world):
            assert same(outs[r], expect(ints, floats)), (step, r)
        if pending is not None:                           # the step that was only completed by this step's post
            p_ints, p_floats, bufs = pending
            for r in range(world):
                assert same((bufs[r][0][:n_ints], bufs[r][1][:n_floats]), expect(p_ints, p_floats)), (step, r, "folded")
            pending = None
    for c in comms:
        c.status()
    for c in comms:
        c.close(barrier=False)


def test_rank_issued_ahead_of_its_peer_still_sums(utils):
    """Rank 0's three steps are issued before rank 1 has issued anything: its fold-collects of steps 1 and 2
    wait (on rank 0's stream only) until rank 1's posts arrive, nothing deadlocks, everything sums."""
    dev = torch.device("cuda:0")
    comms = _make(utils, 2, 1, 1, dev)
    s0, s1 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    vals = torch.arange(1, 7, dtype=torch.int64, device=dev)
    fl = torch.arange(1, 7, dtype=torch.float32, device=dev) * 0.5
    torch.cuda.synchronize(dev)
    got0 = []
    with torch.cuda.stream(s0):
        for k in range(3):                                # rank 0: three steps ahead
            comms[0].all_reduce(vals[k:k + 1], [fl[k]], lazy=True)
    with torch.cuda.stream(s1):
        for k in range(3):
            comms[1].all_reduce(vals[k + 3:k + 4], [fl[k + 3]], lazy=True)
            got = comms[1].result()
            got0.append((int(got[0].item()), float(got[1].item())))
    torch.cuda.synchronize(dev)
    assert got0 == [(1 + 4, 0.5 + 2.0), (2 + 5, 1.0 + 2.5), (3 + 6, 1.5 + 3.0)]
    for c in comms:
        c.status()
        c.close(barrier=False)


def test_missing_peer_times_out_instead_of_hanging(utils):
    os.environ["B200SSL_PEER_TIMEOUT_MS"] = "200"
    try:
        dev = torch.device("cuda:0")
        comms = _make(utils, 2, 1, 0, dev)
    finally:
        os.environ["B200SSL_PEER_TIMEOUT_MS"] = "5000"
    x = torch.ones(1, dtype=torch.int64, device=dev)
    comms[0].all_reduce(x, [], lazy=True)                 # rank 1 never posts
    comms[0].result()
    torch.cuda.synchronize(dev)
    from b200ssl._lib import B200SSLError
    with pytest.raises(B200SSLError):
        comms[0].status()
    for c in comms:
        c.close(barrier=False)


def test_loss_path_step_with_peer_world1(utils):
    """world=1 through the step entry: cm_sum / loss_sum equal the step's own matrix and loss."""
    import b200ssl
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    n, c, h, w = 2, 2, 64, 64
    peer = utils.PeerAllReduce(c * c, 1, dev)
    step = b200ssl.LossPathStep(num_classes=c, sigma_range=(2, 4), peer=peer)
    img = [torch.rand(n, 3, h, w, generator=g).to(dev) for _ in range(2)]
    tea = [torch.randn(n, c, h, w, generator=g).to(dev) for _ in range(2)]
    scores = (torch.randn(n, c, h, w, generator=g) * 3).to(dev)
    blob = torch.nn.functional.avg_pool2d(torch.randn(n, c, h, w, generator=g), 9, 1, 4)
    target = torch.nn.functional.one_hot(blob.argmax(1), c).permute(0, 3, 1, 2).float().contiguous().to(dev)
    params = [torch.randn(100, generator=g).to(dev)]
    ema = [torch.randn(100, generator=g).to(dev)]
    last = None
    for it in range(6):
        out = step(img[0], img[1], tea[0], tea[1], scores, target, params, ema)
        if it % 2:
            peer.result()                                 # explicit flush of this step ...
        torch.cuda.synchronize(dev)
        if last is not None:                              # ... or completed by the next step's post
            assert torch.equal(last["cm_sum"], last["cm"])
            assert float(last["loss_sum"]) == float(last["loss"].double())
        if it % 2:
            assert torch.equal(out["cm_sum"], out["cm"])
            assert float(out["loss_sum"]) == float(out["loss"].double())
        last = out
    peer.status()
    peer.close()


@pytest.mark.skipif(torch.cuda.is_available() and torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_processes_over_cuda_ipc():
    env = dict(os.environ, PYTHONPATH=ROOT, B200SSL_PEER_TIMEOUT_MS="10000")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29641", os.path.join(ROOT, "tests", "peer_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "peer_worker ok" in r.stdout
