"""Launched by tests/test_gpu_peer.py under torchrun (2 ranks, one GPU each): the CUDA-IPC wiring of
the peer all-reduce, StepReducer's collective backend choice, and LossPathStep(peer=...) against the
NCCL sums of the same values."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import b200ssl
    c = 3
    red = b200ssl.utils.StepReducer(c, 2, dev, backend="peer")
    assert red.peer is not None
    gen = torch.Generator().manual_seed(77 + rank)
    for step in range(25):
        cm = torch.randint(0, 2**35, (c, c), generator=gen, dtype=torch.int64).to(dev)
        sc = [torch.randn((), generator=gen).to(dev) for _ in range(2)]
        want_cm = cm.clone()
        want_sc = torch.stack(sc).double()
        dist.all_reduce(want_cm)
        dist.all_reduce(want_sc)
        if step % 2:
            got_cm, got_sc = red.all_reduce(cm, sc)
        else:
            _, h = red.all_reduce(cm, sc, async_op=True)
            h.wait()
            got_cm, got_sc = red.result()
        torch.cuda.synchronize(dev)
        assert torch.equal(got_cm, want_cm), (rank, step)
        assert torch.allclose(got_sc, want_sc, rtol=1e-15, atol=0), (rank, step, got_sc, want_sc)
    red.peer.status()

    # the step entry with the exchange attached
    g = torch.Generator().manual_seed(5 + rank)
    n, cc, h, w = 2, 2, 64, 64
    peer = b200ssl.utils.PeerAllReduce(cc * cc, 1, dev)
    step = b200ssl.LossPathStep(num_classes=cc, sigma_range=(2, 4), peer=peer)
    img = [torch.rand(n, 3, h, w, generator=g).to(dev) for _ in range(2)]
    tea = [torch.randn(n, cc, h, w, generator=g).to(dev) for _ in range(2)]
    scores = (torch.randn(n, cc, h, w, generator=g) * 3).to(dev)
    blob = torch.nn.functional.avg_pool2d(torch.randn(n, cc, h, w, generator=g), 9, 1, 4)
    target = torch.nn.functional.one_hot(blob.argmax(1), cc).permute(0, 3, 1, 2).float().contiguous().to(dev)
    params, ema = [torch.randn(100, generator=g).to(dev)], [torch.randn(100, generator=g).to(dev)]
    def run(step_obj, peer_obj, rounds, flush_every):
        last = None
        for it in range(rounds):
            out = step_obj(img[0], img[1], tea[0], tea[1], scores, target, params, ema)
            want_cm = out["cm"].clone()
            want_loss = out["loss"].double().reshape(1).clone()
            dist.all_reduce(want_cm)
            dist.all_reduce(want_loss)
            flushed = (it % flush_every) == flush_every - 1
            if flushed:
                peer_obj.result()
            torch.cuda.synchronize(dev)
            checks = [(out, want_cm, want_loss)] if flushed else []
            if last is not None:
                checks.append(last)                       # completed by THIS step's post (fold-collect)
            for o, wc, wl in checks:
                assert torch.equal(o["cm_sum"], wc), rank
                assert torch.allclose(o["loss_sum"].reshape(1), wl, rtol=1e-15, atol=0), rank
            last = (dict(cm_sum=out["cm_sum"].clone() if flushed else out["cm_sum"], loss_sum=out["loss_sum"]),
                    want_cm, want_loss)
        peer_obj.result()
        torch.cuda.synchronize(dev)
        assert torch.equal(last[0]["cm_sum"], last[1]), rank

    run(step, peer, 9, 3)
    peer.status()
    peer.close()
    # the same through CUDA graphs: replayed posts/collects take their sequence numbers from device memory
    peer_g = b200ssl.utils.PeerAllReduce(cc * cc, 1, dev)
    step_g = b200ssl.LossPathStep(num_classes=cc, sigma_range=(2, 4), peer=peer_g, graph=True, ring=2)
    run(step_g, peer_g, 12, 4)
    assert step_g.graph_captures < 12 * 2
    peer_g.status()
    peer_g.close()
    # soft-max mode (lovasz_softmax over 5 classes): the matrix kernel runs first and the Lovasz finalising block
    # posts the 25 counts + the loss
    c5 = 5
    probas = torch.softmax(torch.randn(n, c5, h, w, generator=g) * 2, 1).to(dev)
    labels5 = torch.randint(0, c5, (n, h, w), generator=g).to(dev)
    tea5 = [torch.randn(n, c5, h, w, generator=g).to(dev) for _ in range(2)]
    peer_s = b200ssl.utils.PeerAllReduce(c5 * c5, 1, dev)
    step_s = b200ssl.LossPathStep(num_classes=c5, sigma_range=(2, 4), mode="softmax", peer=peer_s, static_outputs=True)
    last = None
    for it in range(7):
        out = step_s(img[0], img[1], tea5[0], tea5[1], probas, labels5, params, ema)
        want_cm = out["cm"].clone()
        want_loss = out["loss"].double().reshape(1).clone()
        dist.all_reduce(want_cm)
        dist.all_reduce(want_loss)
        torch.cuda.synchronize(dev)
        if last is not None:
            assert torch.equal(last[0], last[1]), rank          # completed by this step's post
            assert torch.allclose(last[2].reshape(1), last[3], rtol=1e-15, atol=0), rank
        last = (out["cm_sum"], want_cm, out["loss_sum"], want_loss)
    peer_s.result()
    torch.cuda.synchronize(dev)
    assert torch.equal(last[0], last[1]) and torch.allclose(last[2].reshape(1), last[3], rtol=1e-15, atol=0), rank
    peer_s.status()
    peer_s.close()
    red.peer.close()
    dist.barrier()
    if rank == 0:
        print("peer_worker ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
