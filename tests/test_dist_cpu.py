"""world_size-2 gloo test of the packed per-step all-reduce (confusion matrix || loss scalars) and
of reduce_tensor: the reduced matrix equals the matrix of the concatenated batch exactly."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import b200ssl
    c = 5
    gen = torch.Generator().manual_seed(100)
    labels = torch.randint(0, c, (4, 32, 32), generator=gen)
    preds = torch.randint(0, c, (4, 32, 32), generator=gen)
    shard = slice(rank * 2, rank * 2 + 2)             # DistributedSampler-style split by image
    cm_local, _ = oracle.confusion_matrix(labels[shard].numpy(), preds[shard].numpy(), c)
    red = b200ssl.utils.StepReducer(c, 2, torch.device("cpu"))
    cm, sc = red.all_reduce(torch.from_numpy(cm_local), [torch.tensor(1.5 + rank), torch.tensor(0.25)])
    t = torch.tensor(float(rank + 1))
    same = b200ssl.utils.reduce_tensor(t)
    q.put((rank, cm.clone().numpy(), sc.clone().numpy(), float(same), same is t))
    dist.barrier()
    dist.destroy_process_group()


def test_step_reducer_and_reduce_tensor_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    gen = torch.Generator().manual_seed(100)
    labels = torch.randint(0, 5, (4, 32, 32), generator=gen)
    preds = torch.randint(0, 5, (4, 32, 32), generator=gen)
    whole, _ = oracle.confusion_matrix(labels.numpy(), preds.numpy(), 5)
    for rank, cm, sc, red, same_obj in res:
        assert np.array_equal(cm, whole)                       # integer-exact on every rank
        assert np.allclose(sc, [1.5 + 2.5, 0.5])
        assert same_obj
    assert res[0][3] == 3.0                                     # reduce(dst=0): rank 0 holds the sum
