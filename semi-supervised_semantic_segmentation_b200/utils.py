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
    """Packs the integer confusion matrix and k fp32 scalars and all-reduces them together.

    The confusion matrix travels as int64 (exact); the scalars travel as fp64 in a second tensor of
    the same coalesced call (NCCL SUM is type-homogeneous).  Nothing is read back on the host: the
    caller consumes `cm` / `scalars` lazily, so no step ends in a device sync.
    """

    def __init__(self, num_classes, n_scalars, device, group=None):
        self.num_classes = num_classes
        self.n_scalars = n_scalars
        self.group = group
        self.cm = torch.zeros(num_classes * num_classes, dtype=torch.int64, device=device)
        self.scalars = torch.zeros(max(n_scalars, 1), dtype=torch.float64, device=device)

    def world_size(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def all_reduce(self, cm, scalars, async_op=False):
        """cm: int64 [C,C]; scalars: sequence of 0-dim tensors.  Returns (cm_sum [C,C], scalar_sum [k])
        -- sums over ranks; divide the scalars by world_size() for the reference's averages."""
        self.cm.copy_(cm.reshape(-1))
        if self.n_scalars:
            torch.stack([s.detach().to(torch.float64).reshape(()) for s in scalars], out=self.scalars[:self.n_scalars])
        handles = []
        if self.world_size() > 1:
            handles.append(dist.all_reduce(self.cm, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op))
            if self.n_scalars:
                handles.append(dist.all_reduce(self.scalars, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op))
        out = (self.cm.view(self.num_classes, self.num_classes), self.scalars[:self.n_scalars])
        return (out, handles) if async_op else out
