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

    def __init__(self, num_classes, n_scalars, device, group=None):
        self.num_classes = num_classes
        self.n_scalars = n_scalars
        self.group = group
        self._cc = num_classes * num_classes
        self.buf = torch.zeros(self._cc + max(n_scalars, 1), dtype=torch.float64, device=device)

    def world_size(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def all_reduce(self, cm, scalars, async_op=False):
        """cm: int64 [C,C]; scalars: sequence of 0-dim tensors.  Returns (cm_sum int64 [C,C],
        scalar_sum fp64 [k]) -- sums over ranks; divide the scalars by world_size() for the
        reference's averages.  With async_op=True also returns the work handle (wait before use)."""
        cc = self._cc
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
        cc = self._cc
        return (self.buf[:cc].to(torch.int64).view(self.num_classes, self.num_classes),
                self.buf[cc:cc + self.n_scalars])
