"""Drop-in for the reference's cowmix.py (same function names and argument meaning).

    reference cowmix.py:40-69  generate_cowmix_masks_like  -> b200ssl_cowmix_mask
    reference cowmix.py:72-73  mix_with_mask               -> b200ssl_mix2

Host-side work is limited to what the reference also does on the host: the 2N uniform draws
for p and sigma (CPU generator), the N*K normalised Gaussian taps and the N erfinv threshold
factors; they are uploaded in one pinned, non-blocking copy.  Everything that touches pixels runs
in the CUDA library.
"""
import ctypes as C
import math

import torch

from . import _lib
from ._lib import lib, check, stream_ptr, require_cuda


def kernel_size_for(sigma_max):
    """cowmix.py:30 -- `int(round(sigma_max * 3) * 2) + 1` with Python's half-to-even round."""
    return int(round(sigma_max * 3) * 2) + 1


_x2_cache = {}
_range_cache = {}
_SQRT2 = math.sqrt(2.0)


def _neg_x2(size):
    x2 = _x2_cache.get(size)
    if x2 is None:
        x = torch.arange(-size // 2, size // 2).float()
        if size % 2 == 0:
            x = x + 0.5
        x2 = _x2_cache[size] = (-x.pow(2.0)).unsqueeze(0)
    return x2


def gaussian_taps(size, sigmas, out=None):
    """All per-sample tap vectors at once: [N, size] fp32 on the CPU (written into `out` if given).

    Batched restatement of cowmix.py:6-24.  For odd `size` the abscissa runs from -(k+1) to k-1
    (the reference's arange(-size//2, size//2)), i.e. the peak sits one tap right of centre.
    Every elementwise op is the same ATen CPU op the reference issues per sample and the row sums
    use the same inner reduction as its 1-D `gauss.sum()`, so the values are bit-identical to the
    reference's (tests/test_host_logic.py checks this against the golden vectors).
    """
    denom = sigmas.pow(2).mul_(2)                        # float(2 * sigma ** 2) per sample
    g = torch.div(_neg_x2(size), denom.unsqueeze(1)).exp_()
    return torch.div(g, g.sum(1, keepdim=True), out=out)


def draw_mask_parameters(n, mask_proportion_range, sigma_range):
    """cowmix.py:44-51: p ~ U(lo,hi), sigma ~ logU(lo,hi), both from the CPU generator, p first.
    `Uniform(lo, hi).rsample([n])` is `lo + torch.rand([n]) * (hi - lo)` on fp32 scalars; issuing
    those ops directly (in place, with the fp32 constants cached per range) skips the distribution
    objects and their argument validation -- same bits, same RNG consumption."""
    key = (mask_proportion_range[0], mask_proportion_range[1], sigma_range[0], sigma_range[1])
    c = _range_cache.get(key)
    if c is None:
        lo, hi = torch.tensor(mask_proportion_range[0]), torch.tensor(mask_proportion_range[1])
        l0 = torch.tensor(math.log(float(sigma_range[0])))
        l1 = torch.tensor(math.log(float(sigma_range[1])))
        if not (lo.dtype == hi.dtype == torch.float32):
            raise TypeError("mask_proportion_range must be a tuple of python floats")
        # fp32 values held as python floats: `t * float` and `t + float` round exactly like fp32 tensors
        c = _range_cache[key] = ((hi - lo).item(), lo.item(), (l1 - l0).item(), l0.item())
    p = torch.rand([n]).mul_(c[0]).add_(c[1])
    sigmas = torch.rand([n]).mul_(c[2]).add_(c[3]).exp_()
    return p, sigmas


class _Staging:
    """Ring of pinned host buffers (+ matching device buffers) for the per-step taps/factors upload; a
    slot is reused only after the copy that read it has completed (event per slot), so the host can
    run ahead of the GPU and nothing is allocated per step."""

    def __init__(self, device, slots=8, floats=64 * 200):
        self.device = device
        self.floats = floats
        self.host = [None] * slots
        self.dev = [None] * slots
        self.events = [None] * slots
        self.i = 0

    def slot(self, n_floats):
        i = self.i
        self.i = (i + 1) % len(self.host)
        if self.host[i] is None or self.host[i].numel() < n_floats:
            size = max(n_floats, self.floats)
            self.host[i] = torch.empty(size, dtype=torch.float32, pin_memory=True)
            self.dev[i] = torch.empty(size, dtype=torch.float32, device=self.device)
            self.events[i] = torch.cuda.Event()
        else:
            self.events[i].synchronize()
        return i

    def upload(self, i, n_floats):
        d = self.dev[i]
        d[:n_floats].copy_(self.host[i][:n_floats], non_blocking=True)
        self.events[i].record(torch.cuda.current_stream(self.device))
        return d


_staging = {}


def upload_mask_parameters(p, sigmas, device):
    """Host part of cowmix.py:27-31 and :64: kernel size, taps and erfinv threshold factors, computed
    straight into a pinned staging buffer and uploaded in one non-blocking copy.
    Returns (size, device buffer [N*size taps | N factors]); the buffer belongs to a ring of 8 and is
    overwritten 8 calls later (stream-ordered consumers only)."""
    n = sigmas.shape[0]
    size = kernel_size_for(sigmas.max().item())
    st = _staging.get(device)
    if st is None:
        st = _staging[device] = _Staging(device)
    i = st.slot(n * size + n)
    host = st.host[i]
    gaussian_taps(size, sigmas.float(), out=host[: n * size].view(n, size))
    # cowmix.py:64  torch.erfinv(2 * p - 1) * math.sqrt(2.0)
    torch.mul(p.mul(2).sub_(1).erfinv_(), _SQRT2, out=host[n * size: n * size + n])
    return size, st.upload(i, n * size + n)


def stage_mask_parameters(p, sigmas, taps_dev):
    """As upload_mask_parameters, but into a caller-owned device buffer [>= N*K + N floats] that keeps its
    address from step to step (LossPathStep's static / CUDA-graph mode): the copy is ordered on the current
    stream behind the previous step's readers of the buffer.  Returns the kernel size K."""
    n = sigmas.shape[0]
    size = kernel_size_for(sigmas.max().item())
    nf = n * size + n
    if taps_dev.numel() < nf:
        raise ValueError(f"taps buffer holds {taps_dev.numel()} floats, K={size} needs {nf} "
                         "(sigma outside the configured sigma_range?)")
    device = taps_dev.device
    st = _staging.get(device)
    if st is None:
        st = _staging[device] = _Staging(device)
    i = st.slot(nf)
    host = st.host[i]
    gaussian_taps(size, sigmas.float(), out=host[: n * size].view(n, size))
    torch.mul(p.mul(2).sub_(1).erfinv_(), _SQRT2, out=host[n * size: nf])
    taps_dev[:nf].copy_(host[:nf], non_blocking=True)
    st.events[i].record(torch.cuda.current_stream(device))
    return size


def masks_from_noise(noise, p, sigmas, return_field=False):
    """Deterministic tail of generate_cowmix_masks_like (cowmix.py:56-68) for a given noise field.

    noise: [N,1,H,W] fp32 CUDA; p, sigmas: [N] CPU tensors.  Returns mask [N,1,H,W] (and the
    smoothed field S when return_field is set).
    """
    require_cuda(noise, "noise", torch.float32)
    if noise.dim() != 4 or noise.shape[1] != 1:
        raise ValueError("noise must be [N,1,H,W]")
    n, _, h, w = noise.shape
    assert n == sigmas.shape[0]  # cowmix.py:29
    noise = noise.contiguous()
    mask = torch.empty_like(noise)
    field = torch.empty_like(noise) if return_field else None
    if n == 0 or h == 0 or w == 0:
        return (mask, field) if return_field else mask
    size, dev = upload_mask_parameters(p, sigmas, noise.device)
    ws_bytes = lib.b200ssl_cowmix_workspace_bytes(n, h, w)
    ws = _lib.workspaces.get(noise.device, "cowmix", ws_bytes)
    with torch.cuda.device(noise.device):
        check(lib.b200ssl_cowmix_mask(
            noise.data_ptr(), dev.data_ptr(), size, dev.data_ptr() + 4 * n * size, n, h, w,
            mask.data_ptr(), field.data_ptr() if return_field else None, ws.data_ptr(), ws.numel(),
            stream_ptr(noise.device)), "cowmix_mask")
    return (mask, field) if return_field else mask


def generate_cowmix_masks_like(example_tensor, mask_proportion_range, sigma_range):
    # mask_proportion range: tuple of 2 python floats; sigma_range: tuple of 2 python floats
    require_cuda(example_tensor, "example_tensor")
    if example_tensor.dtype != torch.float32:
        raise TypeError("b200ssl.cowmix: only float32 is supported (the reference trains in fp32)")
    with torch.no_grad():
        n = example_tensor.size(0)
        p, sigmas = draw_mask_parameters(n, mask_proportion_range, sigma_range)
        size = list(example_tensor.size())
        size[1] = 1
        # device generator, drawn after the CPU draws exactly like cowmix.py:53-55
        noise = torch.normal(mean=0, std=1, size=size, dtype=example_tensor.dtype,
                             device=example_tensor.device)
        return masks_from_noise(noise, p, sigmas)


def _as_nchw(t):
    if t.dim() < 2:
        raise ValueError("mix_with_mask expects tensors with a batch and a channel dimension")
    return t.contiguous()


def mix2_with_mask(a0, b0, a1, b1, mask):
    """Fused `mix_with_mask(a0,b0,mask), mix_with_mask(a1,b1,mask)` (train.py:82-86) in one launch.
    a*: [N,C*,H,W]; mask: [N,1,H,W].  Pass a1=b1=None to mix a single pair."""
    require_cuda(mask, "mask", torch.float32)
    a0 = _as_nchw(require_cuda(a0, "tensor_a", torch.float32))
    b0 = _as_nchw(require_cuda(b0, "tensor_b", torch.float32))
    if a0.shape != b0.shape:
        raise ValueError("tensor_a and tensor_b must have the same shape")
    n, c0 = a0.shape[0], a0.shape[1]
    hw = a0[0, 0].numel() if a0.numel() else 0
    mask = mask.contiguous()
    per_channel = mask.shape == a0.shape and c0 != 1
    if not per_channel and (mask.shape[0] != n or mask.shape[1] != 1 or mask[0, 0].numel() != hw):
        raise ValueError(f"mask shape {tuple(mask.shape)} does not broadcast over {tuple(a0.shape)} as [N,1,H,W]")
    out0 = torch.empty_like(a0)
    out1 = None
    c1 = 0
    if a1 is not None:
        if per_channel:
            raise ValueError("a per-channel mask cannot be shared with a second tensor pair")
        a1 = _as_nchw(require_cuda(a1, "tensor_a1", torch.float32))
        b1 = _as_nchw(require_cuda(b1, "tensor_b1", torch.float32))
        if a1.shape != b1.shape or a1.shape[0] != n:
            raise ValueError("second tensor pair must share the batch size with the first")
        c1 = a1.shape[1]
        if a1.dim() == 4 and a0.dim() == 4 and a1.shape[2:] != a0.shape[2:]:
            # row N2: the second pair is at a lower resolution (teacher logits before train.py:72-75's
            # F.interpolate): it is up-sampled bilinearly (align_corners=False) inside the mix
            h, w = a0.shape[2], a0.shape[3]
            out1 = torch.empty((n, c1, h, w), dtype=torch.float32, device=a0.device)
            with torch.cuda.device(a0.device):
                check(lib.b200ssl_mix2_upsampled(
                    a0.data_ptr(), b0.data_ptr(), out0.data_ptr(), c0, a1.data_ptr(), b1.data_ptr(), out1.data_ptr(),
                    c1, a1.shape[2], a1.shape[3], mask.data_ptr(), None, None, n, h, w, stream_ptr(a0.device)),
                    "mix2_upsampled")
            return out0, out1
        if (a1[0, 0].numel() if a1.numel() else 0) != hw:
            raise ValueError("second tensor pair must share batch and spatial size with the first")
        out1 = torch.empty_like(a1)
    with torch.cuda.device(a0.device):
        check(lib.b200ssl_mix2(
            a0.data_ptr(), b0.data_ptr(), out0.data_ptr(), c0,
            a1.data_ptr() if c1 else None, b1.data_ptr() if c1 else None,
            out1.data_ptr() if c1 else None, c1, mask.data_ptr(), c0 if per_channel else 1,
            n, hw, stream_ptr(a0.device)), "mix2")
    return out0, out1


def upsample_bilinear(x, size):
    """F.interpolate(x, size, mode='bilinear', align_corners=False) for no-grad consumers (train.py:72-75:
    teacher predictions): bit-identical to ATen for out >= in.  x: [N,C,h,w] fp32 CUDA."""
    x = require_cuda(x, "x", torch.float32).contiguous()
    if x.dim() != 4:
        raise ValueError("upsample_bilinear expects [N,C,h,w]")
    h, w = int(size[0]), int(size[1])
    out = torch.empty((x.shape[0], x.shape[1], h, w), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.b200ssl_upsample_bilinear(x.data_ptr(), x.shape[0] * x.shape[1], x.shape[2], x.shape[3],
                                            out.data_ptr(), h, w, stream_ptr(x.device)), "upsample_bilinear")
    return out


class _Mix(torch.autograd.Function):
    """Autograd-transparent like the reference's arithmetic: d/da = mask, d/db = 1 - mask."""

    @staticmethod
    def forward(ctx, tensor_a, tensor_b, mask):
        ctx.save_for_backward(mask)
        return mix2_with_mask(tensor_a, tensor_b, None, None, mask)[0]

    @staticmethod
    def backward(ctx, grad):
        (mask,) = ctx.saved_tensors
        zero = torch.zeros_like(grad)
        ga = mix2_with_mask(grad, zero, None, None, mask)[0] if ctx.needs_input_grad[0] else None
        gb = mix2_with_mask(zero, grad, None, None, mask)[0] if ctx.needs_input_grad[1] else None
        return ga, gb, None


def mix_with_mask(tensor_a, tensor_b, mask):
    """cowmix.py:72-73: tensor_a * mask + tensor_b * (1. - mask), bit-identical evaluation order."""
    if torch.is_grad_enabled() and (tensor_a.requires_grad or tensor_b.requires_grad):
        if mask.requires_grad:
            raise NotImplementedError("b200ssl.cowmix.mix_with_mask: gradient w.r.t. the mask is not supported")
        return _Mix.apply(tensor_a, tensor_b, mask)
    return mix2_with_mask(tensor_a, tensor_b, None, None, mask)[0]
