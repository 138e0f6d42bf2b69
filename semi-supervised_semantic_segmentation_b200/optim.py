"""Row N4: the optimiser statements around the loss path, multi-tensor and fused.

    reference train.py:122   torch.nn.utils.clip_grad_norm_(model.module.parameters(), clip_value)
    reference train.py:123   optimizer.step()      torch.optim.SGD(momentum=0.9, weight_decay=5e-4)
    reference train.py:124   optimizer.zero_grad()                    (configs/default_config.py:151-154)
    reference train.py:130   mean_teacher.update_ema_variables(model, ema_model, alpha)

`FusedSGD` keeps torch.optim.SGD's constructor arguments and state layout (`state[p]["momentum_buffer"]`),
and `step()` takes the two neighbours of the optimiser step as options:

    opt = b200ssl.optim.FusedSGD(model.parameters(), lr=..., momentum=0.9, weight_decay=5e-4)
    total_norm = opt.step(max_grad_norm=5.0, ema_params=list(ema_model.parameters()), ema_alpha=0.99,
                          zero_grad=True)

is one reduction + one update launch for all tensors (28-32 B per parameter) instead of ~7 element-wise
passes per tensor.  `clip_grad_norm_` is the stand-alone mirror of the torch function.  The update
arithmetic is torch's op for op (csrc/optim.cu: bit-exact given the same clip coefficient); the total norm
is an fp64 sum of squares in a fixed order, within 1e-6 relative of torch's fp32 norm-of-norms (whose
reduction order is unspecified), so the coefficient can differ from torch's in the last bit.  No CPU fallback.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import lib, check, stream_ptr

_data_ptr = torch.Tensor.data_ptr


class _SgdTable:
    """Device-resident chunk table over (param, grad, momentum, ema) pointers; rebuilt when one moves."""

    def __init__(self):
        self._key = None
        self.table = None
        self.entries = 0
        self.device = None
        self._host = None
        self._ws = None
        self.small = None          # fp32 [2]: total norm, clip coefficient

    def prepare(self, params, grads, moms, emas):
        key = (tuple(map(_data_ptr, params)), tuple(map(_data_ptr, grads)),
               tuple(map(_data_ptr, moms)) if moms is not None else None,
               tuple(map(_data_ptr, emas)) if emas is not None else None)
        if key == self._key:
            return
        device = params[0].device
        for group in (params, grads, moms, emas):
            if group is None:
                continue
            if len(group) != len(params):
                raise ValueError("b200ssl.optim: tensor lists must have equal lengths")
            for t, p in zip(group, params):
                if not t.is_cuda:
                    raise RuntimeError("b200ssl.optim: tensors must live on a CUDA device (no CPU fallback)")
                if t.device != device:
                    raise RuntimeError("b200ssl.optim: all tensors must be on one device")
                if t.dtype != torch.float32:
                    raise TypeError("b200ssl.optim: only float32 tensors are supported")
                if t.shape != p.shape or t.stride() != p.stride() or not (
                        t.is_contiguous() or (t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last))):
                    raise ValueError("b200ssl.optim: gradients / buffers / teacher parameters must be dense and laid "
                                     "out like their parameter")
        n = len(params)
        numels = (C.c_int64 * n)(*[p.numel() for p in params])
        entries = lib.b200ssl_ema_table_entries(numels, n)
        if entries < 0:
            check(int(entries), "ema_table_entries")

        def arr(group):
            return None if group is None else (C.c_void_p * n)(*[t.data_ptr() for t in group])

        host = torch.empty(max(int(entries), 1) * C.sizeof(_lib.SgdChunk), dtype=torch.uint8, pin_memory=True)
        written = lib.b200ssl_sgd_build_table_host(arr(params), arr(grads), arr(moms), arr(emas), numels, n,
                                                   host.data_ptr(), int(entries))
        if written < 0:
            check(int(written), "sgd_build_table_host")
        self.table, self.entries, self._host = host.to(device, non_blocking=True), int(written), host
        self._ws = torch.empty(max(lib.b200ssl_grad_norm_workspace_bytes(self.entries), 256), dtype=torch.uint8,
                               device=device)
        if self.small is None or self.small.device != device:
            self.small = torch.zeros(2, dtype=torch.float32, device=device)
        self._key, self.device = key, device

    def norm(self, max_norm):
        with torch.cuda.device(self.device):
            check(lib.b200ssl_grad_norm_multi(self.table.data_ptr(), self.entries, float(max_norm), self.small.data_ptr(),
                                              self._ws.data_ptr(), self._ws.numel(), stream_ptr(self.device)),
                  "grad_norm_multi")


_clip_tables = {}


def clip_grad_norm_(parameters, max_norm, norm_type=2.0, error_if_nonfinite=False, foreach=None):
    """torch.nn.utils.clip_grad_norm_ (train.py:122) for the L2 norm: returns the total norm (0-dim
    fp32 tensor) and scales every gradient in place by min(max_norm / (norm + 1e-6), 1)."""
    if float(norm_type) != 2.0:
        raise NotImplementedError("b200ssl.optim.clip_grad_norm_: only the L2 norm (the reference's) is implemented")
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    params = [p for p in parameters if p.grad is not None]
    if not params:
        return torch.tensor(0.0)
    with torch.no_grad():
        grads = [p.grad for p in params]
        key = tuple(map(id, params))
        tab = _clip_tables.get(key)
        if tab is None:
            if len(_clip_tables) > 8:
                _clip_tables.clear()
            tab = _clip_tables[key] = _SgdTable()
        tab.prepare(params, grads, None, None)
        tab.norm(max_norm)
        if error_if_nonfinite and not bool(torch.isfinite(tab.small[0])):
            raise RuntimeError("The total norm of order 2.0 for gradients from `parameters` is non-finite, so it "
                               "cannot be clipped.")
        with torch.cuda.device(tab.device):
            check(lib.b200ssl_grad_scale_multi(tab.table.data_ptr(), tab.entries, tab.small.data_ptr() + 4,
                                               stream_ptr(tab.device)), "grad_scale_multi")
        return tab.small[0].clone()


class FusedSGD(torch.optim.Optimizer):
    """torch.optim.SGD's arguments and state; `step` optionally clips before and updates a teacher after."""

    def __init__(self, params, lr=1e-3, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False):
        if lr < 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        if momentum < 0.0:
            raise ValueError(f"Invalid momentum value: {momentum}")
        if weight_decay < 0.0:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")
        super().__init__(params, dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay,
                                      nesterov=nesterov))
        self._tables = [_SgdTable() for _ in self.param_groups]
        self._norm_table = _SgdTable()
        self._cache = {}
        self._frozen_ema = None

    # the cached tensor lists point at state[p]["momentum_buffer"] objects: anything that replaces the
    # state or the groups invalidates them
    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._cache = {}

    def add_param_group(self, param_group):
        super().add_param_group(param_group)
        if hasattr(self, "_tables"):
            self._tables.append(_SgdTable())
            self._cache = {}

    @torch.no_grad()
    def step(self, closure=None, max_grad_norm=None, ema_params=None, ema_alpha=None, zero_grad=False):
        """max_grad_norm: clip_grad_norm_ over ALL parameters of the optimiser first (train.py:122) and
        return the total norm; ema_params (+ ema_alpha): teacher parameters, in the order of the
        optimiser's parameters (pass the same list object every step), updated from the NEW student
        values (train.py:130); zero_grad: write
        zeros into the gradients (optimizer.zero_grad(set_to_none=False))."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        groups = self.param_groups
        single = len(groups) == 1
        if ema_params is not None and ema_alpha is None:
            raise ValueError("FusedSGD.step: ema_params needs ema_alpha")
        # per-group tensor lists are cached; only the data pointers are re-read every step
        plan = []
        offset = 0
        for gi, group in enumerate(groups):
            gparams = group["params"]
            cache = self._cache.get(gi)
            with_grad = tuple(p.grad is not None for p in gparams)
            ema_key = id(ema_params) if ema_params is not None else None
            if cache is None or cache["with_grad"] != with_grad or cache["ema_key"] != ema_key or \
                    cache["n"] != len(gparams):
                params = [p for p in gparams if p.grad is not None]
                moms, fresh = None, None
                if group["momentum"] != 0 and params:
                    fresh = []
                    moms = []
                    for p in params:
                        st = self.state[p]
                        if st.get("momentum_buffer") is None:
                            st["momentum_buffer"] = torch.empty_like(p, memory_format=torch.preserve_format)
                            fresh.append(True)
                        else:
                            fresh.append(False)
                        moms.append(st["momentum_buffer"])
                emas = None
                if ema_params is not None:
                    if not isinstance(ema_params, (list, tuple)):
                        raise TypeError("FusedSGD.step: pass ema_params as a list (it is cached by identity)")
                    if len(ema_params) != sum(len(g["params"]) for g in groups):
                        raise ValueError("FusedSGD.step: ema_params must match the optimiser's parameters")
                    emas = [ema_params[offset + k] for k, p in enumerate(gparams) if p.grad is not None]
                cache = self._cache[gi] = dict(with_grad=with_grad, ema_key=ema_key, n=len(gparams), params=params,
                                               moms=moms, emas=emas, first=bool(fresh) and all(fresh))
                if fresh and any(fresh) and not all(fresh):
                    # some tensors receive their first gradient later than the others (torch.optim.SGD
                    # initialises each buffer on ITS first step): this one step runs as two sub-plans
                    for flag in (True, False):
                        idx = [k for k, f in enumerate(fresh) if f == flag]
                        sub = dict(params=[params[k] for k in idx], moms=[moms[k] for k in idx],
                                   emas=None if emas is None else [emas[k] for k in idx], first=flag)
                        plan.append((group, _SgdTable(), sub))
                    offset += len(gparams)
                    continue
            else:
                cache["first"] = False
            offset += len(gparams)
            if cache["params"]:
                plan.append((group, self._tables[gi], cache))
        coef_ptr, total_norm = None, None
        for group, tab, cache in plan:
            tab.prepare(cache["params"], [p.grad for p in cache["params"]], cache["moms"], cache["emas"])
        if max_grad_norm is not None and plan:
            if single and len(plan) == 1:
                norm_tab = plan[0][1]                       # the update table already lists every gradient
            else:
                norm_tab = self._norm_table
                every = [p for _, _, c in plan for p in c["params"]]
                norm_tab.prepare(every, [p.grad for p in every], None, None)
            norm_tab.norm(max_grad_norm)
            coef_ptr = norm_tab.small.data_ptr() + 4
            total_norm = norm_tab.small[0]
        for group, tab, cache in plan:
            h = _lib.SgdHyper(lr=float(group["lr"]), momentum=float(group["momentum"]),
                              dampening=float(group["dampening"]), weight_decay=float(group["weight_decay"]),
                              ema_alpha=float(ema_alpha) if cache["emas"] is not None else -1.0,
                              nesterov=int(bool(group["nesterov"])), first_step=int(cache["first"]),
                              zero_grad=int(bool(zero_grad)))
            with torch.cuda.device(tab.device):
                check(lib.b200ssl_sgd_ema_multi(tab.table.data_ptr(), tab.entries, coef_ptr, C.byref(h),
                                                stream_ptr(tab.device)), "sgd_ema_multi")
        if ema_params is not None:
            # parameters without a gradient (frozen) are not touched by SGD, but the reference's
            # update_ema_variables (mean_teacher.py:10-11) still moves their teacher copy
            every = [p for g in groups for p in g["params"]]
            frozen = [k for k, p in enumerate(every) if p.grad is None]
            if frozen:
                from . import mean_teacher
                key = (id(ema_params), tuple(frozen))
                if self._frozen_ema is None or self._frozen_ema[0] != key:
                    self._frozen_ema = (key, mean_teacher.EmaUpdater(), [ema_params[k] for k in frozen],
                                        [every[k] for k in frozen])
                _, upd, e_list, p_list = self._frozen_ema
                upd(e_list, p_list, float(ema_alpha))
        if max_grad_norm is not None:
            # a copy: the persistent [norm, coefficient] buffer is overwritten by the next step
            return total_norm.clone() if total_norm is not None else torch.tensor(0.0)
        return loss
