"""The whole loss path of one semi-supervised step as a single call (what bench.py times).

Order (train.py:65-130):
    mask   = cowmix.generate_cowmix_masks_like(image_a, p_range, sigma_range)      train.py:77-80
    mixed  = mix_with_mask(teacher_a, teacher_b, mask), mix_with_mask(image_a, image_b, mask)  :82-86
    loss, dloss/dlogits = Lovasz (binary shim losses.py:239-250, or lovasz_softmax)          :51,:61
    update_ema_variables(model, ema_model, alpha)                                             :130
    confusion matrix of (labels, argmax logits)                                               (new)

Host work per step is what the reference also does on the host (2N uniform draws, N*K taps, N erfinv
factors) plus ONE call into the C library (b200ssl_loss_path_step), which issues every kernel of the
path back to back on the current stream.  Nothing in here synchronises with the host; results are
device tensors.

Three ways to run it, same kernels and same bits:
    LossPathStep(...)                       fresh output tensors every call (safe default)
    LossPathStep(..., static_outputs=True)  outputs come from a ring of `ring` preallocated sets owned by
                                            the object (a set is overwritten `ring` calls later): no
                                            allocator call on the step
    LossPathStep(..., graph=True)           static outputs + the launch sequence (device noise draw,
                                            ~12 kernels on three forked streams, memsets, the peer
                                            exchange) captured ONCE per (tap count K, ring slot, input
                                            addresses) in a CUDA graph and replayed: the host then only draws
                                            p/sigma, builds the taps, copies them (one 12 KB H2D) and launches
                                            one graph.  Pays off when the caller feeds the step from static
                                            buffers (as CUDA-graph users do); inputs at new addresses are
                                            captured again (LRU of `max_graphs`).
"""
import ctypes as C
from collections import OrderedDict

import torch

from . import _lib
from ._lib import lib, check, require_cuda
from . import cowmix, lovasz, mean_teacher


class LossPathStep:
    def __init__(self, num_classes, mask_proportion_range=(0.45, 0.55), sigma_range=(8, 32),
                 ema_alpha=0.99, mode="binary", classes="present", per_image=False, ignore=255, serial=False,
                 peer=None, static_outputs=False, graph=False, ring=2, max_graphs=96):
        if mode not in ("binary", "softmax"):
            raise ValueError("mode must be 'binary' (losses.binary_lovasz_loss_with_logits) or 'softmax'")
        self.num_classes = num_classes
        self.mask_proportion_range = mask_proportion_range
        self.sigma_range = sigma_range
        self.ema_alpha = ema_alpha
        self.mode = mode
        self.classes = classes
        self.per_image = per_image
        self.ignore = ignore
        self.serial = serial      # True: keep every kernel on the current stream (no internal fork/join)
        # utils.PeerAllReduce(num_classes**2, 1, device): every step then ends with the exchange of
        # [confusion matrix || loss] over NVLink peer memory, posted by the block that finalises the loss; the
        # same block completes the PREVIOUS step's exchange, so out["cm_sum"] / out["loss_sum"] of a step are
        # valid after the next step or after peer.result()
        self.peer = peer
        self.graph = bool(graph)
        self.static_outputs = bool(static_outputs) or self.graph
        self.ring = max(int(ring), 1)
        self.max_graphs = max_graphs
        self._ema = mean_teacher.EmaUpdater()
        self._scratch_key = None
        self._scratch = None
        self._one = None
        self._bound = None          # (params list, ema list, table pointer, entries)
        self._desc_key = None
        self._desc = None
        self._n_seg = 0
        self._slots_key = None
        self._slots = []
        self._slot_i = 0
        self._graphs = OrderedDict()
        self._pool = None
        self._warm = set()
        self.graph_captures = 0     # how many graphs have been captured (diagnostics / tests)

    def bind_parameters(self, params, ema_params):
        """Validate the (student, teacher) parameter lists ONCE and keep their chunk table, like an
        optimizer that is constructed over a parameter list: later calls that pass the very same list
        objects skip the per-step pointer sweep (~35 us for a few hundred tensors).  Re-bind after
        anything that re-allocates parameter storage (`model.to(...)`, `load_state_dict(assign=True)`)."""
        table, entries = self._ema.prepare(ema_params, params)
        self._bound = (params, ema_params, table, entries)

    # scratch that never leaves this object is kept across steps.  Its reuse is ordered by the stream the
    # step is issued on, so the key carries the stream: a step issued on another stream gets its own set.
    def _get_scratch(self, scores, desc_l, n_seg, stream_id):
        dev = scores.device
        key = (dev, tuple(scores.shape), n_seg, stream_id if not self.graph else 0)
        if key != self._scratch_key:
            n, c, h, w = scores.shape
            ws_c = lib.b200ssl_cowmix_workspace_bytes(n, h, w)
            ws_l = lib.b200ssl_lovasz_workspace_bytes(C.byref(desc_l))
            if self._scratch is not None:
                for t in self._scratch.values():
                    t.record_stream(torch.cuda.current_stream(dev))
            self._scratch = {
                "segf": torch.empty(max(n_seg, 1), dtype=torch.float32, device=dev),          # seg_loss
                "segi": torch.empty(2 * max(n_seg, 1) + n, dtype=torch.int32, device=dev),    # seg_fg | seg_valid | nonzero
                "ws_c": torch.empty(max(ws_c, 256), dtype=torch.uint8, device=dev),
                "ws_l": torch.empty(max(ws_l, 256), dtype=torch.uint8, device=dev),
            }
            self._one = torch.ones(1, dtype=torch.float32, device=dev)     # constant upstream gradient
            self._scratch_key = key
            self._graphs.clear()
        return self._scratch

    def _lovasz_desc(self, scores, target):
        if self.mode == "binary":
            d = _lib.LovaszDesc()
            d.n_images, d.n_channels, d.hw = scores.shape[0], scores.shape[1], scores.shape[2] * scores.shape[3]
            d.per_image, d.class_mode, d.n_list = 1, _lib.LOVASZ_LIST, 1
            d.class_list[0] = 1
            d.has_ignore, d.ignore_index, d.label_dtype = 1, 255, _lib.U8
            return d
        return lovasz._make_desc(scores, target, self.classes, self.per_image, self.ignore)

    def lovasz_loss_and_grad(self, scores, target):
        """Lovasz loss and its gradient w.r.t. the scores only (no mask / mix / EMA / matrix)."""
        out = self._run(None, None, None, None, scores, target, None, None, None, None, want_cm=False)
        return out["loss"], out["grad"], out["labels"]

    def __call__(self, image_a, image_b, teacher_a, teacher_b, scores, target, params, ema_params,
                 cm_labels=None, cm_out=None):
        """scores: student logits (binary mode) or probabilities (softmax mode) [N,C,H,W];
        target: soft one-hot [N,C,H,W] (binary mode) or integer labels [N,H,W] (softmax mode);
        params / ema_params: lists of student / teacher parameter tensors;
        cm_labels: integer labels for the confusion matrix (defaults to the Lovasz labels);
        cm_out: int64 [C,C] matrix to ACCUMULATE into (default: a zeroed matrix per step)."""
        return self._run(image_a, image_b, teacher_a, teacher_b, scores, target, params, ema_params,
                         cm_labels, cm_out, want_cm=True)

    # ---- output sets ------------------------------------------------------------------------------
    def _new_outputs(self, dev, n, c, h, w, img_c, has_teacher, binary, want_cm, need_noise):
        o = {"small": torch.empty(4, dtype=torch.float32, device=dev),        # [loss, denom, -, -]
             "grad": torch.empty((n, c, h, w), dtype=torch.float32, device=dev)}
        if binary:
            o["labels"] = torch.empty((n, h, w), dtype=torch.uint8, device=dev)
        if want_cm:
            o["cm"] = torch.zeros((c, c), dtype=torch.int64, device=dev)
            if self.peer is not None:
                o["cm_sum"] = torch.zeros(c * c, dtype=torch.int64, device=dev)
                o["loss_sum"] = torch.zeros(1, dtype=torch.float64, device=dev)
        if img_c:
            o["mask"] = torch.empty((n, 1, h, w), dtype=torch.float32, device=dev)
            o["mixed_images"] = torch.empty((n, img_c, h, w), dtype=torch.float32, device=dev)
            if has_teacher:
                o["mixed_teacher"] = torch.empty((n, c, h, w), dtype=torch.float32, device=dev)
            if need_noise:
                o["noise"] = torch.empty((n, 1, h, w), dtype=torch.float32, device=dev)
        return o

    def _next_slot(self, key, make):
        if key != self._slots_key:
            self._slots = [make() for _ in range(self.ring)]
            self._slots_key = key
            self._slot_i = 0
            self._graphs.clear()
            # taps [n*K | n factors] for the largest K of the sigma range, refreshed by one H2D copy per step
            n = key[1][0]
            kmax = cowmix.kernel_size_for(float(self.sigma_range[1])) + 2
            self._taps_dev = torch.empty(n * kmax + n, dtype=torch.float32, device=key[0])
        i = self._slot_i
        self._slot_i = (i + 1) % self.ring
        return i, self._slots[i]

    # ---- one step ---------------------------------------------------------------------------------
    def _run(self, image_a, image_b, teacher_a, teacher_b, scores, target, params, ema_params,
             cm_labels, cm_out, want_cm):
        with torch.no_grad():
            require_cuda(scores, "scores", torch.float32)
            dev = scores.device
            scores = scores.contiguous()
            target = target.contiguous()
            n, c, h, w = scores.shape
            binary = self.mode == "binary"
            if binary:
                require_cuda(target, "target", torch.float32)
                if target.shape != scores.shape:
                    raise ValueError("binary mode: target must be a soft one-hot tensor shaped like the scores")
            else:
                require_cuda(target, "target")
            # the descriptor is rebuilt only when the problem shape changes; per step only pointers move
            dkey = (tuple(scores.shape), target.dtype, dev)
            if dkey != self._desc_key:
                desc_l = self._lovasz_desc(scores, target)
                n_seg = lib.b200ssl_lovasz_num_segments(C.byref(desc_l))
                if n_seg < 0:
                    check(n_seg, "lovasz_num_segments")
                d = _lib.StepDesc()
                d.n, d.classes, d.h, d.w = n, c, h, w
                d.mode = _lib.STEP_BINARY if binary else _lib.STEP_SOFTMAX
                d.lovasz = desc_l
                self._desc, self._desc_key, self._n_seg = d, dkey, n_seg
            d, n_seg = self._desc, self._n_seg
            C.memset(C.byref(d, _lib.StepDesc.noise.offset), 0, C.sizeof(d) - _lib.StepDesc.noise.offset)  # pointers
            d.flags = _lib.STEP_SERIAL if self.serial else 0
            d.K = d.image_channels = d.cm_has_ignore = d.cm_label_dtype = 0
            d.cm_ignore_index = 0
            tstream = torch.cuda.current_stream(dev)
            sc = self._get_scratch(scores, d.lovasz, n_seg, tstream.cuda_stream)
            ns = max(n_seg, 1)
            stream = C.c_void_p(tstream.cuda_stream)
            current = torch.cuda.current_device() == dev.index     # the library launches on the current device
            has_img = image_a is not None
            has_teacher = has_img and teacher_a is not None
            if has_img:
                require_cuda(image_a, "image_a", torch.float32)
                image_a, image_b = image_a.contiguous(), image_b.contiguous()
                if has_teacher:
                    teacher_a, teacher_b = teacher_a.contiguous(), teacher_b.contiguous()
                    if teacher_a.shape != teacher_b.shape or teacher_a.shape[:2] != (n, c):
                        raise ValueError("teacher predictions must both be [N, num_classes, h, w]")
            img_c = image_a.shape[1] if has_img else 0
            # ---- outputs: fresh tensors, or the next set of the ring
            if self.static_outputs:
                skey = (dev, (n, c, h, w), img_c, has_teacher, binary, want_cm, self.peer is not None)
                slot, o = self._next_slot(skey, lambda: self._new_outputs(dev, n, c, h, w, img_c, has_teacher, binary,
                                                                          want_cm, True))
                if want_cm and cm_out is None and not self.graph:
                    o["cm"].zero_()
            else:
                slot, o = 0, self._new_outputs(dev, n, c, h, w, img_c, has_teacher, binary, want_cm, False)
            out = {}
            # Everything the Lovasz / matrix / EMA chains touch is initialised FIRST, so that the fork point
            # can be recorded before the mask parameters and the noise are produced.
            # ---- Lovasz
            small, grad = o["small"], o["grad"]
            d.scores, d.target = scores.data_ptr(), target.data_ptr()
            d.grad, d.small, d.grad_out = grad.data_ptr(), small.data_ptr(), self._one.data_ptr()
            d.seg_loss = sc["segf"].data_ptr()
            d.seg_fg, d.seg_valid = sc["segi"].data_ptr(), sc["segi"].data_ptr() + 4 * ns
            d.nonzero = sc["segi"].data_ptr() + 8 * ns
            d.ws_lovasz, d.ws_lovasz_bytes = sc["ws_l"].data_ptr(), sc["ws_l"].numel()
            if binary:
                labels = o["labels"]
                d.labels_u8 = labels.data_ptr()
            else:
                labels = target
            # ---- confusion matrix
            own_cm = False
            if want_cm:
                if cm_out is None:
                    cm_out, own_cm = o["cm"], True
                d.cm = cm_out.data_ptr()
                d.cm_has_ignore = 0 if self.ignore is None else 1
                d.cm_ignore_index = 0 if self.ignore is None else int(self.ignore)
                if cm_labels is not None:
                    cm_labels = cm_labels.contiguous()
                    d.cm_labels, d.cm_label_dtype = cm_labels.data_ptr(), _lib.label_dtype_code(cm_labels)
                out["cm"] = cm_out
            # ---- EMA
            if params is not None:
                b = self._bound
                if b is not None and params is b[0] and ema_params is b[1]:
                    table, entries = b[2], b[3]
                else:
                    table, entries = self._ema.prepare(ema_params, params)
                if entries:
                    d.ema_table, d.ema_entries, d.ema_alpha = table, entries, float(self.ema_alpha)
            # ---- multi-GPU exchange: this step's sums land in (cm_sum, loss_sum); the previous step's are
            # completed by this step's post
            if self.peer is not None and want_cm:
                if self.peer.n_ints != c * c or self.peer.n_floats != 1:
                    raise ValueError("LossPathStep: peer must be PeerAllReduce(num_classes**2, 1, device)")
                (cm_sum, loss_sum), prev = self.peer.begin_step((o["cm_sum"], o["loss_sum"]) if self.static_outputs else None)
                d.peer = self.peer.handle
                if prev is not None:
                    d.peer_cm_out, d.peer_loss_out = prev[0].data_ptr(), prev[1].data_ptr()
                out["cm_sum"], out["loss_sum"] = cm_sum[:c * c].view(c, c), loss_sum[0]
            # ---- mask + mix (skipped when no images are given)
            use_graph = self.graph
            if has_img:
                if has_teacher:
                    if teacher_a.shape[2:] != (h, w):
                        # row N2: low-resolution teacher logits (train.py:71-75) are up-sampled inside the mix
                        d.teacher_h, d.teacher_w = teacher_a.shape[2], teacher_a.shape[3]
                    mixed_teacher = o["mixed_teacher"]
                    d.teacher_a, d.teacher_b, d.mixed_teacher = teacher_a.data_ptr(), teacher_b.data_ptr(), mixed_teacher.data_ptr()
                    out["mixed_teacher"] = mixed_teacher
                split = not self.serial and not use_graph
                if split:
                    # The Lovasz / matrix / exchange and EMA chains need neither the mask parameters nor the
                    # noise: they are launched NOW (first half of the split step), so the GPU is already busy
                    # while the host draws p / sigma and builds the taps (~0.1 ms) -- this is what a loop that
                    # synchronises every step (loss.item()) sees as latency.
                    d.flags |= _lib.STEP_ISSUE_SIDE
                    if current:
                        check(lib.b200ssl_loss_path_step(C.byref(d), stream), "loss_path_step (side chains)")
                    else:
                        with torch.cuda.device(dev):
                            check(lib.b200ssl_loss_path_step(C.byref(d), stream), "loss_path_step (side chains)")
                    d.flags = (d.flags & ~_lib.STEP_ISSUE_SIDE) | _lib.STEP_ISSUE_MAIN
                try:
                    p, sigmas = cowmix.draw_mask_parameters(n, self.mask_proportion_range, self.sigma_range)
                    if self.static_outputs:
                        size = cowmix.stage_mask_parameters(p, sigmas, self._taps_dev)
                        taps_dev = self._taps_dev
                        noise = o["noise"]
                        if not use_graph:
                            noise.normal_()           # device generator, after the CPU draws (cowmix.py:44-55)
                    else:
                        size, taps_dev = cowmix.upload_mask_parameters(p, sigmas, dev)
                        noise = torch.normal(mean=0, std=1, size=(n, 1, h, w), dtype=torch.float32, device=dev)
                except BaseException:
                    if split:
                        # the side chains of this step are already running: join them back into the caller's
                        # stream (second half with no mask / mix work) before the error leaves this call
                        with torch.cuda.device(dev):
                            lib.b200ssl_loss_path_step(C.byref(d), stream)
                    raise
                d.K, d.image_channels = size, img_c
                d.noise, d.taps, d.thr_factor = noise.data_ptr(), taps_dev.data_ptr(), taps_dev.data_ptr() + 4 * n * size
                d.image_a, d.image_b = image_a.data_ptr(), image_b.data_ptr()
                d.mask, d.mixed_images = o["mask"].data_ptr(), o["mixed_images"].data_ptr()
                out["mask"], out["mixed_images"] = o["mask"], o["mixed_images"]
                d.ws_cowmix, d.ws_cowmix_bytes = sc["ws_c"].data_ptr(), sc["ws_c"].numel()
            if use_graph:
                self._replay(d, dev, o if has_img else None, o["cm"] if own_cm else None, slot)
            elif current:
                check(lib.b200ssl_loss_path_step(C.byref(d), stream), "loss_path_step")
            else:
                with torch.cuda.device(dev):
                    check(lib.b200ssl_loss_path_step(C.byref(d), stream), "loss_path_step")
            out["loss"], out["grad"], out["labels"] = small[0], grad, labels
            return out

    # ---- CUDA graph of the launch sequence ----------------------------------------------------------
    def _replay(self, d, dev, o_img, own_cm, slot):
        """Replay (capturing first if needed) the graph of this step's launch sequence.  A graph is
        identified by every value the launches carry: the raw bytes of the descriptor (extents, K, all
        pointers) and the ring slot."""
        key = (slot, bytes(d))
        g = self._graphs.get(key)
        if g is None:
            shape_key = (d.n, d.classes, d.h, d.w, d.mode, dev.index)
            with torch.cuda.device(dev):
                if shape_key not in self._warm:
                    # one eager pass first, so that lazily created side streams and per-device function
                    # attributes exist before a capture.  It runs on a copy of the descriptor WITHOUT the
                    # stages that have effects beyond the step's own outputs (EMA update, confusion-matrix
                    # accumulation, peer exchange); what it writes is overwritten by the replay.
                    w = _lib.StepDesc.from_buffer_copy(d)
                    w.ema_table, w.ema_entries, w.cm, w.peer = None, 0, None, None
                    w.peer_cm_out = w.peer_loss_out = None
                    check(lib.b200ssl_loss_path_step(C.byref(w), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                          "loss_path_step (warm-up)")
                    self._warm.add(shape_key)
                g = torch.cuda.CUDAGraph()
                if self._pool is None:
                    self._pool = torch.cuda.graph_pool_handle()
                rc = [0]
                with torch.cuda.graph(g, pool=self._pool):
                    if o_img is not None:
                        o_img["noise"].normal_()      # device generator: replays draw fresh numbers (graph-safe philox)
                    if own_cm is not None:
                        own_cm.zero_()
                    d.flags &= ~_lib.STEP_PREFORKED
                    rc[0] = lib.b200ssl_loss_path_step(C.byref(d), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
                check(rc[0], "loss_path_step (graph capture)")
            self._graphs[key] = g
            self.graph_captures += 1
            while len(self._graphs) > self.max_graphs:
                self._graphs.popitem(last=False)
        else:
            self._graphs.move_to_end(key)
        g.replay()
