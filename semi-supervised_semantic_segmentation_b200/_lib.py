"""ctypes binding of libb200ssl.so (the C ABI declared in include/b200ssl.h).

The product path has no CPU fallback: if the CUDA library has not been built this module
raises at import time, and every op raises if it is handed a non-CUDA tensor.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libb200ssl.so")

MAX_LIST = 64
I64, I32, U8 = 0, 1, 2
LOVASZ_ALL, LOVASZ_PRESENT, LOVASZ_LIST = 0, 1, 2
LOVASZ_ERR_ABS, LOVASZ_ERR_HINGE = 0, 1
EMA_CHUNK = 4096


class LovaszDesc(C.Structure):
    _fields_ = [
        ("n_images", C.c_int32),
        ("n_channels", C.c_int32),
        ("hw", C.c_int64),
        ("per_image", C.c_int32),
        ("class_mode", C.c_int32),
        ("n_list", C.c_int32),
        ("class_list", C.c_int32 * MAX_LIST),
        ("has_ignore", C.c_int32),
        ("ignore_index", C.c_int64),
        ("label_dtype", C.c_int32),
        ("error_mode", C.c_int32),
    ]


class EmaChunk(C.Structure):
    _fields_ = [("ema", C.c_void_p), ("param", C.c_void_p), ("count", C.c_int32), ("pad_", C.c_int32)]


class SgdChunk(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("momentum", C.c_void_p), ("ema", C.c_void_p),
                ("count", C.c_int32), ("tensor", C.c_int32)]


class SgdHyper(C.Structure):
    _fields_ = [("lr", C.c_double), ("momentum", C.c_double), ("dampening", C.c_double), ("weight_decay", C.c_double),
                ("ema_alpha", C.c_double), ("nesterov", C.c_int32), ("first_step", C.c_int32), ("zero_grad", C.c_int32),
                ("reserved_", C.c_int32)]


_vp, _i, _i64, _sz, _d = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_double

STEP_BINARY, STEP_SOFTMAX = 0, 1
STEP_SERIAL, STEP_PREFORKED, STEP_ISSUE_SIDE, STEP_ISSUE_MAIN = 1, 2, 4, 8


class StepDesc(C.Structure):
    """b200ssl_step_desc (include/b200ssl.h)."""
    _fields_ = (
        [(k, C.c_int32) for k in ("n", "classes", "h", "w", "image_channels", "K", "mode", "cm_has_ignore")] +
        [("cm_ignore_index", C.c_int64), ("cm_label_dtype", C.c_int32), ("flags", C.c_int32),
         ("lovasz", LovaszDesc)] +
        [(k, C.c_void_p) for k in ("noise", "taps", "thr_factor", "image_a", "image_b", "teacher_a", "teacher_b",
                                   "scores", "target", "cm_labels", "mask", "mixed_images", "mixed_teacher",
                                   "grad", "cm", "small", "seg_loss", "seg_fg", "seg_valid",
                                   "nonzero", "labels_u8")] +
        [("ws_cowmix", C.c_void_p), ("ws_cowmix_bytes", C.c_size_t), ("ws_lovasz", C.c_void_p),
         ("ws_lovasz_bytes", C.c_size_t), ("ema_table", C.c_void_p), ("ema_entries", C.c_int64),
         ("ema_alpha", C.c_double), ("peer", C.c_void_p), ("peer_cm_out", C.c_void_p),
         ("peer_loss_out", C.c_void_p), ("teacher_h", C.c_int32), ("teacher_w", C.c_int32),
         ("grad_out", C.c_void_p)])


# name -> (restype, argtypes); mirrors include/b200ssl.h one to one
SIGNATURES = {
    "b200ssl_version": (_i, []),
    "b200ssl_last_error": (C.c_char_p, []),
    "b200ssl_launch_count": (C.c_longlong, []),
    "b200ssl_prof_enable": (None, [_i]),
    "b200ssl_prof_report": (C.c_longlong, [C.c_char_p, _sz]),
    "b200ssl_ema_table_entries": (_i64, [_vp, _i]),
    "b200ssl_ema_build_table_host": (_i64, [_vp, _vp, _vp, _i, _vp, _i64]),
    "b200ssl_ema_multi": (_i, [_vp, _i64, _d, _vp]),
    "b200ssl_mix2": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _i, _i64, _i64, _vp]),
    "b200ssl_cowmix_workspace_bytes": (_sz, [_i, _i, _i]),
    "b200ssl_cowmix_mask": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "b200ssl_cowmix_field": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "b200ssl_mix2_field": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i64, _i64, _vp]),
    "b200ssl_lovasz_num_segments": (C.c_int32, [C.POINTER(LovaszDesc)]),
    "b200ssl_lovasz_workspace_bytes": (_sz, [C.POINTER(LovaszDesc)]),
    "b200ssl_lovasz_forward": (_i, [C.POINTER(LovaszDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200ssl_lovasz_forward_backward": (_i, [C.POINTER(LovaszDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200ssl_lovasz_seg_scale": (_i, [C.POINTER(LovaszDesc), _vp, _vp, _vp, _vp, _vp]),
    "b200ssl_lovasz_backward": (_i, [C.POINTER(LovaszDesc), _vp, _vp, _vp, _vp]),
    "b200ssl_binary_lovasz_fused": (_i, [_vp, _vp, _i, _i, _i64, _i] + [_vp] * 10 + [_i, _i64, _vp, _sz, _vp]),
    "b200ssl_argmax_channels": (_i, [_vp, _i, _i, _i64, _vp, _i, _vp, _vp]),
    "b200ssl_binary_lovasz_reduce": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "b200ssl_binary_lovasz_scale": (_i, [_vp, _vp, _vp, _i, _vp, _vp]),
    "b200ssl_confusion_matrix": (_i, [_vp, _vp, _i64, _i, _i, _i, _i64, _i, _i, _i64, _vp, _vp, _vp]),
    "b200ssl_confusion_from_logits": (_i, [_vp, _vp, _i, _i, _i64, _i, _i64, _i, _i, _vp, _vp, _vp]),
    "b200ssl_dice_workspace_bytes": (_sz, [_i, _i64]),
    "b200ssl_dice_metric": (_i, [_vp, _vp, _i, _i64, _vp, _vp, _sz, _vp]),
    "b200ssl_dice_from_cm": (_i, [_vp, _i, _vp, _vp]),
    "b200ssl_validation_cm": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _i, C.c_float, _i, _vp, _vp]),
    "b200ssl_consistency_workspace_bytes": (_sz, [_i, _i64]),
    "b200ssl_consistency_forward": (_i, [_vp, _vp, _i, _i, _i64, C.c_float, _vp, _vp, _sz, _vp]),
    "b200ssl_consistency_backward": (_i, [_vp, _vp, _i, _i, _i64, C.c_float, _vp, _vp, _vp, _vp]),
    "b200ssl_consistency_mixed_workspace_bytes": (_sz, [_i, _i, _i]),
    "b200ssl_consistency_mixed_forward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, C.c_float, _vp, _vp, _vp, _sz, _vp]),
    "b200ssl_consistency_mixed_backward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, C.c_float, _vp, _vp, _vp, _vp, _vp]),
    "b200ssl_loss_path_step": (_i, [C.POINTER(StepDesc), _vp]),
    "b200ssl_loss_path_fork": (_i, [_vp]),
    "b200ssl_sizeof": (_sz, [_i]),
    "b200ssl_softmax_stats": (_i, [_vp, _i, _i, _i64, _vp, _vp, _vp]),
    "b200ssl_lovasz_forward_logits": (_i, [C.POINTER(LovaszDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200ssl_softmax_backward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i64, _vp]),
    "b200ssl_softmax_forward": (_i, [_vp, _i, _i, _i64, _vp, _vp]),
    "b200ssl_softmax_backward_probas": (_i, [_vp, _vp, _i, _i, _i64, _vp]),
    "b200ssl_mix2_upsampled": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i64, _i, _i, _vp]),
    "b200ssl_upsample_bilinear": (_i, [_vp, _i64, _i, _i, _vp, _i, _i, _vp]),
    "b200ssl_upsample_bilinear_backward": (_i, [_vp, _i64, _i, _i, _vp, _i, _i, _vp]),
    "b200ssl_binary_lovasz_lowres": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i] + [_vp] * 10 + [_vp, _sz, _vp]),
    "b200ssl_sgd_build_table_host": (_i64, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _i64]),
    "b200ssl_grad_norm_workspace_bytes": (_sz, [_i64]),
    "b200ssl_grad_norm_multi": (_i, [_vp, _i64, _d, _vp, _vp, _sz, _vp]),
    "b200ssl_grad_scale_multi": (_i, [_vp, _i64, _vp, _vp]),
    "b200ssl_sgd_ema_multi": (_i, [_vp, _i64, _vp, C.POINTER(SgdHyper), _vp]),
    "b200ssl_peer_create": (_i, [_i, _i, C.POINTER(C.c_void_p), _vp]),
    "b200ssl_peer_mailbox": (_vp, [_vp]),
    "b200ssl_peer_connect": (_i, [_vp, _vp]),
    "b200ssl_peer_connect_ptrs": (_i, [_vp, _vp]),
    "b200ssl_peer_post": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp, _vp]),
    "b200ssl_peer_collect": (_i, [_vp, _vp, _vp, _vp]),
    "b200ssl_peer_allreduce": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp, _vp]),
    "b200ssl_peer_status": (_i, [_vp]),
    "b200ssl_peer_destroy": (_i, [_vp]),
}
PEER_HANDLE_BYTES, PEER_MAX_RANKS, PEER_MAX_WORDS, PEER_MAX_FLOATS, PEER_DEPTH = 64, 16, 4096, 8, 4


class B200SSLError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C semi-supervised_semantic_segmentation_b200/csrc` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means header and library are out of sync
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()
for _which, _struct in enumerate((EmaChunk, LovaszDesc, StepDesc, SgdChunk, SgdHyper)):
    if lib.b200ssl_sizeof(_which) != C.sizeof(_struct):
        raise ImportError(f"b200ssl: ctypes layout of {_struct.__name__} ({C.sizeof(_struct)} B) does not match "
                          f"libb200ssl.so ({lib.b200ssl_sizeof(_which)} B); rebuild the library")


def check(rc, what=""):
    if rc != 0:
        msg = lib.b200ssl_last_error().decode("utf-8", "replace")
        raise B200SSLError(f"{what or 'b200ssl'} failed (rc={rc}): {msg}")


def launch_count():
    return int(lib.b200ssl_launch_count())


def kernel_times(enable=None):
    """enable=True/False switches per-kernel event timing; enable=None collects
    {kernel name: (launches, total_ms, min_ms)} since the last collection (synchronises)."""
    if enable is not None:
        lib.b200ssl_prof_enable(1 if enable else 0)
        return None
    buf = C.create_string_buffer(1 << 16)
    lib.b200ssl_prof_report(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, n, total, mn = line.split()
        out[name] = (int(n), float(total), float(mn))
    return out


def stream_ptr(device=None):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t, name, dtype=None):
    """The product path runs on the GPU or not at all."""
    import torch
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(
            f"b200ssl: {name} is on {t.device}; the B200 kernels have no CPU fallback "
            "(use the reference implementation for CPU tensors)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"b200ssl: {name} must be {dtype}, got {t.dtype}")
    return t


def label_dtype_code(t):
    import torch
    if t.dtype == torch.int64:
        return I64
    if t.dtype == torch.int32:
        return I32
    if t.dtype == torch.uint8:
        return U8
    raise TypeError(f"b200ssl: labels must be int64, int32 or uint8, got {t.dtype}")


class WorkspaceCache:
    """One growing scratch tensor per (device, CUDA stream, tag); allocation is PyTorch's (caching
    allocator), the library itself never allocates.  A scratch buffer holds keys, status words and tickets
    of the kernels in flight, so it is only ever shared between launches that are ordered on ONE stream:
    calls issued on different streams get different buffers.  When a buffer is replaced by a larger one
    the old tensor is `record_stream`ed on the stream whose kernels may still be using it before the
    reference is dropped, so the allocator cannot hand it out while they run."""

    def __init__(self):
        self._bufs = {}

    def get(self, device, tag, nbytes):
        import torch
        index = device.index if device.index is not None else torch.cuda.current_device()
        stream = torch.cuda.current_stream(index)
        key = (index, stream.cuda_stream, tag)
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            if buf is not None:
                buf.record_stream(stream)
            buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=torch.device("cuda", index))
            self._bufs[key] = buf
        return buf

    def clear(self):
        self._bufs.clear()


workspaces = WorkspaceCache()
