"""The C-ABI library loads on a machine without a GPU and exports every symbol the header declares.
No compute calls here: only argument validation paths that return before any CUDA call, host-side
table builders and size queries."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200ssl.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200ssl_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from b200ssl import _lib
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} is declared in include/b200ssl.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in _lib.py"
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"
    assert _lib.lib.b200ssl_version() == 100


def test_only_the_c_abi_is_exported():
    import subprocess
    from b200ssl import _lib
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    syms = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert syms and all(s.startswith("b200ssl_") for s in syms), syms
    assert set(declared_symbols()) <= set(syms)


def test_struct_layouts_match_the_header():
    from b200ssl import _lib
    assert C.sizeof(_lib.EmaChunk) == 24
    # int32 x2, int64, int32 x3, int32[64], int32, (pad) int64, int32 x2
    assert C.sizeof(_lib.LovaszDesc) == 8 + 8 + 12 + 256 + 4 + 8 + 8
    assert _lib.LovaszDesc.ignore_index.offset % 8 == 0
    for which, st in enumerate((_lib.EmaChunk, _lib.LovaszDesc, _lib.StepDesc)):
        assert _lib.lib.b200ssl_sizeof(which) == C.sizeof(st)
    assert _lib.lib.b200ssl_sizeof(99) == 0


def test_argument_errors_without_gpu():
    from b200ssl import _lib
    lib = _lib.lib
    assert lib.b200ssl_mix2(None, None, None, 3, None, None, None, 0, None, 1, -1, 10, None) == -1
    assert b"negative" in lib.b200ssl_last_error()
    assert lib.b200ssl_mix2(None, None, None, 3, None, None, None, 0, None, 1, 2, 10, None) == -1   # null mask
    assert lib.b200ssl_cowmix_mask(None, None, 4, None, 1, 8, 8, None, None, None, 0, None) == -1   # even K
    assert b"odd" in lib.b200ssl_last_error()
    assert lib.b200ssl_confusion_matrix(None, None, 10, 0, 0, 0, 0, 0, 0, 0, None, None, None) == -1
    assert lib.b200ssl_ema_multi(None, 5, 0.99, None) == -1
    assert lib.b200ssl_ema_multi(None, 0, 0.99, None) == 0                                          # empty: no-op
    d = _lib.LovaszDesc()
    d.n_images, d.n_channels, d.hw, d.class_mode, d.n_list = 2, 3, 100, _lib.LOVASZ_LIST, 2
    d.class_list[0], d.class_list[1] = 1, 1
    assert lib.b200ssl_lovasz_num_segments(C.byref(d)) == -1 and b"duplicate" in lib.b200ssl_last_error()
    d.class_list[1] = 7
    assert lib.b200ssl_lovasz_num_segments(C.byref(d)) == -1 and b"out of range" in lib.b200ssl_last_error()


def test_size_queries_and_segment_counts():
    from b200ssl import _lib
    lib = _lib.lib
    d = _lib.LovaszDesc()
    d.n_images, d.n_channels, d.hw = 4, 21, 512 * 512
    d.class_mode, d.per_image = _lib.LOVASZ_PRESENT, 0
    assert lib.b200ssl_lovasz_num_segments(C.byref(d)) == 21
    d.per_image = 1
    assert lib.b200ssl_lovasz_num_segments(C.byref(d)) == 84
    ws = lib.b200ssl_lovasz_workspace_bytes(C.byref(d))
    assert ws >= 2 * 84 * 512 * 512 * 8            # two key buffers
    d.class_mode, d.n_list = _lib.LOVASZ_LIST, 1
    d.class_list[0] = 1
    assert lib.b200ssl_lovasz_num_segments(C.byref(d)) == 4
    assert lib.b200ssl_cowmix_workspace_bytes(16, 512, 512) >= 2 * 16 * 512 * 512 * 4
    assert lib.b200ssl_cowmix_workspace_bytes(0, 512, 512) == 0
    assert lib.b200ssl_dice_workspace_bytes(3, 1000) > 0


def test_ema_chunk_table_host_builder():
    """b200ssl_ema_build_table_host is pure host code: tensors are cut into <= 4096-element chunks in
    order, every element is covered exactly once."""
    from b200ssl import _lib
    lib = _lib.lib
    numels = [1, 4096, 4097, 0, 10000]
    n = len(numels)
    arr = (C.c_int64 * n)(*numels)
    entries = lib.b200ssl_ema_table_entries(arr, n)
    assert entries == 1 + 1 + 2 + 0 + 3
    base_e, base_p = 0x10000000, 0x20000000
    offs = [0, 64, 64 + 4096 * 4, 0, 1 << 20]
    e_ptrs = (C.c_void_p * n)(*[base_e + o for o in offs])
    p_ptrs = (C.c_void_p * n)(*[base_p + o for o in offs])
    table = (_lib.EmaChunk * entries)()
    assert lib.b200ssl_ema_build_table_host(e_ptrs, p_ptrs, arr, n, table, entries) == entries
    covered = {}
    for ch in table:
        t = ch.pad_
        assert 1 <= ch.count <= _lib.EMA_CHUNK
        assert (ch.ema - base_e) == (ch.param - base_p)
        start = (ch.ema - base_e - offs[t]) // 4
        covered.setdefault(t, []).append((start, ch.count))
    for t, numel in enumerate(numels):
        spans = sorted(covered.get(t, []))
        pos = 0
        for s, c in spans:
            assert s == pos
            pos += c
        assert pos == numel
    assert lib.b200ssl_ema_build_table_host(e_ptrs, p_ptrs, arr, n, table, entries - 1) == -2    # too small


def test_product_path_refuses_cpu_tensors_and_never_touches_the_oracle():
    import torch
    import b200ssl
    x = torch.zeros(1, 2, 4, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        b200ssl.cowmix.mix_with_mask(x, x, x[:, :1])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        b200ssl.lovasz.lovasz_softmax(x, torch.zeros(1, 4, 4, dtype=torch.long))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        b200ssl.mean_teacher.EmaUpdater()([x], [x], 0.99)
    with pytest.raises(RuntimeError):
        b200ssl.metrics.confusion_matrix(torch.zeros(4, dtype=torch.long), torch.zeros(4, dtype=torch.long), 2)
    pkg = os.path.join(ROOT, "semi-supervised_semantic_segmentation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "liboracle" not in src, f


def test_optim_and_peer_refuse_cpu_and_bad_arguments():
    """Rows N4 / 8e host logic without a GPU: constructor validation mirrors torch.optim.SGD, CPU tensors
    are refused (no fallback), and a peer exchange that cannot fit its mailbox row is rejected up front."""
    import torch
    import b200ssl
    p = torch.nn.Parameter(torch.zeros(4))
    with pytest.raises(ValueError):
        b200ssl.optim.FusedSGD([p], lr=-1.0)
    with pytest.raises(ValueError):
        b200ssl.optim.FusedSGD([p], lr=0.1, momentum=0.0, nesterov=True)
    opt = b200ssl.optim.FusedSGD([p], lr=0.1, momentum=0.9)
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError):
        opt.step()
    with pytest.raises(RuntimeError):
        b200ssl.optim.clip_grad_norm_([p], 1.0)
    with pytest.raises(NotImplementedError):
        b200ssl.optim.clip_grad_norm_([p], 1.0, norm_type=1.0)
    with pytest.raises(ValueError):
        b200ssl.utils.PeerAllReduce(3000, 1, "cuda:0")          # 2*3000+1 words > one mailbox row
    with pytest.raises(ValueError):
        b200ssl.utils.StepReducer(2, 1, torch.device("cpu"), backend="nvlink")
    with pytest.raises(RuntimeError):
        b200ssl.cowmix.upsample_bilinear(torch.zeros(1, 1, 4, 4), (8, 8))
    with pytest.raises(RuntimeError):
        b200ssl.lovasz.lovasz_softmax_with_logits(torch.zeros(1, 3, 4, 4), torch.zeros(1, 4, 4, dtype=torch.int64))
