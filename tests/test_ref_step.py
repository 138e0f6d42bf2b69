"""The staged UNMODIFIED reference (oracle/_ref, oracle/stage_ref.py) driven in train.py order must agree
bit for bit with the restated port (oracle/torch_port.py) on the whole loss-path step: this pins the port --
and through it the C oracle and the CUDA path -- on the reference's own call sequence, and it is what
bench.py's `--impl reference` arm times."""
import numpy as np
import pytest
import torch

from oracle import stage_ref, torch_port


@pytest.fixture(scope="module")
def ref_step():
    stage_ref.stage()
    if not stage_ref.available():
        pytest.skip("oracle/_ref is not staged and /root/reference is not mounted")
    from oracle import ref_step as rs
    return rs


def _inputs(n, c, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    d = dict(image_a=torch.rand(n, 3, h, w, generator=g), image_b=torch.rand(n, 3, h, w, generator=g),
             teacher_a=torch.randn(n, c, h, w, generator=g) * 3, teacher_b=torch.randn(n, c, h, w, generator=g) * 3,
             scores=torch.randn(n, c, h, w, generator=g) * 3)
    lab = torch.nn.functional.avg_pool2d(torch.randn(n, c, h, w, generator=g), 9, 1, 4).argmax(1)
    d["labels"] = lab
    d["target"] = torch.nn.functional.one_hot(lab, c).permute(0, 3, 1, 2).float().contiguous()
    d["params"] = [torch.randn(s, generator=g) for s in [(7,), (33, 5), (4099,), (1,)]]
    d["ema"] = [torch.randn(p.shape, generator=g) for p in d["params"]]
    return d


@pytest.mark.parametrize("mode,c", [("binary", 2), ("softmax", 5)])
def test_port_equals_staged_reference(ref_step, mode, c):
    d = _inputs(2, c, 64, 48, 3)
    outs = []
    for fn in (ref_step.loss_path_step, torch_port.loss_path_step):
        ema = [e.clone() for e in d["ema"]]
        torch.manual_seed(0)
        if mode == "binary":
            scores, target = d["scores"], d["target"]
        else:
            scores, target = torch.softmax(d["scores"], 1), d["labels"]
        o = fn(d["image_a"], d["image_b"], d["teacher_a"], d["teacher_b"], scores, target, d["params"], ema,
               mode=mode, sigma_range=(2, 6), num_classes=c, ignore=255)
        outs.append((o, ema))
    (a, ea), (b, eb) = outs
    for k in ("mask", "mixed_images", "mixed_teacher", "grad", "cm"):
        assert torch.equal(a[k], b[k]), k
    assert float(a["loss"]) == float(b["loss"])
    assert all(torch.equal(x, y) for x, y in zip(ea, eb))


def test_staged_files_are_the_reference(ref_step):
    import json, os, hashlib
    with open(os.path.join(stage_ref.REF_DIR, "MANIFEST.json")) as f:
        man = json.load(f)
    for name, meta in man["files"].items():
        with open(os.path.join(stage_ref.REF_DIR, name), "rb") as f:
            assert hashlib.sha256(f.read()).hexdigest() == meta["sha256"], name
        src = os.path.join(stage_ref.REF_ROOT, name)
        if os.path.exists(src):
            with open(src, "rb") as f:
                assert hashlib.sha256(f.read()).hexdigest() == meta["sha256"], name
