"""oracle.stage_ref -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Recipe that makes the UNMODIFIED reference implementation of the loss-and-mixing path available to
bench.py's `--impl reference` arm and `cpu_baseline` leg on the GPU box, where /root/reference does not
exist: the five reference modules on the path are copied byte for byte from the read-only
reference tree into `oracle/_ref/`, which is git-ignored (never part of the history: no reference source
is committed) but NOT gpurun-ignored, so it travels with the snapshot like a built `.so`.
`__graft_entry__.build()` runs this wherever /root/reference is mounted; elsewhere the staged copy is used
as it is, and if it is absent bench.py falls back to the restated port (`kind: "port"`).

    python -m oracle.stage_ref            # stage + print the manifest
"""
import hashlib
import json
import os
import shutil

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"
REF_DIR = os.path.join(_HERE, "_ref")
# module -> the lines of it that are on the path
FILES = {
    "cowmix.py": "generate_gaussian :6-11, gaussian_kernel_2d_vertical :14-24, dual_pass_gaussian_fileter2d :27-37, "
                 "generate_cowmix_masks_like :40-69, mix_with_mask :72-73",
    "lovasz.py": "lovasz_grad :19-31, iou :54-73, lovasz_hinge :79-111, flatten_binary_scores :114-126, "
                 "lovasz_softmax :155-170, lovasz_softmax_flat :173-201, flatten_probas :204-220, mean :235-253",
    "losses.py": "CalculateLoss :8-22, binary_lovasz_loss_with_logits :239-250",
    "mean_teacher.py": "update_ema_variables :5-18, detach_model_parameters :20-22",
    "metrics.py": "dice_metric :1-7",
}


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def stage(src=REF_ROOT, force=False):
    """Copy the on-path reference modules into oracle/_ref/.  Returns the manifest, or None when the
    reference tree is not mounted (the GPU box)."""
    if not os.path.isdir(src):
        return None
    os.makedirs(REF_DIR, exist_ok=True)
    manifest = {"source": src, "files": {}}
    for name, what in FILES.items():
        s, d = os.path.join(src, name), os.path.join(REF_DIR, name)
        if force or not os.path.exists(d) or _sha(s) != _sha(d):
            shutil.copyfile(s, d)
        manifest["files"][name] = {"sha256": _sha(d), "on_path": what}
    with open(os.path.join(REF_DIR, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    return manifest


def available():
    return all(os.path.exists(os.path.join(REF_DIR, n)) for n in FILES)


if __name__ == "__main__":
    print(json.dumps(stage(force=True), indent=1))
