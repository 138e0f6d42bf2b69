"""Experiment: class-group size of the multi-class key-build (B200SSL_KEY_GROUP / B200SSL_KEY_MULTI)."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import b200ssl
sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
import kernels as K
dev = torch.device("cuda:0"); gen = torch.Generator(device=dev).manual_seed(0)
n, c, h, w = 4, 21, 512, 512
logits = torch.randn(n, c, h, w, device=dev, generator=gen) * 2
probas = torch.softmax(logits, 1)
labels = K.coherent(n, c, h, w, gen); labels[torch.rand(n, h, w, device=dev, generator=gen) < 0.03] = 255
step = b200ssl.LossPathStep(num_classes=c, mode="softmax", classes="present", per_image=False, ignore=255)
ms_p = K.timeit(lambda: step.lovasz_loss_and_grad(probas, labels), reps=5, inner=4)
def fused():
    x = logits.detach().requires_grad_(True)
    b200ssl.lovasz.lovasz_softmax_with_logits(x, labels, ignore=255).backward()
ms_f = K.timeit(fused, reps=5, inner=4)
b200ssl._lib.kernel_times(True)
for _ in range(5): fused(); step.lovasz_loss_and_grad(probas, labels)
torch.cuda.synchronize()
kt = b200ssl._lib.kernel_times()
print(json.dumps({"group": os.environ.get("B200SSL_KEY_GROUP"), "multi": os.environ.get("B200SSL_KEY_MULTI"),
                  "probas_ms": round(ms_p, 4), "from_logits_ms": round(ms_f, 4),
                  "kernels_us": {k: round(v[1] / v[0] * 1e3, 1) for k, v in kt.items()}}))
