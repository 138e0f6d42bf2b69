"""Multi-class Lovasz forward+backward only (for ncu captures of its kernels)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import b200ssl
sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
import kernels as K
dev = torch.device("cuda:0"); gen = torch.Generator(device=dev).manual_seed(0)
n, c, h, w = 4, 21, 512, 512
probas = torch.softmax(torch.randn(n, c, h, w, device=dev, generator=gen) * 2, 1)
labels = K.coherent(n, c, h, w, gen); labels[torch.rand(n, h, w, device=dev, generator=gen) < 0.03] = 255
step = b200ssl.LossPathStep(num_classes=c, mode="softmax", classes="present", per_image=False, ignore=255)
for _ in range(4):
    step.lovasz_loss_and_grad(probas, labels)
torch.cuda.synchronize()
print("ok")
