import sys, torch
sys.path.insert(0, '.')
import b200ssl
from b200ssl import _lib
dev = torch.device('cuda:0')
gen = torch.Generator(device=dev).manual_seed(0)
def coherent(n, c, h, w):
    x = torch.randn(n, c, h // 32, w // 32, device=dev, generator=gen)
    return torch.nn.functional.interpolate(x, size=(h, w), mode="bilinear").argmax(1)
for (n, c, h, w, per_image) in [(4, 21, 512, 512, False), (4, 21, 512, 512, True), (1, 19, 1024, 2048, False)]:
    probas = torch.softmax(torch.randn(n, c, h, w, device=dev, generator=gen) * 2, 1)
    labels = coherent(n, c, h, w)
    labels[coherent(n, 30, h, w) == 0] = 255
    step = b200ssl.LossPathStep(num_classes=c, mode="softmax", classes="present", per_image=per_image, ignore=255)
    for _ in range(3):
        step.lovasz_loss_and_grad(probas, labels)
    torch.cuda.synchronize()
    _lib.kernel_times(True)
    for _ in range(5):
        step.lovasz_loss_and_grad(probas, labels)
    torch.cuda.synchronize()
    kt = _lib.kernel_times(); _lib.kernel_times(False)
    tot = sum(v[1] for v in kt.values()) / 5
    print((n, c, h, w, per_image), "total ms", round(tot, 4), {k: round(v[1] / 5, 4) for k, v in kt.items()})
