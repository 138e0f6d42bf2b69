"""Host-side cost of one LossPathStep call, by section (queue kept short so the host never blocks on the GPU)."""
import sys, time, collections, torch
sys.path.insert(0, '.')
import bench, b200ssl
from b200ssl import _lib, cowmix
import b200ssl.step as stepmod
dev = torch.device('cuda:0')
inp = bench.make_inputs(dev, 0)
step = b200ssl.LossPathStep(num_classes=2, mode="binary"); step.bind_parameters(inp["params"], inp["ema_params"])
acc = collections.defaultdict(float)
def wrap(obj, name, tag):
    f = getattr(obj, name)
    def g(*a, **k):
        t = time.perf_counter(); r = f(*a, **k); acc[tag] += time.perf_counter() - t; return r
    setattr(obj, name, g)
wrap(cowmix, "draw_mask_parameters", "draw p/sigma (CPU RNG)")
wrap(cowmix, "upload_mask_parameters", "taps + factors + H2D")
wrap(torch, "normal", "torch.normal launch")
wrap(torch, "empty", "torch.empty"); wrap(torch, "empty_like", "torch.empty_like"); wrap(torch, "zeros", "torch.zeros")
class L:  # proxy for the ctypes library
    def __init__(s, lib): s._l = lib
    def __getattr__(s, n):
        f = getattr(s._l, n)
        def g(*a):
            t = time.perf_counter(); r = f(*a); acc["C:" + n] += time.perf_counter() - t; return r
        return g
stepmod.lib = L(_lib.lib)
def one():
    return step(inp["image_a"], inp["image_b"], inp["teacher_a"], inp["teacher_b"], inp["scores"], inp["target"], inp["params"], inp["ema_params"])
for _ in range(20): one()
torch.cuda.synchronize(); acc.clear()
N, tot = 0, 0.0
for rep in range(8):
    t0 = time.perf_counter()
    for _ in range(25): one()
    tot += time.perf_counter() - t0; N += 25
    torch.cuda.synchronize()
print(f"host total {tot / N * 1e6:.1f} us/step")
for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
    print(f"  {v / N * 1e6:7.1f} us  {k}")
print(f"  {(tot - sum(acc.values())) / N * 1e6:7.1f} us  (python glue in step.py: descriptor fields, data_ptr, ...)")
