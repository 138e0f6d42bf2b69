import sys, time, torch
sys.path.insert(0, '.')
import bench, b200ssl
from b200ssl import _lib
dev = torch.device('cuda:0')
inp = bench.make_inputs(dev, 0)
W = bench.WORKLOAD
step = b200ssl.LossPathStep(num_classes=2, mode="binary"); step.bind_parameters(inp["params"], inp["ema_params"])
def one():
    return step(inp["image_a"], inp["image_b"], inp["teacher_a"], inp["teacher_b"], inp["scores"], inp["target"], inp["params"], inp["ema_params"])
for _ in range(5): one()
torch.cuda.synchronize()
ts = []
t0 = time.perf_counter()
for i in range(30):
    a = time.perf_counter(); one(); ts.append((time.perf_counter() - a) * 1e3)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host per-step ms:", [round(x, 2) for x in ts])
print("host loop ms/step", (t1 - t0) / 30 * 1e3, "incl. drain", (t2 - t0) / 30 * 1e3)
# cProfile one step
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(20): one()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
