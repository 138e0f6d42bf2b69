import sys, time, torch, gc
sys.path.insert(0, '.')
import bench, b200ssl
dev = torch.device('cuda:0')
inp = bench.make_inputs(dev, 0)
step = b200ssl.LossPathStep(num_classes=2, mode="binary")
def one():
    return step(inp["image_a"], inp["image_b"], inp["teacher_a"], inp["teacher_b"], inp["scores"], inp["target"], inp["params"], inp["ema_params"])
use_nvml = len(sys.argv) > 1 and sys.argv[1] == "nvml"
cs = bench.ClockSampler(0)
if use_nvml: cs.start()
t0 = time.perf_counter(); n = 0
while time.perf_counter() - t0 < 1.2:
    one(); n += 1
torch.cuda.synchronize()
print("warmup steps", n, "nvml", use_nvml)
for rep in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    h0 = time.perf_counter()
    e0.record()
    hs = []
    for i in range(100):
        a = time.perf_counter(); one(); hs.append(time.perf_counter() - a)
    e1.record()
    h1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"region {rep}: gpu {e0.elapsed_time(e1)/100:.4f} ms/step, host loop {(h1-h0)*10:.4f} ms/step, max host step {max(hs)*1e3:.2f} ms at {hs.index(max(hs))}, gc {gc.get_count()}")
