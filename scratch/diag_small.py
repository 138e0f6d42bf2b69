import sys, torch
sys.path.insert(0, '.')
import bench, b200ssl
dev = torch.device('cuda:0')
inp = bench.make_inputs(dev, 0)
step = b200ssl.LossPathStep(num_classes=2, mode="binary")
for _ in range(4):
    step(inp["image_a"], inp["image_b"], inp["teacher_a"], inp["teacher_b"], inp["scores"], inp["target"], inp["params"], inp["ema_params"])
torch.cuda.synchronize()
print("ok")
