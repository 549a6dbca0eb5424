"""GPU time of one training step by kernel (torch profiler), eager, fp32 and bf16 autocast: where does the step go?"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "image-compression-for-machine_b200")); sys.path.insert(1, REPO)
import torch
from torch.profiler import profile, ProfilerActivity
from compressai.training import Trainer
from compressai.zoo import models
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
for label, ac in (("fp32", None), ("bf16", torch.bfloat16)):
    torch.manual_seed(0)
    net = models["stf"]().to(dev).train()
    tr = Trainer(net, autocast=ac)
    x = torch.rand(B, 3, 256, 256, device=dev)
    for _ in range(3):
        tr.step(x)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        tr.step(x)
        torch.cuda.synchronize()
    rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
    tot = sum(e.device_time_total for e in rows)
    print(f"== {label}: {tot / 1e3:.1f} ms of GPU kernel time in {sum(e.count for e in rows)} launches")
    for e in rows[:22]:
        print(f"{e.device_time_total / 1e3:8.2f} ms {e.count:6d}  {e.key[:110]}")
    tr.buckets.remove()
    del tr, net
