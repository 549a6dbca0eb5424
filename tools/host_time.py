"""Host enqueue time vs GPU time of one compress / decompress step (is the pipeline launch-bound?)."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "image-compression-for-machine_b200")); sys.path.insert(1, REPO)
import torch
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
model = bench.make_model(dev)
model.micro_batches = mb
x = bench.make_images(B, 0).to(dev)
for _ in range(3):
    c = model.compress(x, device_strings=True); d = model.decompress(c["strings"], c["shape"])
torch.cuda.synchronize()
for it in range(2):
    t0 = time.perf_counter(); c = model.compress(x, device_strings=True); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    d = model.decompress(c["strings"], c["shape"]); t3 = time.perf_counter(); torch.cuda.synchronize(); t4 = time.perf_counter()
    print(f"B={B} mb={mb}: compress host {1e3*(t1-t0):.1f} ms, total {1e3*(t2-t0):.1f} ms | decompress host {1e3*(t3-t2):.1f} ms, total {1e3*(t4-t2):.1f} ms")
