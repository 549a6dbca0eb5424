"""WACNN (cnn) 768x512 throughput: plain API per-family times and the stream pipeline.  python tools/wacnn_bench.py [B] [steps]"""
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "image-compression-for-machine_b200"))
sys.path.insert(1, REPO)
import torch  # noqa: E402

import bench  # noqa: E402
from compressai import _native  # noqa: E402
from compressai.utils.pipeline import RoundTripPipeline  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda", 0)
model = bench.make_model(dev, arch="cnn")
x = bench.make_images(B, 0).to(dev)
model.micro_batches = 1
for _ in range(2):
    c = model.compress(x, device_strings=True)
    model.decompress(c["strings"], c["shape"])
torch.cuda.synchronize()
with _native.Profile() as prof:
    c = model.compress(x, device_strings=True)
    model.decompress(c["strings"], c["shape"])
    summ = prof.summary()
tot = sum(v[1] for v in summ.values())
print(f"WACNN B={B} 3x768x512 plain API: {tot:.1f} ms of kernels per step")
for k, v in sorted(summ.items(), key=lambda kv: -kv[1][1]):
    extra = f"  {v[2] / v[1] / 1e9:7.1f} TFLOP/s" if v[2] else ""
    print(f"  {k:40s} calls {v[0]:4d}  {v[1]:8.3f} ms{extra}")
pipe = RoundTripPipeline(model)
pipe.roundtrip([x] * 6, keep_outputs=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
pipe.roundtrip([x] * steps, keep_outputs=False)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"pipeline: {steps} steps of {B} images in {ms:.1f} ms = {steps * B / ms * 1e3:.1f} images/s")
