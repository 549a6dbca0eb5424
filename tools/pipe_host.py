"""Is the stream pipeline host-bound?  Wall time of the enqueue loop vs the GPU time of K pipelined steps.

    python tools/pipe_host.py [K] [part] [lag]
"""
import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "image-compression-for-machine_b200")); sys.path.insert(1, REPO)
import torch
import bench
from compressai import _native
from compressai.utils.pipeline import RoundTripPipeline
K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
part = int(sys.argv[2]) if len(sys.argv) > 2 else 32
lag = int(sys.argv[3]) if len(sys.argv) > 3 else 8
dev = torch.device("cuda", 0)
model = bench.make_model(dev)
x = bench.make_images(64, 0).to(dev)
pipe = RoundTripPipeline(model, n_streams=12, part=part, decoder_streams_per_cta=8, lag=lag, chains=2)
pipe.roundtrip([x] * 6, keep_outputs=False)
torch.cuda.synchronize()
for it in range(3):
    l0 = _native.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    pipe.roundtrip([x] * K, keep_outputs=False)
    t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize(); t2 = time.perf_counter()
    n = _native.launch_count() - l0
    print(f"K={K} part={part} lag={lag}: host enqueue {1e3*(t1-t0)/K:.1f} ms/step, wall {1e3*(t2-t0)/K:.1f} ms/step, GPU {e0.elapsed_time(e1)/K:.1f} ms/step, "
          f"{n/K:.0f} launches/step = {1e6*(t1-t0)/n:.1f} us of host time per launch", flush=True)
