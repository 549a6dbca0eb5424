"""Where does the host thread spend the e2e step?  cProfile of K pipelined host-io steps (top cumulative entries)."""
import cProfile, io, os, pstats, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "image-compression-for-machine_b200")); sys.path.insert(1, REPO)
import torch
import bench
from compressai.utils.pipeline import RoundTripPipeline
K = int(sys.argv[1]) if len(sys.argv) > 1 else 6
dev = torch.device("cuda", 0)
model = bench.make_model(dev)
x = bench.make_images(64, 0).pin_memory()
outs = [torch.empty_like(x).pin_memory() for _ in range(2)]
pipe = RoundTripPipeline(model, n_streams=12, part=32, decoder_streams_per_cta=8, lag=8, chains=2, decode_priority=True)
pipe.roundtrip([x] * 6, host_io=True, out_host=[outs[i % 2] for i in range(6)])
torch.cuda.synchronize()
t0 = time.perf_counter()
pipe.roundtrip([x] * K, host_io=True, out_host=[outs[i % 2] for i in range(K)])
torch.cuda.synchronize()
print(f"e2e wall {1e3 * (time.perf_counter() - t0) / K:.1f} ms/step")
pr = cProfile.Profile()
pr.enable()
pipe.roundtrip([x] * K, host_io=True, out_host=[outs[i % 2] for i in range(K)])
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
print(s.getvalue()[:5000])
