"""Timeline (CUDA events, ms since step start) of the entry points of one decompress, per CUDA stream."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "image-compression-for-machine_b200")); sys.path.insert(1, REPO)
import torch
import bench
from compressai import _native
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
what = sys.argv[2] if len(sys.argv) > 2 else "decompress"
dev = torch.device("cuda", 0)
model = bench.make_model(dev)
x = bench.make_images(B, 0).to(dev)
for _ in range(2):
    c = model.compress(x, device_strings=True); d = model.decompress(c["strings"], c["shape"])
torch.cuda.synchronize()
log = []
orig = _native.Profile._work
def work(name, args):
    log.append((name, torch.cuda.current_stream().cuda_stream))
    return orig(name, args)
_native.Profile._work = staticmethod(work)
base = torch.cuda.Event(enable_timing=True)
with _native.Profile() as prof:
    base.record()
    if what == "compress":
        c = model.compress(x, device_strings=True)
    else:
        d = model.decompress(c["strings"], c["shape"])
    torch.cuda.synchronize()
    order = []
    counters = {}
    for name, st in log:
        k = counters.get(name, 0); counters[name] = k + 1
        s, e, _ = prof.records[name][k]
        order.append((base.elapsed_time(s), base.elapsed_time(e), st, name))
streams = sorted({o[2] for o in order})
print("streams", streams)
for t0, t1, st, name in order:
    if name in ("icm_rans_decoder_step", "icm_rans_encode_batch") or (t1 - t0) > 0.8:
        print(f"{streams.index(st)} {t0:8.2f} -> {t1:8.2f} ({t1 - t0:6.2f})  {name}")
print("end", max(o[1] for o in order))
