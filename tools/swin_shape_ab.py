"""A/B of the fused Swin block's launch shape (csrc/swin_fused.cu launch_shaped): one process per ICM_SWIN_SHAPE value, since the
library reads the variable once.  Prints the time per block at benchmark scale and a digest of one application on seeded
input (every shape must give the same bits: the per-window arithmetic does not depend on how windows are dealt to warps).

    for s in 0 82 122 101 121; do ICM_SWIN_SHAPE=$s python tools/swin_shape_ab.py; done
"""
import hashlib, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "image-compression-for-machine_b200")); sys.path.insert(1, REPO)
import torch
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
model = bench.make_model(dev)
eng = model._engine
shape = os.environ.get("ICM_SWIN_SHAPE", "0")
for C in (48, 96):
    stage = {48: 0, 96: 1}[C]
    H, W = 384 >> stage, 256 >> stage
    for shifted in (False, True):
        blk = model.layers[stage].blocks[1 if shifted else 0]
        g = torch.Generator(device=dev); g.manual_seed(C)
        x = torch.randn(B * H * W, C, device=dev, generator=g)
        eng.swin_block(x, B, H, W, blk, shifted)
        torch.cuda.synchronize()
        digest = hashlib.sha1(x.cpu().numpy().tobytes()).hexdigest()[:12]
        x = torch.randn(B * H * W, C, device=dev, generator=g)
        for _ in range(3):
            eng.swin_block(x, B, H, W, blk, shifted)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            eng.swin_block(x, B, H, W, blk, shifted)
        e1.record(); torch.cuda.synchronize()
        print(f"shape={shape} C={C} shifted={int(shifted)} B={B}: {e0.elapsed_time(e1) / 20:.3f} ms per block  digest {digest}", flush=True)
