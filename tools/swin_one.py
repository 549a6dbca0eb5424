"""One fused Swin block at benchmark scale (ncu / timing target): python tools/swin_one.py C [B] [parts]"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "image-compression-for-machine_b200")); sys.path.insert(1, REPO)
import torch
import bench
C = int(sys.argv[1]) if len(sys.argv) > 1 else 48
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda", 0)
model = bench.make_model(dev)
eng = model._engine
stage = {48: 0, 96: 1, 192: 2, 384: 3}[C]
blk = model.layers[stage].blocks[1]
H, W = 384 >> stage, 256 >> stage
x = torch.randn(B * H * W, C, device=dev)
for fused in (True, False):
    eng.fused_block = fused
    for _ in range(3):
        eng.swin_block(x, B, H, W, blk, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eng.swin_block(x, B, H, W, blk, True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gb = x.numel() * 4 * (2 if C == 48 else 4) / 1e9
    print(f"C={C} B={B} fused={fused}: {ms:.3f} ms per block ({gb / ms * 1e3:.0f} GB/s of the fused kernel's minimal traffic)")
eng.fused_block = True
