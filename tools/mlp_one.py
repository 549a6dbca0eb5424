"""Time icm_swin_mlp against the two-launch path for one shape:  python tools/mlp_one.py M C"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "image-compression-for-machine_b200"))
import torch  # noqa: E402

from compressai.models._engine import Engine  # noqa: E402

M, C = int(sys.argv[1]), int(sys.argv[2])
eng = Engine(None)
mlp = torch.nn.Module()
mlp.fc1 = torch.nn.Linear(C, 4 * C)
mlp.fc2 = torch.nn.Linear(4 * C, C)
mlp = mlp.cuda()
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(M, C, device="cuda", generator=g)
xn = torch.randn(M, C, device="cuda", generator=g).bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for fused in (True, False):
    eng.fused_mlp = fused
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = 0.0
    for it in range(13):
        flush.zero_()
        e0.record()
        eng.mlp(xn, x, mlp)
        e1.record()
        torch.cuda.synchronize()
        if it >= 3:
            ms += e0.elapsed_time(e1) / 10
    by = M * C * (2 + 4 + 4)
    print(f"M={M} C={C} fused={fused}: {ms:.3f} ms  {16.0 * M * C * C / ms / 1e9:.1f} TFLOP/s  {by / ms / 1e6:.0f} GB/s (fused minimum traffic)")
