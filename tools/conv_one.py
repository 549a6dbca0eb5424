"""Run one icm_conv2d shape a few times (for ncu source-level captures).
    python tools/conv_one.py M Cin Cout act out_dtype residual [B H W k]   e.g.  conv_one.py 6291456 48 192 1 0 0"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "image-compression-for-machine_b200"))
import torch  # noqa: E402

from compressai.models._engine import Engine, PackedConv  # noqa: E402

M, Cin, Cout, act, odt, res = (int(v) for v in sys.argv[1:7])
k = int(sys.argv[10]) if len(sys.argv) > 10 else 1
B, H, W = (int(v) for v in sys.argv[7:10]) if len(sys.argv) > 9 else (1, 1, M)
eng = Engine(None)
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(B * H * W, Cin, device="cuda", generator=g).bfloat16()
pk = PackedConv(torch.randn(Cout, Cin, k, k, device="cuda", generator=g) / (Cin * k * k) ** 0.5, torch.randn(Cout, device="cuda", generator=g), 1, k // 2, 0)
r = torch.randn(B * H * W, Cout, device="cuda", generator=g) if res else None
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
out = eng.conv(x, B, H, W, pk, act=act, out_dtype=odt, residual=r)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
N_IT = int(os.environ.get("CONV_ONE_ITERS", "10"))
ms = 0.0
for it in range(3 + N_IT):
    flush.zero_()  # evict L2 between launches
    e0.record()
    eng.conv(x, B, H, W, pk, out=out, act=act, out_dtype=odt, residual=r)
    e1.record()
    torch.cuda.synchronize()
    if it >= 3:
        ms += e0.elapsed_time(e1) / N_IT
by = x.numel() * 2 + out.numel() * out.element_size() * (2 if res else 1)
print(f"{ms:.3f} ms  {2.0 * B * H * W * Cin * Cout * k * k / ms / 1e9:.1f} TFLOP/s  {by / ms / 1e6:.0f} GB/s")
