"""Per-shape time of every icm_conv2d call in one compress+decompress step (CUDA events per call)."""
import collections
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "image-compression-for-machine_b200"))
sys.path.insert(1, REPO)
import torch  # noqa: E402

import bench  # noqa: E402
from compressai import _native  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    dev = torch.device("cuda", 0)
    model = bench.make_model(dev)
    model.micro_batches = 1
    x = bench.make_images(B, 0).to(dev)
    for _ in range(2):
        c = model.compress(x, device_strings=True)
        model.decompress(c["strings"], c["shape"])
    keys = {"icm_conv2d": [], "icm_conv2d_grouped": []}
    orig = _native.Profile._work

    def work(name, args):
        if name in ("icm_conv2d", "icm_conv2d_grouped"):
            a = args[0]._obj
            G = args[1]._obj.groups if name == "icm_conv2d_grouped" else 1
            keys[name].append((a.B, a.H, a.W, a.Cin, a.Cout, a.KH, a.stride, a.act, a.out_dtype, a.pixel_shuffle, bool(a.residual), G))
        return orig(name, args)

    _native.Profile._work = staticmethod(work)
    with _native.Profile() as prof:
        c = model.compress(x, device_strings=True)
        model.decompress(c["strings"], c["shape"])
        torch.cuda.synchronize()
        recs = [r for name in keys for r in prof.records.get(name, [])]
        agg = collections.OrderedDict()
        for k, (s, e, w) in zip([k for name in keys for k in keys[name]], recs):
            t = s.elapsed_time(e)
            a = agg.setdefault(k, [0, 0.0, 0.0])
            a[0] += 1
            a[1] += t
            a[2] += w
    tot = sum(v[1] for v in agg.values())
    print(f"B={B}: {len(recs)} conv calls, {tot:.2f} ms, {sum(v[2] for v in agg.values()) / tot / 1e9:.1f} TFLOP/s average")
    print("  B    H    W  Cin Cout k s act odt ps res  G | calls      ms   share  TFLOP/s   GB/s(min traffic)")
    for k, (n, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        Bq, H, W, Cin, Cout, K, s, act, odt, ps, res, G = k
        Ho, Wo = (H + 2 * (K // 2) - K) // s + 1, (W + 2 * (K // 2) - K) // s + 1
        bytes_min = n * G * (Bq * H * W * Cin * 2 + Bq * Ho * Wo * Cout * (4 if odt else 2) * (2 if res else 1) + Cout * Cin * K * K * 2)
        print(f"{Bq:3d} {H:4d} {W:6d} {Cin:4d} {Cout:4d} {K} {s} {act:3d} {odt:3d} {ps:2d} {int(res):3d} {G:2d} | {n:5d} {ms:7.3f} {ms / tot:7.3f} {fl / ms / 1e9:8.1f} {bytes_min / ms / 1e6:8.1f}")


if __name__ == "__main__":
    main()
