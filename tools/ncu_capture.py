"""Drive one compress+decompress step so that ncu captures ONE launch of every distinct (entry point, shape).

    ncu --set full --clock-control none --profile-from-start off -o gpurun_out/r01_full python tools/ncu_capture.py 64

The step runs twice untimed first; in the third pass cudaProfilerStart/Stop bracket the first call of each
distinct key, and the keys (+ how often each occurs in the step) are written to gpurun_out/<tag>_keys.json in
capture order so the per-kernel ncu rows can be weighted back into per-step totals (tools/ncu_summarise.py)."""
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


def key_of(name, args):
    if name in ("icm_conv2d", "icm_conv2d_grouped"):
        a = args[0]._obj
        G = args[1]._obj.groups if name == "icm_conv2d_grouped" else 1
        return (name, a.B, a.H, a.W, a.Cin, a.Cout, a.KH, a.stride, a.act, a.out_dtype, a.pixel_shuffle, int(bool(a.residual)), a.res_mode, G)
    ints = tuple(int(v) for v in args if isinstance(v, int) and 0 <= v < (1 << 24))
    return (name,) + ints


class Capture(_native.Profile):
    def __init__(self, enabled):
        super().__init__()
        self.enabled = enabled
        self.seen = collections.OrderedDict()
        self.flops = {}
        self.rt = torch.cuda.cudart()

    def __getattr__(self, name):
        fn = getattr(self._L, name)
        if name in self._HOST or not name.startswith("icm_"):
            return fn

        def w(*args):
            k = key_of(name, args)
            first = k not in self.seen
            self.seen[k] = self.seen.get(k, 0) + 1
            if first:
                self.flops[k] = self._work(name, args)
                if self.enabled:
                    self.rt.cudaProfilerStart()
            rc = fn(*args)
            if first and self.enabled:
                self.rt.cudaProfilerStop()
            return rc

        return w


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    arch = sys.argv[2] if len(sys.argv) > 2 else "stf"
    tag = sys.argv[3] if len(sys.argv) > 3 else "r01_full"
    dev = torch.device("cuda", 0)
    model = bench.make_model(dev, arch=arch)
    model.micro_batches = 1
    x = bench.make_images(B, 0).to(dev)
    for _ in range(2):
        c = model.compress(x, device_strings=True)
        model.decompress(c["strings"], c["shape"])
    torch.cuda.synchronize()
    with Capture(True) as cap:
        c = model.compress(x, device_strings=True)
        model.decompress(c["strings"], c["shape"])
        torch.cuda.synchronize()
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    with open(os.path.join(REPO, "gpurun_out", tag + "_keys.json"), "w") as f:
        json.dump([{"key": list(k), "calls_per_step": n, "flops": cap.flops[k]} for k, n in cap.seen.items()], f)
    print(len(cap.seen), "distinct (entry, shape) keys;", sum(cap.seen.values()), "calls per step")


if __name__ == "__main__":
    main()
