"""Limits of the reference itself, reproduced: (1) do its zig-zag / LRCP variants (compressai/models/stf5.py,
stf6.py) run at all?  (2) does its STF accept image sizes that are not multiples of 64?

TEST INFRASTRUCTURE, build container only (reads /root/reference through oracle/refshim.py).  Imports the two UNMODIFIED
files next to the reference's stf.py, builds each model and calls forward() and compress() on one 64x64 image.  Result
recorded in DESIGN.md section 7: forward() runs, compress() raises in both (stf5: a ModuleList is called; stf6: an attribute the
constructor never creates), i.e. the variants have no working bit-stream path to reproduce.

    python oracle/check_variants.py
"""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import refshim  # noqa: E402


def main():
    h = refshim.install("port")
    pkg = os.path.join(h["root"], "compressai", "models")
    for f in ("stf5.py", "stf6.py"):
        os.symlink(os.path.join(refshim.REF, "compressai", "models", f), os.path.join(pkg, f))
    for name in ("stf5", "stf6"):
        mod = importlib.import_module("compressai.models." + name)
        cls = [getattr(mod, n) for n in dir(mod) if n.startswith("SymmetricalTransFormer")][0]
        torch.manual_seed(0)
        m = cls().eval()
        m.update(force=True)
        x = torch.rand(1, 3, 64, 64)
        for what in ("forward", "compress"):
            try:
                with torch.no_grad():
                    out = m(x) if what == "forward" else m.compress(x)
                print(f"{name}.{cls.__name__}.{what}: runs ({sorted(out)})")
            except Exception as e:  # noqa: BLE001
                print(f"{name}.{cls.__name__}.{what}: FAILS with {type(e).__name__}: {str(e)[:120]}")
    # Does the reference STF accept sizes that are not multiples of 64 (VERDICT r1 "in-model padding")?  Its Swin blocks pad
    # internally (stf.py:158-163), but the context model concatenates the 4x-upsampled hyper-prior output with the y slices
    # (stf.py:613): the sizes only agree when H / 16 is a multiple of 4.
    stf = importlib.import_module("compressai.models.stf")
    torch.manual_seed(0)
    m = stf.SymmetricalTransFormer().eval()
    m.update(force=True)
    for hw in ((64, 64), (96, 96), (100, 150), (64, 96), (128, 72)):
        for what in ("forward", "compress"):
            try:
                with torch.no_grad():
                    _ = m(torch.rand(1, 3, *hw)) if what == "forward" else m.compress(torch.rand(1, 3, *hw))
                print(f"stf {hw} {what}: runs")
            except Exception as e:  # noqa: BLE001
                print(f"stf {hw} {what}: FAILS with {type(e).__name__}: {str(e)[:90]}")
    refshim.uninstall(h)


if __name__ == "__main__":
    main()
