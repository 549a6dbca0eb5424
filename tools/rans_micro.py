"""Coder-only micro-benchmark: S streams x N symbols.

    python tools/rans_micro.py S kind [layouts...]
kind: idx0 (default-init-like statistics: table 0, |sym| <= 3), stress (tables 0..46, ~26 % escapes), lowrate, uniform
layouts: "streams_per_cta:kernel" for the decoder (kernel 1 = round-1 warp-search kernel).  Default sweep if none given.
"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "image-compression-for-machine_b200")); sys.path.insert(1, REPO)
import numpy as np, torch
from compressai import ans
from compressai._native import check, lib
from compressai.entropy_models import GaussianConditional
from compressai.models.stf import get_scale_table

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
kind = sys.argv[2] if len(sys.argv) > 2 else "idx0"
layouts = sys.argv[3:] or ["0:1", "1:0", "2:0", "4:0", "8:0", "16:0"]
N = 49152 * 12
gc = GaussianConditional(None); gc.update_scale_table(get_scale_table()); gc = gc.cuda()
T = gc.device_tables()
rng = np.random.default_rng(0)
table = gc.scale_table.cpu().numpy()
if kind == "idx0":
    idx = np.zeros((S, N), np.int32); sym = np.rint(rng.normal(0, 0.8, (S, N))).astype(np.int32)
elif kind == "stress":
    idx = rng.integers(0, 47, (S, N)).astype(np.int32); sym = np.rint(rng.normal(0, 2.2 * table[idx] + 0.6)).astype(np.int32)
elif kind == "model":  # symbols drawn from the tables themselves, narrow tables, (almost) no escapes: the pure common path
    idx = rng.integers(0, 24, (S, N)).astype(np.int32); sym = np.clip(np.rint(rng.normal(0, table[idx])), -((gc._cdf_length.cpu().numpy()[idx] - 3) // 2), (gc._cdf_length.cpu().numpy()[idx] - 3) // 2).astype(np.int32)
elif kind == "lowrate":
    idx = np.minimum(rng.geometric(0.15, (S, N)) - 1, 63).astype(np.int32); sym = np.rint(rng.normal(0, table[idx])).astype(np.int32)
else:
    idx = rng.integers(0, 64, (S, N)).astype(np.int32); sym = np.rint(rng.normal(0, table[idx])).astype(np.int32)
ds, di = torch.from_numpy(sym).cuda(), torch.from_numpy(idx).cuda()
steps = [di[:, k * 49152:(k + 1) * 49152].contiguous() for k in range(12)]
esc = float(np.mean(np.abs(sym - 0) > ((gc._cdf_length.cpu().numpy()[idx] - 3) // 2)))
print(f"{kind}: S={S} N={N} escapes {100 * esc:.1f} %")


def run(label):
    best_e, best_d, ok, nb = 1e9, 1e9, True, 0
    for it in range(3):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        packed, sizes = ans.encode_streams(T, ds, di, return_device="async")
        e1.record()
        dec = ans.acquire_decoder(S); dec.set_streams_device(packed, sizes)
        outs = [dec.decode_step(T, st) for st in steps]
        e2.record(); torch.cuda.synchronize()
        ok &= torch.equal(torch.cat(outs, 1), ds)
        nb = int(sizes[:S].sum())
        best_e, best_d = min(best_e, e0.elapsed_time(e1)), min(best_d, e1.elapsed_time(e2))
        ans.release_decoder(dec)
    print(f"  {label:>14}: encode {best_e:7.2f} ms ({N / best_e / 1e3:5.1f} Msym/s/stream)  decode {best_d:7.2f} ms "
          f"({N / best_d / 1e3:5.1f} Msym/s/stream)  {8 * nb / S / N:.2f} bit/sym  roundtrip {ok}", flush=True)


for lay in layouts:
    a, b = (int(v) for v in lay.split(":"))
    check(lib().icm_set_decoder_layout(a, b), "icm_set_decoder_layout")
    run(lay)
