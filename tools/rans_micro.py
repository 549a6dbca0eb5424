"""Coder-only micro-benchmark: S streams x N symbols, default-init-like statistics (idx 0, |sym| <= 3) or mixed tables."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "image-compression-for-machine_b200")); sys.path.insert(1, REPO)
import numpy as np, torch
from compressai import ans
from compressai.entropy_models import GaussianConditional
from compressai.models.stf import get_scale_table
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
kind = sys.argv[2] if len(sys.argv) > 2 else "idx0"
N = 49152 * 12
gc = GaussianConditional(None); gc.update_scale_table(get_scale_table()); gc = gc.cuda()
T = gc.device_tables()
rng = np.random.default_rng(0)
table = gc.scale_table.cpu().numpy()
if kind == "idx0":
    idx = np.zeros((S, N), np.int32); sym = np.rint(rng.normal(0, 0.8, (S, N))).astype(np.int32)
elif kind == "lowrate":
    idx = np.minimum(rng.geometric(0.15, (S, N)) - 1, 63).astype(np.int32); sym = np.rint(rng.normal(0, table[idx])).astype(np.int32)
else:
    idx = rng.integers(0, 64, (S, N)).astype(np.int32); sym = np.rint(rng.normal(0, table[idx])).astype(np.int32)
ds, di = torch.from_numpy(sym).cuda(), torch.from_numpy(idx).cuda()
for it in range(3):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    packed, sizes = ans.encode_streams(T, ds, di, return_device="async")
    e1.record()
    dec = ans.acquire_decoder(S); dec.set_streams_device(packed, sizes)
    outs = [dec.decode_step(T, di[:, k * 49152:(k + 1) * 49152].contiguous()) for k in range(12)]
    e2.record(); torch.cuda.synchronize()
    ok = torch.equal(torch.cat(outs, 1), ds)
    nb = int(sizes[:S].sum())
    print(f"{kind} S={S}: encode {e0.elapsed_time(e1):.2f} ms ({N/e0.elapsed_time(e1)/1e3:.1f} Msym/s/stream), decode {e1.elapsed_time(e2):.2f} ms ({N/e1.elapsed_time(e2)/1e3:.1f} Msym/s/stream), {8*nb/S/N:.2f} bit/sym, roundtrip {ok}")
    ans.release_decoder(dec)
