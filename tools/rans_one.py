"""One encode + a few decode steps of S streams (ncu target): python tools/rans_one.py S kind streams_per_cta kernel [steps]"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "image-compression-for-machine_b200")); sys.path.insert(1, REPO)
import numpy as np, torch
from compressai import ans
from compressai._native import check, lib
from compressai.entropy_models import GaussianConditional
from compressai.models.stf import get_scale_table

S, kind, spc, lanes = int(sys.argv[1]), sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
nsteps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
N = 49152 * nsteps
gc = GaussianConditional(None); gc.update_scale_table(get_scale_table()); gc = gc.cuda()
T = gc.device_tables()
rng = np.random.default_rng(0)
table = gc.scale_table.cpu().numpy()
if kind == "idx0":
    idx = np.zeros((S, N), np.int32); sym = np.rint(rng.normal(0, 0.8, (S, N))).astype(np.int32)
elif kind == "stress":
    idx = rng.integers(0, 47, (S, N)).astype(np.int32); sym = np.rint(rng.normal(0, 2.2 * table[idx] + 0.6)).astype(np.int32)
else:
    idx = rng.integers(0, 64, (S, N)).astype(np.int32); sym = np.rint(rng.normal(0, table[idx])).astype(np.int32)
ds, di = torch.from_numpy(sym).cuda(), torch.from_numpy(idx).cuda()
check(lib().icm_set_decoder_layout(spc, lanes), "layout")
for it in range(2):
    packed, sizes = ans.encode_streams(T, ds, di, return_device="async")
    dec = ans.acquire_decoder(S); dec.set_streams_device(packed, sizes)
    outs = [dec.decode_step(T, di[:, k * 49152:(k + 1) * 49152].contiguous()) for k in range(nsteps)]
    torch.cuda.synchronize()
    assert torch.equal(torch.cat(outs, 1), ds)
    ans.release_decoder(dec)
print("ok")
