"""Per-job timeline of the stream pipeline: when each job's C phase, encoders, decode loop and S phase begin/end.

    python tools/pipe_trace.py [steps] [bench flags ...]   e.g.  python tools/pipe_trace.py 10 --chains 2 --lag 6
"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.join(REPO, "image-compression-for-machine_b200"))
sys.path.insert(1, REPO)
import argparse  # noqa: E402

import torch  # noqa: E402

import bench  # noqa: E402
from compressai.utils.pipeline import RoundTripPipeline  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("steps", type=int, nargs="?", default=10)
    ap.add_argument("--streams", type=int, default=12)
    ap.add_argument("--part", type=int, default=16)
    ap.add_argument("--dec-per-cta", type=int, default=4)
    ap.add_argument("--lag", type=int, default=6)
    ap.add_argument("--chains", type=int, default=2)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--repeat", type=int, default=2)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    model = bench.make_model(dev)
    x = bench.make_images(a.batch, 0).to(dev)
    pipe = RoundTripPipeline(model, n_streams=a.streams, part=a.part, decoder_streams_per_cta=a.dec_per_cta, lag=a.lag, chains=a.chains)
    marks = []

    def mark(tag):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream())
        marks.append((tag, ev))

    orig_c, orig_d, orig_s = model._compress_part, model._decode_part, model._synthesis
    state = {"job": 0, "sjob": 0}

    def c_part(xd, phase=None):
        j = state["job"]
        state["job"] += 1
        mark((j, "start"))

        def ph(what):
            if phase:
                if what == "begin":
                    phase(what)
            mark((j, "C_" + what))
            if phase and what == "end":
                phase(what)

        r = orig_c(xd, phase=ph)
        mark((j, "enc_end"))
        return r

    def d_part(*args, **kw):
        j = state["job"] - 1
        r = orig_d(*args, **kw)
        mark((j, "dec_end"))
        return r

    def s_part(*args, **kw):
        j = state["sjob"]
        state["sjob"] += 1
        mark((j, "S_begin"))
        r = orig_s(*args, **kw)
        mark((j, "S_end"))
        return r

    pipe.roundtrip([x] * 4)
    torch.cuda.synchronize()
    model._compress_part, model._decode_part, model._synthesis = c_part, d_part, s_part
    for rep in range(a.repeat):
        marks.clear()
        state["job"] = state["sjob"] = 0
        t0 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0.record()
        pipe.roundtrip([x] * a.steps)
        torch.cuda.synchronize()
        rows = {}
        for (j, tag), ev in marks:
            rows.setdefault(j, {})[tag] = t0.elapsed_time(ev)
        end = max(r["S_end"] for r in rows.values())
        print(f"--- repeat {rep}: {a.steps} steps in {end:.1f} ms = {a.steps * a.batch / end * 1e3:.1f} img/s")
        print(" job   start  C_begin   C_end  enc_end  dec_end  S_begin   S_end |  C    enc   dec  wait   S   total")
        for j in sorted(rows):
            r = rows[j]
            print(f"{j:4d} {r['start']:7.1f} {r['C_begin']:8.1f} {r['C_end']:7.1f} {r['enc_end']:8.1f} {r['dec_end']:8.1f} {r['S_begin']:8.1f} {r['S_end']:7.1f} |"
                  f" {r['C_end'] - r['C_begin']:5.1f} {r['enc_end'] - r['C_end']:5.1f} {r['dec_end'] - r['enc_end']:5.1f} {r['S_begin'] - r['dec_end']:5.1f} {r['S_end'] - r['S_begin']:5.1f} {r['S_end'] - r['start']:6.1f}")


if __name__ == "__main__":
    main()
