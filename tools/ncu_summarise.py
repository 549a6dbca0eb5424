"""Join an ncu raw CSV (ncu -i X.ncu-rep --page raw --csv) with the key list written by tools/ncu_capture.py and
print per-(entry, shape) rows plus per-kernel per-step totals: duration, DRAM bytes, tensor-pipe and DRAM
utilisation.   python tools/ncu_summarise.py gpurun_out/r01_full_raw.csv gpurun_out/r01_full_keys.json"""
import collections
import csv
import io
import json
import sys


def num(v):
    try:
        return float(v.replace(",", ""))
    except (ValueError, AttributeError):
        return float("nan")


def main():
    raw, keys = sys.argv[1], json.load(open(sys.argv[2]))
    lines = [l for l in open(raw) if not l.startswith("==")]
    rd = list(csv.reader(io.StringIO("".join(lines))))
    hdr, units, rows = rd[0], rd[1], rd[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def get(r, name, scale_unit=None):
        i = col.get(name)
        if i is None:
            return float("nan")
        v = num(r[i])
        u = units[i]
        if scale_unit == "us":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        if scale_unit == "B":
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        return v

    # launches issued by one entry point may be several kernels (e.g. rans encode: records, encode, scan, pack)
    per_kernel = collections.OrderedDict()
    print(f"{'kernel':34s} {'grid':>8s} {'us':>10s} {'dramMB':>9s} {'dram%':>6s} {'tensor%':>7s} {'sm%':>6s} {'regs':>4s}  key")
    ki = 0
    pending = []
    for r in rows:
        name = r[col["Kernel Name"]].split("(")[0].replace("icm::", "").replace("void ", "")
        us = get(r, "gpu__time_duration.sum", "us")
        db = get(r, "dram__bytes_read.sum", "B") + get(r, "dram__bytes_write.sum", "B")
        dp = get(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
        tp = get(r, "sm__inst_executed_pipe_tensor.sum.pct_of_peak_sustained_active")
        if tp != tp:
            tp = get(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
        sp = get(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed")
        regs = get(r, "launch__registers_per_thread")
        grid = r[col["Grid Size"]] if "Grid Size" in col else ""
        pending.append((name, grid, us, db, dp, tp, sp, regs))
    # assign rows to keys in order: every key owns >= 1 consecutive rows; split by expected kernel names
    owner = {"icm_conv2d": ["conv_igemm_kernel"], "icm_swin_mlp": ["swin_mlp_kernel"], "icm_rans_encode_batch": ["rans_records_kernel", "rans_encode_kernel", "rans_scan_kernel", "rans_pack_kernel"]}
    pi = 0
    for k in keys:
        entry = k["key"][0]
        n = len(owner.get(entry, [None]))
        if entry == "icm_swin_block" and k["key"][4] == 96:
            n = 2  # C = 96 runs the attention half and the MLP half as two kernels
        for j in range(n):
            if pi >= len(pending):
                break
            name, grid, us, db, dp, tp, sp, regs = pending[pi]
            pi += 1
            print(f"{name[:34]:34s} {grid:>8s} {us:10.1f} {db / 1e6:9.2f} {dp:6.1f} {tp:7.1f} {sp:6.1f} {regs:4.0f}  {k['key']} x{k['calls_per_step']}")
            a = per_kernel.setdefault(name, [0, 0.0, 0.0, 0.0])
            a[0] += k["calls_per_step"]
            a[1] += us * k["calls_per_step"]
            a[2] += db * k["calls_per_step"]
            a[3] += k["flops"] * k["calls_per_step"] if j == 0 else 0.0
    print()
    print("per step (each captured launch weighted by how often its shape occurs in one compress+decompress):")
    for name, (n, us, db, fl) in per_kernel.items():
        extra = f"  {fl / us / 1e6:8.1f} TFLOP/s" if fl else ""
        print(f"{name[:40]:40s} launches {n:5d}  {us / 1e3:9.2f} ms  DRAM {db / 1e9:8.3f} GB  {db / us / 1e3:8.1f} GB/s  per-launch {db / n / 1e6:8.2f} MB{extra}")


if __name__ == "__main__":
    main()
