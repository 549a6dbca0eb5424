"""Per-kernel counts of the SASS mnemonics that prove which hardware path a kernel uses (B200_PROFILING.md): tcgen05 MMA
(UTCHMMA), TMEM loads (LDTM), TMA (UTMALDG / UTMASTG), tensor-memory barriers (UTCBAR), mma.sync (HMMA), ldmatrix (LDSM).

    python tools/sass_summary.py > profiles/r02_sass_summary.txt        (needs only cuobjdump; no GPU)"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "image-compression-for-machine_b200", "lib", "libicm_b200.so")
PAT = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTMAPF", "SYNCS", "HMMA", "LDSM", "MOVM", "MUFU", "IMAD.WIDE", "LDG", "STG", "LDS", "STS"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.splitlines()
    counts, order, cur, k = {}, [], None, 0
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = names[k].split("(")[0].replace("void ", "").replace("icm::", "")
            k += 1
            if cur not in counts:
                counts[cur] = collections.Counter()
                order.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for p in PAT:
                if op.startswith(p):
                    counts[cur][p] += 1
    stamp = os.path.join(os.path.dirname(LIB), "build.stamp")
    print(f"# cuobjdump -sass {os.path.relpath(LIB, REPO)} (build stamp {open(stamp).read()[:12] if os.path.exists(stamp) else '?'}), instructions per kernel by mnemonic prefix")
    cols = [p for p in PAT if any(counts[n][p] for n in order)]
    print(f"{'kernel':46s} {'total':>6s} " + " ".join(f"{c:>9s}" for c in cols))
    for n in order:
        print(f"{n[:46]:46s} {counts[n]['_total']:6d} " + " ".join(f"{counts[n][c]:9d}" for c in cols))


if __name__ == "__main__":
    sys.exit(main())
