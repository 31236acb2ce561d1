"""Per-CUDA-source-line summary of an ncu source page exported with
    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > X.src.csv
Aggregates warp-stall samples and executed instructions per (file, line) and prints the heaviest lines with their
dominant stall reasons.  Usage: python scripts/ncu_lines.py X.src.csv [top_n] [kernel-substring]"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    want = sys.argv[3] if len(sys.argv) > 3 else ""
    rows = list(csv.reader(open(path, newline="")))
    per = defaultdict(lambda: defaultdict(float))
    text = {}
    f = fn = None
    hdr = None
    active = True
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            f = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            fn = r[1]
            active = want in fn
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or not active or len(r) < len(hdr) - 2:
            continue
        if r[0] == "":        # SASS row nested under a CUDA line: already aggregated into the line row
            continue
        try:
            line = int(r[0])
        except ValueError:
            continue
        key = (f, line)
        text[key] = r[1].strip()[:110]
        for i, h in enumerate(hdr):
            if h in ("# Samples", "Instructions Executed", "Thread Instructions Executed") or \
                    (h.startswith("stall_") and "Not Issued" not in h):
                try:
                    per[key][h] += float(r[i] or 0)
                except ValueError:
                    pass
    tot = sum(v["# Samples"] for v in per.values()) or 1.0
    toti = sum(v["Instructions Executed"] for v in per.values()) or 1.0
    print(f"total samples {tot:.0f}, warp instructions {toti:.0f}")
    agg = defaultdict(float)
    for v in per.values():
        for h, x in v.items():
            if h.startswith("stall_"):
                agg[h] += x
    print("stall mix: " + ", ".join(f"{h[6:]} {100 * x / tot:.1f}%" for h, x in sorted(agg.items(), key=lambda kv: -kv[1])[:9]))
    for key, v in sorted(per.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
        st = sorted(((h[6:], x) for h, x in v.items() if h.startswith("stall_") and x > 0), key=lambda kv: -kv[1])[:3]
        print(f"{100 * v['# Samples'] / tot:6.2f}% smp {100 * v['Instructions Executed'] / toti:6.2f}% inst  "
              f"{key[0]}:{key[1]:<5d} {' '.join(f'{a}={100 * b / max(1.0, v[chr(35) + chr(32) + chr(83) + chr(97) + chr(109) + chr(112) + chr(108) + chr(101) + chr(115)]):.0f}%' for a, b in st):32s} | {text[key]}")


main()
