"""roofline.traffic of bench.py: DRAM bytes per launch of every config's kernel, from ONE ncu pass over the bench command

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/traffic.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-ess

Usage: python scripts/traffic_from_ncu.py gpurun_out/traffic.csv profiles/r02_traffic.json
The bench runs, per config and in CONFIG order, 1 warm-up + 2 timed device-resident launches and then the end-to-end arm
(two handles of half the chains each); the 2nd and 3rd launch of each config's kernel group are the timed device-resident
ones -- their average is the figure.  bench.py reads the JSON (it never runs under a profiler itself)."""
import csv, json, sys
from collections import OrderedDict

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
KERNEL_OF = [("c2", "walnutspy_kernel<DiagT, 128, 4, 128, 4"), ("c1", "package_kernel<StdNormalT"),
             ("c2_nuts", "walnutspy_kernel<DiagSmT"), ("c3", "walnutspy_kernel<FunnelT"),
             ("c4", "walnutspy_kernel<LogRegMma104T"), ("c5", "walnutspy_kernel<StockWatsonT"),
             ("c2_1m", "walnutspy_kernel<DiagT, 128, 4, 128, 4")]


def main(src, dst):
    rows = list(csv.reader(l for l in open(src, newline="") if l.startswith('"')))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    launches = OrderedDict()
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        d = launches.setdefault(int(r[ix["ID"]]), {"kernel": r[ix["Kernel Name"]].replace("wn::", ""), "grid": r[ix["Grid Size"]]})
        name, unit, val = r[ix["Metric Name"]], r[ix["Metric Unit"]], float(r[ix["Metric Value"]].replace(",", ""))
        if name.startswith("dram__bytes"):
            d[name] = val * UNIT[unit]
        elif name == "gpu__time_duration.sum":
            d["ms"] = val * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)   # unit "nsecond" etc. handled below
    seq = [v for v in launches.values() if "walnutspy_kernel" in v["kernel"] or "package_kernel" in v["kernel"]]
    groups = []                     # consecutive launches of one kernel
    for v in seq:
        if groups and groups[-1][0]["kernel"] == v["kernel"]:
            groups[-1].append(v)
        else:
            groups.append([v])
    out, gi = {}, 0
    for name, pat in KERNEL_OF:
        while gi < len(groups) and pat not in groups[gi][0]["kernel"]:
            gi += 1
        if gi == len(groups):
            break
        g = groups[gi]
        gi += 1
        # the headline kernel group of c2 also holds the min-ESS launch when that leg runs: the first three are the device leg
        timed = g[1:3] if len(g) >= 3 else g
        rd = sum(x.get("dram__bytes_read.sum", 0.0) for x in timed) / len(timed)
        wr = sum(x.get("dram__bytes_write.sum", 0.0) for x in timed) / len(timed)
        out[name] = {"kernel": g[0]["kernel"], "grid": g[0]["grid"], "launches_in_group": len(g),
                     "dram_read_bytes_per_launch": rd, "dram_write_bytes_per_launch": wr,
                     "traffic_bytes_per_launch": rd + wr}
    json.dump({"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over "
                         "`python bench.py --steps 2 --warmup 1 --no-cpu --no-ess` (scripts/traffic_from_ncu.py); "
                         "average of the two timed device-resident launches of each config",
               "workloads": out}, open(dst, "w"), indent=1)
    for k, v in out.items():
        print(f"{k:8s} {v['traffic_bytes_per_launch'] / 1e9:9.3f} GB/launch  (read {v['dram_read_bytes_per_launch'] / 1e9:.3f}, "
              f"write {v['dram_write_bytes_per_launch'] / 1e9:.3f})  x{v['launches_in_group']}  {v['kernel'][:70]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
