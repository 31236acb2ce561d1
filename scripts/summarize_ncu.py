"""Summarise an ncu report (.ncu-rep) and a launch list (csv) into small text files under profiles/."""
import csv, io, subprocess, sys, collections

def raw_metrics(rep):
    # a .ncu-rep, or the CSV that `ncu -i rep --page raw --csv` printed on the GPU box (the reports with imported
    # source exceed the 64 MiB that travel back)
    if rep.endswith(".csv"):
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]

WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]

def main():
    rep, launches, out = sys.argv[1], sys.argv[2], sys.argv[3]      # launches: a csv, or "-" for none
    hdr, units, rows = raw_metrics(rep)
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on  ({rep})\n")
        for r in rows:
            for w in WANT:
                for i, h in enumerate(hdr):
                    if h == w:
                        f.write(f"{h} [{units[i]}] = {r[i]}\n")
            f.write("\n")
        if launches == "-":
            f.flush()
            print(open(out).read())
            return
        # launch list
        f.write(f"# launch list: ncu --metrics gpu__time_duration.sum --clock-control none  ({launches})\n")
        txt = open(launches).read()
        start = txt.index('"ID"')
        agg = collections.OrderedDict()
        for r in csv.DictReader(io.StringIO(txt[start:])):
            if r.get("Metric Name") != "gpu__time_duration.sum":
                continue
            name = r["Kernel Name"][:90]
            v = float(r["Metric Value"].replace(",", ""))
            u = r["Metric Unit"]
            ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += ms
        tot = sum(a[1] for a in agg.values())
        f.write(f"total {tot:.3f} ms over {sum(a[0] for a in agg.values())} launches\n")
        for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
            f.write(f"{a[1]:12.3f} ms  {100 * a[1] / tot:6.2f}%  x{a[0]:<4d} {name}\n")
    print(open(out).read())

main()
