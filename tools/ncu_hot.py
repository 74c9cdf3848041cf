"""Summarise an .ncu-rep: headline metrics + the hottest SASS instructions by stall samples.
usage: python tools/ncu_hot.py file.ncu-rep [min_pct]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.8
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, v = rows[0], rows[2]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "sm__cycles_elapsed.max",
        "lts__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]
for k in keys:
    for i, x in enumerate(h):
        if x == k:
            print(f"{k:60s} {rows[1][i]:14s} {v[i]}")
st = []
for i, x in enumerate(h):
    if "issue_stalled" in x and x.endswith("per_issue_active.ratio"):
        try:
            st.append((float(v[i]), x.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        except ValueError:
            pass
print("stall cycles per issued instruction:", ", ".join(f"{b} {a:.2f}" for a, b in sorted(st, reverse=True)[:7]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hd = rows[1]
ix = {k: i for i, k in enumerate(hd)}
data = rows[2:]
tot = sum(int(r[ix["# Samples"]]) for r in data)
print("total samples", tot, "sass instructions", len(data))
for n, r in enumerate(data):
    s = int(r[ix["# Samples"]])
    if s > tot * minpct / 100:
        print(n, r[ix["Source"]][:64].ljust(64), f"{100*s/tot:5.1f}%", "exec", r[ix["Instructions Executed"]], "thr/inst",
              r[ix["Avg. Threads Executed"]], "lsb", r[ix["stall_long_sb"]], "ssb", r[ix["stall_short_sb"]], "wait", r[ix["stall_wait"]],
              "br", r[ix["stall_branch_resolving"]])
