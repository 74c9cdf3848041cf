"""Cycle attribution along a kernel's hot loop from an .ncu-rep with source-level samples.
usage: python tools/ncu_attr.py file.ncu-rep steps_per_warp [min_exec]   (prints every 8th instruction + memory/collective ops)"""
import csv, io, subprocess, sys
rep, steps = sys.argv[1], float(sys.argv[2])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
cyc = float(rr[2][rr[0].index("sm__cycles_elapsed.max")]) / steps
rows = list(csv.reader(io.StringIO(src)))
hd = rows[1]
ix = {k: i for i, k in enumerate(hd)}
data = rows[2:]
tot = sum(int(r[ix["# Samples"]]) for r in data)
emax = max(int(r[ix["Instructions Executed"]]) for r in data)
acc = 0
k = 0
print("cycles per step", round(cyc))
for n, r in enumerate(data):
    e = int(r[ix["Instructions Executed"]])
    s = int(r[ix["# Samples"]])
    if e >= emax * 0.05:
        acc += s
        k += 1
        t = r[ix["Source"]]
        if k % 8 == 0 or any(x in t for x in ("SHFL", "LDS", "LDG", "VOTE", "STG", "STS", "ATOM", "RED", "BSYNC", "BRA")):
            print(n, t[:60].ljust(60), "e=%.2f" % (e / emax), "t=" + r[ix["Avg. Threads Executed"]], "cyc=%6.1f" % (s / tot * cyc), "cum=%6.0f" % (acc / tot * cyc))
