"""Print an ncu launch-list CSV (gpu__time_duration.sum) as: id, kernel, block, grid, ms."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        print(d["ID"], d["Kernel Name"][:58].ljust(58), d["Block Size"], d["Grid Size"], round(float(d["Metric Value"]) / 1e6, 3), "ms")
