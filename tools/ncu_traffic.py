"""Sum dram bytes per logical kernel over its size-class launches from `ncu --set full` reports and write
profiles/traffic.json (read by bench.py for roofline.traffic).
usage: python tools/ncu_traffic.py n_ids zipf_s source-note rep1.ncu-rep [rep2.ncu-rep ...]"""
import csv, io, json, subprocess, sys
from pathlib import Path

n_ids, zipf_s, note = int(float(sys.argv[1])), float(sys.argv[2]), sys.argv[3]
out = {}
for rep in sys.argv[4:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    def col(name):
        return h.index(name)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    for r in rows[2:]:
        name = r[col("Kernel Name")]
        key = "k_roc_encode" if "k_roc_encode" in name else "k_roc_decode" if "k_roc_decode" in name else name.split("(")[0]
        if "nan" in r[col("dram__bytes_read.sum")].lower():
            continue  # ncu sometimes fails to collect the dram counters of a launch: re-capture it alone and pass that report too
        rd = float(r[col("dram__bytes_read.sum")]) * scale[units[col("dram__bytes_read.sum")]]
        wr = float(r[col("dram__bytes_write.sum")]) * scale[units[col("dram__bytes_write.sum")]]
        ms = float(r[col("gpu__time_duration.sum")]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[units[col("gpu__time_duration.sum")]]
        e = out.setdefault(key, {"dram_bytes": 0.0, "dram_read": 0.0, "dram_write": 0.0, "launches": 0, "ms_serialised": 0.0})
        e["dram_bytes"] += rd + wr
        e["dram_read"] += rd
        e["dram_write"] += wr
        e["launches"] += 1
        e["ms_serialised"] += ms
Path("profiles").mkdir(exist_ok=True)
json.dump({"n_ids": n_ids, "zipf_s": zipf_s, "source": note, "kernels": out}, open("profiles/traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
