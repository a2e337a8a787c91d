"""Write profiles/chain_traffic.json from ncu reports: DRAM bytes per launch of the dominant kernel at a named launch
shape (read by bench.py for `roofline.traffic`).  Runs in the build container (ncu reads .ncu-rep files without a GPU).

   python tools/update_traffic.py <precision> <workload> <report.ncu-rep> "<command the report was captured from>"
"""
import csv, json, os, subprocess, sys, io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
prec, workload, rep, how = sys.argv[1:5]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, vals = rows[0], rows[2]
m = dict(zip(hdr, vals))
unit = dict(zip(hdr, rows[1]))


def mbytes(name):
    v, u = float(m[name]), unit[name]
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


rec = {"dram_bytes_per_launch": mbytes("dram__bytes_read.sum") + mbytes("dram__bytes_write.sum"),
       "dram_bytes_read": mbytes("dram__bytes_read.sum"), "dram_bytes_write": mbytes("dram__bytes_write.sum"),
       "kernel": m["Kernel Name"], "grid": m["launch__grid_size"], "duration_us_under_ncu": float(m["gpu__time_duration.sum"]),
       "tensor_pipe_pct_of_elapsed": float(m["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]),
       "report": os.path.relpath(rep, ROOT), "captured_with": how}
f = os.path.join(ROOT, "profiles", "chain_traffic.json")
d = json.load(open(f)) if os.path.exists(f) else {}
d.setdefault(prec, {})[workload] = rec
json.dump(d, open(f, "w"), indent=1, sort_keys=True)
print(json.dumps(rec, indent=1))
