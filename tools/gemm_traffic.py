"""Writes profiles/rNN_gemm_traffic_<precision>.json from an ncu launch list of one forward step
(`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`, tools/profile_step.py):
DRAM bytes per launch of dana::conv_gemm_kernel, stamped with the commit the capture was taken at.  bench.py reports
the newest such file of its precision as roofline.traffic.

  python tools/gemm_traffic.py gpurun_out/launches_mixed.csv mixed profiles/r02_gemm_traffic_mixed.json"""
import csv
import json
import subprocess
import sys

path, precision, out = sys.argv[1], sys.argv[2], sys.argv[3]
lines = [l for l in open(path) if not l.startswith("==")]
by = {}
for row in csv.DictReader(lines):
    if "conv_gemm_kernel" not in row["Kernel Name"]:
        continue
    d = by.setdefault(int(row["ID"]), {"rd": 0.0, "wr": 0.0, "ns": 0.0})
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    if row["Metric Name"] == "dram__bytes_read.sum":
        d["rd"] += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    elif row["Metric Name"] == "dram__bytes_write.sum":
        d["wr"] += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    elif row["Metric Name"] == "gpu__time_duration.sum":
        d["ns"] += v * {"ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(unit, 1)
n = len(by)
tot = sum(d["rd"] + d["wr"] for d in by.values())
try:
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
except OSError:
    commit = None
json.dump({"kernel": "dana::conv_gemm_kernel", "precision": precision, "launches_per_step": n,
           "dram_bytes_per_step": tot, "dram_bytes_per_launch": tot / max(n, 1),
           "gemm_time_ms_under_ncu": sum(d["ns"] for d in by.values()) / 1e6, "commit": commit,
           "source": "%s (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                     "--clock-control none, one forward step of tools/profile_step.py --precision %s)" % (path, precision)},
          open(out, "w"), indent=1)
print(open(out).read())
