"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list
by kernel name: launches, summed time, share of the step, and (when captured) DRAM bytes per launch."""
import collections
import csv
import re
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
lines = [l for l in open(path) if not l.startswith("==")]
tot, cnt = collections.defaultdict(float), collections.Counter()
dram = collections.defaultdict(float)
UNIT_T = {"ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1.0}
UNIT_B = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = re.sub(r"^void ", "", name)[:64]
    v = float(row["Metric Value"].replace(",", ""))
    if row.get("Metric Name") == "gpu__time_duration.sum":
        tot[name] += v * UNIT_T.get(row["Metric Unit"], 1.0)
        cnt[name] += 1
    elif row.get("Metric Name") in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        dram[name] += v * UNIT_B.get(row["Metric Unit"], 1.0)
total = sum(tot.values())
print("total %.3f ms over %d launches (per-launch times are cold-cache, serialised: compare shares)" % (total / 1e6, sum(cnt.values())))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:top]:
    extra = "  dram %8.2f MB/launch" % (dram[k] / cnt[k] / 1e6) if dram else ""
    print("  %-66s n=%4d %8.3f ms %5.1f%%%s" % (k, cnt[k], v / 1e6, 100 * v / total, extra))
