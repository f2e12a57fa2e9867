"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
lines = [l for l in open(path) if not l.startswith("==")]
tot, cnt = collections.defaultdict(float), collections.Counter()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(row["Metric Unit"], 1.0)
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = re.sub(r"^void ", "", name)[:64]
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
print("total %.3f ms over %d launches (per-launch times are cold-cache, serialised: compare shares)" % (total / 1e6, sum(cnt.values())))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:top]:
    print("  %-66s n=%4d %8.3f ms %5.1f%%" % (k, cnt[k], v / 1e6, 100 * v / total))
