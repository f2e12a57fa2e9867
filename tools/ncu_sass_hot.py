"""Per-opcode and hottest-instruction summary of one kernel from an .ncu-rep (source page, SASS view).
usage: python tools/ncu_sass_hot.py rep.ncu-rep [kernel-regex] [launch-skip]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
cmd = ["ncu", "-i", rep, "--page", "source", "--csv"]
if len(sys.argv) > 2:
    cmd += ["--kernel-name", "regex:" + sys.argv[2]]
if len(sys.argv) > 3:
    cmd += ["--launch-skip", sys.argv[3], "--launch-count", "1"]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
hdr = rows[0]
ix = {n: i for i, n in enumerate(hdr)}
ops = collections.Counter()
samples = collections.Counter()
tot = 0
tot_s = 0
recs = []
for r in rows[1:]:
    if len(r) < len(hdr) or not r[ix["Instructions Executed"]].isdigit():
        continue
    n = int(r[ix["Instructions Executed"]])
    s = int(r[ix["# Samples"]] or 0)
    op = r[ix["Source"]].split()[0] if not r[ix["Source"]].startswith("@") else r[ix["Source"]].split()[1]
    op = op.split(".")[0]
    ops[op] += n
    samples[op] += s
    tot += n
    tot_s += s
    recs.append((s, n, r[ix["Source"]][:90]))
print("total warp instructions %d, samples %d" % (tot, tot_s))
print("by opcode (executed, %% of total | stall samples %%):")
for op, n in ops.most_common(25):
    print("  %-12s %12d %5.1f%% | %5.1f%%" % (op, n, 100.0 * n / tot, 100.0 * samples[op] / max(tot_s, 1)))
print("hottest instructions by stall samples:")
for s, n, src in sorted(recs, reverse=True)[:25]:
    print("  %6d samples %10d exec  %s" % (s, n, src))
