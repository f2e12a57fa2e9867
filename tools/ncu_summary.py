"""Extract the metrics the roofline discussion needs from .ncu-rep files into a markdown table.
usage: python tools/ncu_summary.py out.md rep1.ncu-rep [rep2.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time [us]"),
    ("dram__bytes_read.sum", "DRAM read [MB]"),
    ("dram__bytes_write.sum", "DRAM write [MB]"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "TMA load [GB]"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("smsp__inst_executed.sum", "warp instr"),
]


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    for r in rows[2:]:
        yield dict(zip(hdr, r))


def main():
    dst, reps = sys.argv[1], sys.argv[2:]
    lines = ["| report | kernel | " + " | ".join(t for _, t in METRICS) + " |", "|---|---|" + "---|" * len(METRICS)]
    for rep in reps:
        for r in rows_of(rep):
            name = r.get("Kernel Name", "?").split("(")[0][:60]
            vals = []
            for key, _ in METRICS:
                v = r.get(key, "")
                try:
                    v = "%.4g" % float(v.replace(",", ""))
                except ValueError:
                    pass
                vals.append(v)
            lines.append("| %s | `%s` | %s |" % (rep.split("/")[-1], name, " | ".join(vals)))
    with open(dst, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
