"""BASELINE.json configs[4] -- the CISA microbenchmark sweep: query 38x50x1024 against `units = way x shot` support maps
of Ns x 1024 (Ns = 196: BASELINE's 14x14; Ns = 400: the reference's native 20x20), way, shot in 1..10.

  python tools/cisa_sweep.py --out gpurun_out/cisa_sweep.json                 # timing pass (CUDA-graph replay, L2 flushed)
  ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum \
      --clock-control none -k regex:conv_gemm --csv --log-file gpurun_out/cisa_sweep_ncu.csv \
      python tools/cisa_sweep.py --ncu-pass                                   # one eager block per point, 4 GEMMs each
  python tools/cisa_sweep.py --merge gpurun_out/cisa_sweep.json gpurun_out/cisa_sweep_ncu.csv   # adds tensor-pipe %

Per point: ms, algorithmic TFLOP/s and GB/s (SURVEY.md section 8d formulas) against the measured peaks, and -- after
the merge -- the tensor-pipe utilisation of the four tensor-core launches of the block (k-projection, q-projection,
logits + softmax, P.V), each and time-weighted."""
import argparse
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
C, D, NQ = 1024, 256, 1900


def points(ns_list):
    units = sorted({w * s for w in range(1, 11) for s in range(1, 11)})
    return [(ns, u) for ns in ns_list for u in units]


def make_runner(units, hs, prec):
    import torch
    import dana_b200  # noqa: F401
    from dana_b200 import ops
    from dana_b200.engine import DanaEngine
    from dana_b200.ops import Pair
    from dana_b200.synthetic import synthetic_state_dict
    global _SD
    try:
        _SD
    except NameError:
        _SD = synthetic_state_dict(1996)
    eng = DanaEngine(_SD, n_shot=units, precision=prec)
    split = prec != "bf16"
    g = torch.Generator(device="cuda").manual_seed(units * 1000 + hs)
    base = torch.relu(torch.randn(1, C, 38, 50, device="cuda", generator=g))
    sup = torch.relu(torch.randn(units, C, hs, hs, device="cuda", generator=g))
    bp = ops.split_f32(base.permute(0, 2, 3, 1).contiguous(), split).view(NQ, C)
    dense = Pair.empty((NQ, C), "cuda", split)
    supp = ops.split_f32(sup.permute(0, 2, 3, 1).contiguous(), split)
    return lambda: eng.rpn_attention(bp, dense, 1, NQ, supp, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--ns", default="196,400")
    ap.add_argument("--precision", default="bf16x3")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--ncu-pass", action="store_true")
    ap.add_argument("--merge", nargs=2)
    a = ap.parse_args()
    if a.merge:
        return merge(*a.merge)
    import torch
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) \
        else {"bf16_tflops": 1590.0, "hbm_gbs": 6650.0}
    ns_list = [int(v) for v in a.ns.split(",")]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    rows = []
    for ns, units in points(ns_list):
        hs = int(round(ns ** 0.5))
        run = make_runner(units, hs, a.precision)
        if a.ncu_pass:
            flush.zero_()
            run()
            torch.cuda.synchronize()
            continue
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            run()
            with torch.cuda.graph(graph, stream=side):
                run()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph.replay()
        tot = 0.0
        for _ in range(a.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        ms = tot / a.iters
        flops = 2.0 * NQ * C * D + units * (2.0 * ns * C * D + 2.0 * NQ * ns * D + 2.0 * NQ * ns * C + 8.0 * ns * C)
        byts = 2.0 * C * (NQ + units * ns) + 4.0 * NQ * C + 1.05e6
        rows.append({"ns": ns, "units": units, "ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1),
                     "pct_peak_tf": round(100 * flops / ms / 1e9 / peaks["bf16_tflops"], 2),
                     "gbs": round(byts / ms / 1e6), "pct_peak_hbm": round(100 * byts / ms / 1e6 / peaks["hbm_gbs"], 2)})
        print(json.dumps(rows[-1]), flush=True)
        del graph
    if a.out and not a.ncu_pass:
        with open(a.out, "w") as f:
            json.dump({"workload": "BA+CISA block, Nq 1900 x C 1024, d 256, units = way x shot (way, shot in 1..10)",
                       "precision": a.precision, "timing": "CUDA-graph replay, L2 flushed, mean of %d" % a.iters,
                       "peaks": peaks, "points": rows}, f, indent=1)


def merge(json_path, csv_path):
    data = json.load(open(json_path))
    lines = [l for l in open(csv_path) if not l.startswith("==")]
    per_launch = {}
    for row in csv.DictReader(lines):
        i = int(row["ID"])
        v = float(row["Metric Value"].replace(",", ""))
        d = per_launch.setdefault(i, {})
        if row["Metric Name"].startswith("sm__pipe_tensor"):
            d["pipe"] = v
        elif row["Metric Name"] == "gpu__time_duration.sum":
            d["us"] = v / 1e3 if row["Metric Unit"] in ("ns", "nsecond") else v
    launches = [per_launch[i] for i in sorted(per_launch)]
    names = ("k_proj", "q_proj", "logits_softmax", "pv")
    assert len(launches) == 4 * len(data["points"]), (len(launches), len(data["points"]))
    for pi, pt in enumerate(data["points"]):
        g = launches[4 * pi:4 * pi + 4]
        pt["tensor_pipe_pct"] = {n: round(x["pipe"], 1) for n, x in zip(names, g)}
        pt["gemm_us_under_ncu"] = {n: round(x["us"], 1) for n, x in zip(names, g)}
        tot = sum(x["us"] for x in g)
        pt["tensor_pipe_pct_time_weighted"] = round(sum(x["us"] * x["pipe"] for x in g) / tot, 1)
    data["tensor_pipe_metric"] = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active per tensor-core launch " \
                                 "(ncu, one eager block per point, L2 flushed); time-weighted over the four launches"
    with open(json_path, "w") as f:
        json.dump(data, f, indent=1)
    print("merged %d points" % len(data["points"]))


if __name__ == "__main__":
    main()
