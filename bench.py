"""bench.py -- query-images/sec of the DAnA forward hot path (BASELINE.json configs[1]:
res50 DAnA 2-way 3-shot, bs=4 per GPU, 600x1000 synthetic queries, full eval forward).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16x3|bf16]

One step = one eval forward over a batch of 4 episodes (4 queries + 4 x 6 support crops) per GPU.
N > 1: launched by torchrun, one rank per GPU; episodes are sharded by rank with NO collective on the
data path (weak scaling); torch.distributed only agrees on max step time / summed units.
Prints ONE JSON line (see the task contract): value = device-timed throughput with inputs resident
in HBM; e2e = same metric through DAnARCNN.forward with pinned-host inputs copied H2D and results
read back D2H inside the timed region; roofline = the tensor-core implicit-GEMM kernel (dominant
kernel of the step); cpu_baseline = the oracle port of the reference timed on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH_PER_GPU = 4
HEIGHT, WIDTH = 600, 1000
WAYS, SHOTS = 2, 3
WORKLOAD = "res50 DAnA 2-way 3-shot, bs=4/GPU, 600x1000 synthetic queries, full eval forward"
# algorithmic-minimum FLOPs of one query: trunk 75.4 + 6 x 12.75 + CISA 9.2 + RPN 45.4
# + layer4 143.4 + 2 head passes x 19.1
GF_PER_QUERY = 388.0


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


def gemm_traffic(precision):
    """DRAM bytes per launch of the GEMM kernel (dram__bytes_read.sum + dram__bytes_write.sum averaged over the GEMM
    launches of one step).  ncu cannot run inside a timed bench, so this is the figure of the newest committed ncu
    launch list of the same step and precision (profiles/rNN_gemm_traffic_<precision>.json, written by
    tools/gemm_traffic.py from the capture, stamped with its commit); None when there is none for this precision.
    The run's own algorithmic bytes are printed beside it (roofline.algorithmic_bytes_per_launch)."""
    import glob
    cands = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_gemm_traffic_%s.json" % precision)))
    if not cands:
        return None, None
    with open(cands[-1]) as f:
        d = json.load(f)
    return round(d["dram_bytes_per_launch"]), {"file": os.path.relpath(cands[-1], ROOT), "commit": d.get("commit"),
                                               "launches_per_step": d.get("launches_per_step")}


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i] == "Active" for s in self.samples)]
        out = {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(self.samples[0][1]),
               "reasons": reasons, "samples": len(self.samples)}
        try:
            pw = [float(s[6]) for s in self.samples if len(s) > 6]
            if pw:
                out["power_w_median"], out["power_w_max"] = round(statistics.median(pw), 1), round(max(pw), 1)
        except ValueError:
            pass
        return out


def pin_host_cores():
    """One rank per GPU shares the host with its siblings: give every rank its own contiguous slice of the cores the
    process may use (launch-thread migration across sockets showed up as +1.3 % ms/step at 8 ranks in round 1) and a
    matching OpenMP width (torchrun otherwise sets OMP_NUM_THREADS=1 with a warning).  Returns the slice."""
    world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    try:
        cores = sorted(os.sched_getaffinity(0))
    except AttributeError:
        return None
    per = max(1, len(cores) // max(world, 1))
    mine = cores[local * per:(local + 1) * per] if world > 1 else cores
    if world > 1 and mine:
        try:
            os.sched_setaffinity(0, mine)
        except OSError:
            pass
    os.environ["OMP_NUM_THREADS"] = str(max(1, len(mine)))
    return [mine[0], mine[-1]] if mine else None


def dist_setup(n_gpus):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


# ------------------------------------------------------------------------------------------ ours
def run_ours(args):
    import torch

    import dana_b200  # noqa: F401
    from dana_b200 import ops
    from dana_b200.config import cfg, cfg_from_file, cfg_from_list
    from dana_b200.dana import DAnARCNN
    from dana_b200.sharding import aggregate_throughput
    from dana_b200.synthetic import synthetic_episode, synthetic_state_dict

    rank, world, local = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    cfg_from_file(os.path.join(ROOT, "cfgs", "res50.yml"))
    cfg_from_list(["ANCHOR_SCALES", "[4,8,16,32]", "ANCHOR_RATIOS", "[0.5,1,2]", "MAX_NUM_GT_BOXES", "50"])
    net = DAnARCNN(["bg", "fg"], "concat", 256, 256, pretrained=False, semantic_enhance=True, num_way=WAYS,
                   num_shot=SHOTS, precision=args.precision, use_cuda_graph=not args.no_graph)
    net.create_architecture()
    net.load_state_dict(synthetic_state_dict(1996), strict=False)
    net.to(dev).eval()
    eng = net.engine()

    b = BATCH_PER_GPU
    im_h, info_h, sup_h = synthetic_episode(1000 + rank, b, HEIGHT, WIDTH, WAYS * SHOTS, pin=True)
    gt_h, nb_h = torch.zeros(b, 1, 5).pin_memory(), torch.zeros(b).pin_memory()
    im_d, info_d, sup_d = im_h.to(dev), info_h.to(dev), sup_h.to(dev)
    gt_d, nb_d = gt_h.to(dev), nb_h.to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value)
    for _ in range(args.warmup):
        net(im_d, info_d, gt_d, nb_d, sup_d)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ops.LAUNCHES = 0
    evs = []
    for _ in range(args.steps):
        flush.zero_()                               # L2 flush between timed iterations (not timed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        net(im_d, info_d, gt_d, nb_d, sup_d)
        e1.record()
        evs.append((e0, e1))
    barrier()
    dev_s = sum(a.elapsed_time(c) for a, c in evs) / 1e3
    # kernels launched per step: counted on one eager forward (a graph replay launches the same kernel nodes)
    ops.LAUNCHES = 0
    eng.forward(im_d, info_d, sup_d)
    torch.cuda.synchronize()
    launches = ops.LAUNCHES * args.steps

    # ---- end-to-end timing through the public module API (e2e): every step copies its inputs from pinned host
    # memory (EpisodePrefetcher: the copy of step i+1 overlaps the forward of step i) and reads its results back
    # (ResultFetcher: into pinned buffers behind the step; the host waits for step i's results after it has enqueued
    # step i+1, so neither direction of the PCIe traffic leaves the GPU idle).  Every step's inputs cross H2D and every
    # step's results cross D2H inside the timed region.
    from dana_b200.pipeline import EpisodePrefetcher, ResultFetcher
    host_batch = (im_h, info_h, gt_h, nb_h, sup_h)

    def e2e_run(n):
        pf = EpisodePrefetcher(dev)
        fetch = ResultFetcher(dev)
        res_, pending = None, None
        for holders in pf.run(host_batch for _ in range(n)):
            rois, cls_prob, bbox_pred, *_ = net(*holders)
            ticket = fetch.start((rois, cls_prob, bbox_pred))        # async D2H of the step's results
            if pending is not None:
                res_ = fetch.wait(pending)                           # previous step's results are on the host
            pending = ticket
        if pending is not None:
            res_ = fetch.wait(pending)
        return res_

    res = e2e_run(max(2, args.warmup // 2))
    barrier()
    t0 = time.perf_counter()
    res = e2e_run(args.steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    sampler.stop_flag = True
    sampler.join(2)
    h2d = sum(t.numel() * t.element_size() for t in (im_h, info_h, gt_h, nb_h, sup_h))
    d2h = sum(t.numel() * t.element_size() for t in res)

    # ---- roofline of the dominant kernel + per-stage breakdown: one extra instrumented eager step, CUDA events around
    # every tensor-core GEMM launch and at every stage boundary, on the launching stream; algorithmic FLOPs =
    # 2*M*N*K and algorithmic bytes (every operand / output element once) from the launch shapes
    ops.GEMM_TRACE = []
    eng.stage_events = []
    eng.forward(im_d, info_d, sup_d)          # eager launches: events around every GEMM on the launching stream
    torch.cuda.synchronize()
    trace, ops.GEMM_TRACE = ops.GEMM_TRACE, None
    marks, eng.stage_events = eng.stage_events, None
    ops.STAGE = ""
    g_ms = sum(t[0].elapsed_time(t[1]) for t in trace)
    g_flops = sum(t[2] for t in trace)
    g_issued = sum(t[2] * t[4]["mma_per_product"] for t in trace)
    g_bytes = sum(t[4]["bytes"] for t in trace)
    peaks = load_peaks()
    stage_ms = {marks[i + 1][0]: marks[i][1].elapsed_time(marks[i + 1][1]) for i in range(len(marks) - 1)}
    eager_ms = sum(stage_ms.values())
    r_rois = b * 300
    hbm_stage_bytes = {"roi_align": r_rois * 49 * 1024 * (6 if args.precision == "mixed" else 4 if args.precision == "bf16x3" else 2)
                       + b * 38 * 63 * 1024 * 4}
    stages = []
    for name, ms in stage_ms.items():
        tr = [t for t in trace if t[4]["stage"] == name]
        fl = sum(t[2] for t in tr)
        ent = {"stage": name, "ms_eager": round(ms, 3), "share": round(ms / eager_ms, 4),
               "gemm_launches": len(tr), "gemm_ms": round(sum(t[0].elapsed_time(t[1]) for t in tr), 3),
               "algorithmic_gflop": round(fl / 1e9, 1)}
        if name in hbm_stage_bytes:
            ent.update(bound="hbm", algorithmic_mb=round(hbm_stage_bytes[name] / 1e6, 1),
                       achieved_gbs=round(hbm_stage_bytes[name] / (ms * 1e-3) / 1e9, 1),
                       frac=round(hbm_stage_bytes[name] / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], 4))
        elif fl > 0:
            ent.update(bound="tensor", achieved_tflops=round(fl / (ms * 1e-3) / 1e12, 1),
                       frac=round(fl / (ms * 1e-3) / 1e12 / peaks["tf_sustained"], 4),
                       issued_frac=round(sum(t[2] * t[4]["mma_per_product"] for t in tr) / (ms * 1e-3) / 1e12 / peaks["tf_sustained"], 4))
        stages.append(ent)
    traffic, traffic_src = gemm_traffic(args.precision)

    value, t_max, units = aggregate_throughput(units=b * args.steps, seconds=dev_s, device=dev)
    e2e_value, _, _ = aggregate_throughput(units=b * args.steps, seconds=e2e_s, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {
        "metric": "query-images/sec", "value": round(value, 2), "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * t_max / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"bf16x3": "bf16x3 (split-bf16 operands, fp32 accumulate; fp32-equivalent)", "bf16": "bf16",
                  "mixed": "bf16x3 + f16 (split-bf16 operands everywhere except layer4 and the RPN convs, which run on "
                           "single fp16 planes; fp32 accumulate; every stage within 1e-3 of the fp32 reference)"}[args.precision],
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": b, "query_hw": [HEIGHT, WIDTH], "ways": WAYS, "shots": SHOTS,
                   "support_hw": [320, 320], "rois_per_image": 300, "precision": args.precision,
                   "l2": "flushed (256 MiB memset) between timed iterations", "sharding": "episodes by rank, no collective",
                   "host_cores_rank0": getattr(args, "host_cores", None),
                   "launch": "eager" if args.no_graph else "cuda graph replay (one graph per input shape)"},
        "e2e": {"value": round(e2e_value, 2), "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "clocks": sampler.summary(),
        "roofline": {"bound": "tensor", "kernel": "dana::conv_gemm_kernel (tcgen05 implicit GEMM, all launches of a step)",
                     "achieved": round(g_flops / (g_ms * 1e-3) / 1e12, 1) if g_ms > 0 else None,
                     "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                     "frac": round(g_flops / (g_ms * 1e-3) / 1e12 / peaks["tf_sustained"], 4) if g_ms > 0 else None,
                     "issued_frac": round(g_issued / (g_ms * 1e-3) / 1e12 / peaks["tf_sustained"], 4) if g_ms > 0 else None,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes_per_launch": round(g_bytes / max(len(trace), 1)),
                     "algorithmic_bytes_per_step": round(g_bytes),
                     "peak_source": peaks["source"] + " sustained bf16 (kernel timed inside a step)",
                     "launches_per_step": len(trace),
                     "gemm_ms_per_step": round(g_ms, 3),
                     "gemm_ms_note": "sum of CUDA-event intervals around each GEMM launch of one EAGER step",
                     "algorithmic_gflop_per_step": round(g_flops / 1e9, 1),
                     "issued_gflop_per_step": round(g_issued / 1e9, 1),
                     "step_gflop_model": GF_PER_QUERY * b},
        "stages": {"timing": "one eager step, CUDA events at the stage boundaries (host launch gaps included: %.3f ms "
                             "eager vs ms_per_step under graph replay); shares are the comparable quantity" % eager_ms,
                   "list": stages},
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(sample_batch=1)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ CPU legs
def cpu_forward_fn(batch):
    """The oracle port of the reference forward (test infrastructure, timed here as the CPU baseline)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dana_oracle as O

    import dana_b200  # noqa: F401
    from dana_b200.synthetic import synthetic_episode, synthetic_state_dict
    torch.set_num_threads(os.cpu_count() or 1)
    p = synthetic_state_dict(1996)
    im, info, sup = synthetic_episode(1000, batch, HEIGHT, WIDTH, WAYS * SHOTS)
    ref_nms = ref_roi = None
    try:                                     # the reference's own compiled CPU operators, when they travelled
        import build_ref
        ref_c = build_ref.load()
        if ref_c is not None:
            ref_nms = lambda bx, sc, th: ref_c.nms(bx.contiguous(), sc.contiguous(), th)  # noqa: E731
            ref_roi = lambda f, r, s, ph, pw, sr: ref_c.roi_align_forward(f.contiguous(), r.contiguous(), s, ph, pw, sr)  # noqa: E731
    except Exception:  # noqa: BLE001
        pass

    def step():
        with torch.no_grad():
            return O.dana_forward_eval(p, im, info, sup, SHOTS, roi_align_fn=ref_roi, nms_fn=ref_nms)
    kind = "port (oracle/dana_oracle.py; NMS + RoIAlign from the reference's compiled CPU kernels)" if ref_nms else \
        "port (oracle/dana_oracle.py)"
    return step, kind


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.lower().startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_operator_baseline():
    """The reference's two custom CPU operators on ONE host thread (they are single-threaded, ROIAlign_cpu.cpp:133,
    nms_cpu.cpp:17-64) at the benchmark's per-image sizes: RoIAlign of 300 RoIs on a [1,1024,38,63] map and NMS of
    6000 boxes.  Uses the reference's own compiled operators when oracle/_ref travelled, else
    the oracle's C restatement."""
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dana_oracle as O
    roi_fn, nms_fn, kind = O.roi_align_forward, O.nms, "port (oracle/c/dana_oracle.c)"
    try:
        import build_ref
        ref_c = build_ref.load()
        if ref_c is not None:
            roi_fn = lambda f, r, s, ph, pw, sr: ref_c.roi_align_forward(torch.as_tensor(f), torch.as_tensor(r), s, ph, pw, sr)  # noqa: E731
            nms_fn = lambda bx, sc, th: ref_c.nms(torch.as_tensor(bx), torch.as_tensor(sc), th)  # noqa: E731
            kind = "reference (oracle/_ref: lib/model/csrc/cpu compiled in place)"
    except Exception:  # noqa: BLE001
        pass
    rs = np.random.RandomState(0)
    feat = rs.standard_normal((1, 1024, 38, 63)).astype(np.float32)
    cx, cy = rs.uniform(0, 1000, 300), rs.uniform(0, 600, 300)
    bw, bh = np.exp(rs.uniform(np.log(16), np.log(600), 300)), np.exp(rs.uniform(np.log(16), np.log(500), 300))
    rois = np.stack([np.zeros(300), np.clip(cx - bw / 2, 0, 999), np.clip(cy - bh / 2, 0, 599),
                     np.clip(cx + bw / 2, 0, 999), np.clip(cy + bh / 2, 0, 599)], 1).astype(np.float32)
    n = 6000
    x1, y1 = rs.uniform(0, 900, n), rs.uniform(0, 500, n)
    boxes = np.stack([x1, y1, x1 + rs.uniform(1, 200, n), y1 + rs.uniform(1, 200, n)], 1).astype(np.float32)
    scores = (rs.permutation(n) / n).astype(np.float32)
    prev = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        t0 = time.perf_counter()
        roi_fn(feat, rois, 1.0 / 16, 7, 7, 0)
        t_roi = time.perf_counter() - t0
        t0 = time.perf_counter()
        nms_fn(boxes, scores, 0.7)
        t_nms = time.perf_counter() - t0
    finally:
        torch.set_num_threads(prev)
    return {"threads": 1, "kind": kind, "roi_align_300_rois_ms": round(t_roi * 1e3, 1), "nms_6000_boxes_ms": round(t_nms * 1e3, 1)}


def cpu_baseline(sample_batch=1, reps=2):
    step, kind = cpu_forward_fn(sample_batch)
    step()
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    dt = (time.perf_counter() - t0) / reps
    out = {"value": round(sample_batch / dt, 4), "unit": "images/s", "cores": os.cpu_count(), "cpu": cpu_model(),
           "kind": "port", "detail": kind, "sample": "%d query 600x1000 + 6 support crops per step, %d timed steps (+1 warm-up), torch CPU fp32, "
                                     "%d threads" % (sample_batch, reps, os.cpu_count() or 1)}
    try:
        out["operators_single_thread"] = cpu_operator_baseline()
    except Exception as e:  # noqa: BLE001  -- the operator timings are an extra, never a reason to lose the line
        out["operators_single_thread"] = {"error": repr(e)[:200]}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_batch = 1
    step, kind = cpu_forward_fn(sample_batch)
    # one 600x1000 query + its 6 support crops per step (~0.7 s on 16 cores): --steps / --warmup are honoured as given
    # up to 60 steps / 10 warm-ups (a bound, so that a sustained-run request of the GPU arm does not run for hours here)
    warm = max(0, min(args.warmup, 10))
    steps = max(1, min(args.steps, 60))
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    v = round(sample_batch * steps / dt, 4)
    line = {"impl": "reference", "metric": "query-images/sec", "value": v, "unit": "images/s",
            "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": steps, "warmup": warm,
            "ms_per_step": round(1e3 * dt / steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": "1 query + 6 supports per step on the host cores",
                       "steps_requested": args.steps, "warmup_requested": args.warmup,
                       "process": "one CPU process on rank 0 whatever --gpus is (the reference path does not shard)"},
            "cpu_baseline": {"value": v, "unit": "images/s", "cores": os.cpu_count(), "cpu": cpu_model(), "kind": "port",
                             "detail": kind,
                             "sample": "%d steps x 1 query 600x1000 (2-way 3-shot), torch CPU fp32, %d threads" % (steps, os.cpu_count() or 1)},
            "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="mixed", choices=["mixed", "bf16x3", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--workload", default="forward", choices=["forward", "train"],
                    help="forward: the headline (BASELINE configs[1]); train: the training step of configs[3] "
                         "(res101, 800x1333, fwd + bwd + NCCL all-reduce + SGD) -- delegates to tools/train_bench.py")
    a = ap.parse_args()
    if a.workload == "train":
        import runpy
        sys.argv = [os.path.join(ROOT, "tools", "train_bench.py"), "--gpus", str(a.gpus), "--steps", str(a.steps),
                    "--warmup", str(max(a.warmup, 3))] + (["--eager"] if a.no_graph else [])
        runpy.run_path(sys.argv[0], run_name="__main__")
        sys.exit(0)
    if a.impl == "reference":
        run_reference(a)
    else:
        a.warmup = max(a.warmup, 3)
        a.host_cores = pin_host_cores()          # before torch is imported (OMP width is read at import)
        run_ours(a)
