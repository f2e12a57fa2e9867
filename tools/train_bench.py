"""Training-step benchmark (BASELINE.json configs[3]: res101, 800x1333 queries, 5 shots per support set, one episode
per GPU, fwd + bwd + NCCL gradient all-reduce + SGD) -- the one workload of the north star with a collective.

  python tools/train_bench.py                                   # one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
      tools/train_bench.py --gpus 8                             # weak scaling: one episode per rank

Prints one JSON line: whole-job train-images/s (sum of per-rank batches / max-over-ranks device time), ms per step and
the forward / backward / all-reduce wait / SGD split of one instrumented step.  The reference trains 2 support sets
(positive + negative, dana.py:100-108) whatever `way` says, so "5-way 5-shot" is 2 x 5 support crops per image here.
Weights: synthetic (reference init distributions), data: synthetic episodes with 1-8 random boxes per image."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--layers", type=int, default=101)
    ap.add_argument("--batch", type=int, default=1, help="images per GPU")
    ap.add_argument("--height", type=int, default=800)
    ap.add_argument("--width", type=int, default=1333)
    ap.add_argument("--shots", type=int, default=5)
    ap.add_argument("--bucket-mb", type=int, default=25)
    ap.add_argument("--eager", action="store_true", help="launch every kernel from Python instead of replaying the two CUDA graphs")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import dana_b200  # noqa: F401
    from dana_b200 import ops
    from dana_b200.config import cfg_from_file, cfg_from_list, reset_cfg
    from dana_b200.dana import DAnARCNN
    from dana_b200.sharding import aggregate_throughput
    from dana_b200.synthetic import synthetic_episode, synthetic_state_dict
    from dana_b200.train_step import SGDTrainer
    reset_cfg()
    cfg_from_file(os.path.join(ROOT, "cfgs", "res50.yml"))
    cfg_from_list(["ANCHOR_SCALES", "[4,8,16,32]", "ANCHOR_RATIOS", "[0.5,1,2]", "MAX_NUM_GT_BOXES", "50"])
    torch.manual_seed(1996)
    net = DAnARCNN(["bg", "fg"], "concat", 256, 256, pretrained=False, semantic_enhance=True, num_layers=args.layers,
                   num_way=2, num_shot=args.shots, precision="bf16x3")
    net.create_architecture()
    net.load_state_dict(synthetic_state_dict(1996, num_layers=args.layers), strict=False)
    net.cuda().train()
    trainer = SGDTrainer(net, bucket_bytes=args.bucket_mb << 20, cuda_graph=not args.eager)
    n_params = sum(g[1].numel() for g in trainer.groups)

    b = args.batch
    im, info, sup = synthetic_episode(100 + rank, b, args.height, args.width, 2 * args.shots, pin=True)
    rs = np.random.RandomState(7 + rank)
    gt = torch.zeros(b, 50, 5)
    nb = torch.zeros(b, dtype=torch.long)
    for i in range(b):
        n = rs.randint(1, 9)
        w, h = rs.uniform(64, 400, n), rs.uniform(64, 400, n)
        x1, y1 = rs.uniform(0, args.width - w - 1), rs.uniform(0, args.height - h - 1)
        gt[i, :n] = torch.from_numpy(np.stack([x1, y1, x1 + w, y1 + h, np.ones(n)], 1).astype(np.float32))
        nb[i] = n
    np.random.seed(1996 + rank)          # the target layers sample with numpy's global RNG
    d_im, d_info, d_gt, d_nb, d_sup = [t.to(dev) for t in (im, info, gt, nb, sup)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    losses = []
    for _ in range(args.warmup):
        loss, _ = trainer.step(d_im, d_info, d_gt, d_nb, d_sup)
        losses.append(float(loss))
    # device-resident timing
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ops.LAUNCHES = 0
    e0.record()
    for _ in range(args.steps):
        loss, _ = trainer.step(d_im, d_info, d_gt, d_nb, d_sup)
    e1.record()
    barrier()
    launches = ops.LAUNCHES
    secs = e0.elapsed_time(e1) / 1e3
    losses.append(float(loss))
    value, max_s, units = aggregate_throughput(b * args.steps, secs, dev)
    # end to end: inputs from pinned host memory every step, loss read back
    barrier()
    e0.record()
    for _ in range(args.steps):
        t = [x.to(dev, non_blocking=True) for x in (im, info, gt, nb, sup)]
        loss, _ = trainer.step(*t)
        float(loss)
    e1.record()
    barrier()
    e2e, _, _ = aggregate_throughput(b * args.steps, e0.elapsed_time(e1) / 1e3, dev)
    # one instrumented step
    trainer.events = []
    trainer.step(d_im, d_info, d_gt, d_nb, d_sup)
    torch.cuda.synchronize()
    ev = trainer.events
    split = {ev[i + 1][0] + "_ms": round(ev[i][1].elapsed_time(ev[i + 1][1]), 3) for i in range(len(ev) - 1)}
    trainer.events = None
    err = ops.device_error()
    if rank == 0:
        print(json.dumps({
            "metric": "train-images/sec", "value": round(value, 3), "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(max_s / args.steps * 1e3, 2),
            "higher_is_better": True, "scaling": "weak", "dtype": "bf16x3 (split-bf16 operands, fp32 accumulate), fp32 master weights",
            "data": "synthetic", "e2e": {"value": round(e2e, 3), "unit": "images/s",
                                         "h2d_bytes_per_step": int(sum(x.numel() * x.element_size() for x in (im, info, gt, nb, sup))),
                                         "d2h_bytes_per_step": 4},
            "config": {"workload": "res%d DAnA training step (fwd + bwd + grad all-reduce + SGD), %d image/GPU, %dx%d, 2 sets x %d shots"
                       % (args.layers, b, args.height, args.width, args.shots), "trainable_params": n_params,
                       "allreduce_bytes_per_step": 4 * n_params if world > 1 else 0, "bucket_mb": args.bucket_mb,
                       "launch": "eager (Python launches)" if args.eager else "two CUDA graphs around the host target layers",
                       "collective": ("NCCL all-reduce (sum / world) of the gradient arenas, " + ("bucketed, launched from post-accumulate hooks during backward" if args.eager else "after the backward graph")) if world > 1 else "none"},
            "split_one_step": split, "gpu_launches": launches, "loss_first_last": [round(losses[0], 4), round(losses[-1], 4)],
            "device_error": err}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
