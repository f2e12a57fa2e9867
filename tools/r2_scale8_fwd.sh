#!/bin/bash
O=gpurun_out/r2s
mkdir -p $O
: > $O/fwd_scale_final.jsonl
timeout 200 python bench.py --gpus 1 --no-cpu-baseline 2>/dev/null | tail -1 >> $O/fwd_scale_final.jsonl
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --no-cpu-baseline 2>/dev/null | tail -1 >> $O/fwd_scale_final.jsonl
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 tools/train_bench.py --gpus 8 --steps 10 2>/dev/null | tail -1 > $O/train_scale8_final.jsonl
cut -c1-130 $O/fwd_scale_final.jsonl $O/train_scale8_final.jsonl
