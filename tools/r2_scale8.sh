#!/bin/bash
# 8-GPU box: training-step scaling 1/2/4/8 (weak: one episode per GPU) and the forward bench at 8
O=gpurun_out/r2s
mkdir -p $O
: > $O/train_scale.jsonl
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 300 python tools/train_bench.py --gpus 1 --steps 10 2>$O/train_$n.err | tail -1 >> $O/train_scale.jsonl
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29540+n)) tools/train_bench.py --gpus $n --steps 10 2>$O/train_$n.err | tail -1 >> $O/train_scale.jsonl
  fi
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 tools/train_bench.py --gpus 8 --steps 10 --eager 2>$O/train_8_eager.err | tail -1 > $O/train_scale_eager8.jsonl
cut -c1-130 $O/train_scale.jsonl $O/train_scale_eager8.jsonl
: > $O/fwd_scale.jsonl
timeout 300 python bench.py --gpus 1 --no-cpu-baseline 2>/dev/null | tail -1 >> $O/fwd_scale.jsonl
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --no-cpu-baseline 2>/dev/null | tail -1 >> $O/fwd_scale.jsonl
cut -c1-130 $O/fwd_scale.jsonl
