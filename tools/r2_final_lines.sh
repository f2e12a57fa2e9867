#!/bin/bash
# refresh of the bench lines after the last code change of the round (the full capture is tools/round_capture.sh)
O=gpurun_out/cap2
mkdir -p $O
timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench_mixed.log 2>&1
timeout 300 python bench.py --steps 1000 --warmup 20 --no-cpu-baseline > $O/bench_mixed_sustained.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --precision bf16x3 --no-cpu-baseline > $O/bench_x3.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --precision bf16 --no-cpu-baseline > $O/bench_bf16.log 2>&1
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.log 2>&1
timeout 300 python tools/train_bench.py --steps 20 > $O/train_bench.log 2>&1
timeout 300 python tools/train_bench.py --steps 10 --eager > $O/train_bench_eager.log 2>&1
timeout 300 python tools/train_bench.py --steps 10 --layers 50 --batch 4 --height 600 --width 1000 --shots 3 > $O/train_bench_res50_bs4.log 2>&1
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6) > $O/pytest_gpu.txt
for f in $O/bench_*.log $O/train_bench*.log; do tail -1 $f | cut -c1-150; done; tail -2 $O/pytest_gpu.txt
