#!/bin/bash
O=gpurun_out/r2a
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > $O/pytest.txt
timeout 900 python tools/parity_report.py --out $O/parity.jsonl > $O/parity.log 2>&1
timeout 300 python bench.py --precision mixed --no-cpu-baseline > $O/bench_mixed.log 2>&1
timeout 300 python bench.py --precision bf16x3 --no-cpu-baseline > $O/bench_x3.log 2>&1
DANA_CLUSTER=3 timeout 300 python bench.py --precision mixed --no-cpu-baseline > $O/bench_mixed_cl3.log 2>&1
DANA_TRUNK_CHUNK=1 timeout 300 python bench.py --precision mixed --no-cpu-baseline > $O/bench_mixed_chunk1.log 2>&1
DANA_TRUNK_CHUNK=2 timeout 300 python bench.py --precision mixed --no-cpu-baseline > $O/bench_mixed_chunk2.log 2>&1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/launches_mixed.csv python tools/profile_step.py --precision mixed > $O/prof.log 2>&1
python tools/summarize_launches.py $O/launches_mixed.csv 40 > $O/launches_mixed_summary.txt 2>&1
timeout 300 python tools/gemm_bench.py --precision all > $O/gemm_bench.txt 2>&1
DANA_CLUSTER=3 timeout 300 python tools/gemm_bench.py --precision bf16x3 > $O/gemm_bench_cl3.txt 2>&1
cat $O/pytest.txt; tail -3 $O/parity.log | cut -c1-300; for f in $O/bench_*.log; do echo $f; tail -1 $f | cut -c1-200; done
