#!/bin/bash
O=gpurun_out/r2h
mkdir -p $O
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > $O/pytest.txt
tail -4 $O/pytest.txt
for only in 0 2 3 6 9 14; do timeout 100 python tools/gemm_bench.py --precision bf16x3 --only $only | tail -1; DANA_RES_CROSS=0 timeout 100 python tools/gemm_bench.py --precision bf16x3 --only $only | tail -1; done
timeout 300 python bench.py --no-cpu-baseline > $O/bench_mixed.log 2>&1; tail -1 $O/bench_mixed.log | cut -c1-140
DANA_RES_CROSS=0 timeout 300 python bench.py --no-cpu-baseline > $O/bench_mixed_nocross.log 2>&1; tail -1 $O/bench_mixed_nocross.log | cut -c1-140
timeout 300 python bench.py --no-cpu-baseline > $O/bench_mixed2.log 2>&1; tail -1 $O/bench_mixed2.log | cut -c1-140
