#!/bin/bash
O=gpurun_out/r2l
mkdir -p $O
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > $O/pytest.txt
tail -4 $O/pytest.txt
timeout 300 python tools/gemm_bench.py --precision bf16x3 > $O/gemm_wide.txt 2>&1
DANA_WIDE=0 timeout 300 python tools/gemm_bench.py --precision bf16x3 > $O/gemm_nowide.txt 2>&1
paste $O/gemm_wide.txt $O/gemm_nowide.txt | cut -c1-60,85-130
timeout 300 python bench.py --no-cpu-baseline > $O/bench_mixed.log 2>&1; tail -1 $O/bench_mixed.log | cut -c1-140
DANA_WIDE=0 timeout 300 python bench.py --no-cpu-baseline > $O/bench_mixed_nowide.log 2>&1; tail -1 $O/bench_mixed_nowide.log | cut -c1-140
timeout 300 python bench.py --no-cpu-baseline > $O/bench_mixed2.log 2>&1; tail -1 $O/bench_mixed2.log | cut -c1-140
