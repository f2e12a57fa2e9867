#!/bin/bash
O=gpurun_out/r2f
mkdir -p $O
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > $O/pytest.txt
tail -8 $O/pytest.txt
timeout 300 python bench.py --no-cpu-baseline > $O/bench_mixed.log 2>&1
tail -1 $O/bench_mixed.log | cut -c1-160
DANA_SIDE_STREAM=0 timeout 300 python bench.py --no-cpu-baseline > $O/bench_mixed_noside.log 2>&1
tail -1 $O/bench_mixed_noside.log | cut -c1-160
DANA_WIDE_SOFTMAX=0 timeout 300 python bench.py --no-cpu-baseline > $O/bench_mixed_nowide.log 2>&1
tail -1 $O/bench_mixed_nowide.log | cut -c1-160
timeout 300 python tools/cisa_bench.py --iters 10 --ns 400 --units 1,3,6,25,100 --precision bf16x3 > $O/cisa_wide.txt 2>&1
DANA_WIDE_SOFTMAX=0 timeout 300 python tools/cisa_bench.py --iters 10 --ns 400 --units 1,3,6,25,100 --precision bf16x3 > $O/cisa_nowide.txt 2>&1
cat $O/cisa_wide.txt $O/cisa_nowide.txt
