#!/bin/bash
O=gpurun_out/r2i
mkdir -p $O
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > $O/pytest.txt
tail -4 $O/pytest.txt
timeout 300 python bench.py --no-cpu-baseline > $O/bench_mixed.log 2>&1; tail -1 $O/bench_mixed.log | cut -c1-140
DANA_SIDE_STREAM=0 timeout 300 python bench.py --no-cpu-baseline > $O/bench_mixed_noside.log 2>&1; tail -1 $O/bench_mixed_noside.log | cut -c1-140
timeout 300 python bench.py --no-cpu-baseline > $O/bench_mixed2.log 2>&1; tail -1 $O/bench_mixed2.log | cut -c1-140
