#!/bin/bash
O=gpurun_out/r2k
mkdir -p $O
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > $O/pytest.txt
tail -4 $O/pytest.txt
timeout 300 python tools/cisa_bench.py --iters 10 --ns 400 --units 1,3,6,25,100 --precision bf16x3 > $O/cisa_wide.txt 2>&1; cat $O/cisa_wide.txt
timeout 300 ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum --clock-control none -k regex:conv_gemm --launch-skip 4 -c 4 python tools/cisa_bench.py --eager --iters 1 --units 100 --ns 400 --precision bf16x3 2>&1 | grep -E "conv_gemm|duration|pipe_tensor" | head -20
timeout 300 python bench.py --no-cpu-baseline > $O/bench_mixed.log 2>&1; tail -1 $O/bench_mixed.log | cut -c1-140
