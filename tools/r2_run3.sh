#!/bin/bash
O=gpurun_out/r2c
mkdir -p $O
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > $O/pytest.txt
timeout 300 python bench.py --precision mixed --no-cpu-baseline > $O/bench_mixed.log 2>&1
timeout 600 python tools/cisa_sweep.py --out $O/cisa_sweep.json > $O/cisa_sweep.log 2>&1
timeout 900 ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum --clock-control none -k regex:conv_gemm --csv --log-file $O/cisa_sweep_ncu.csv python tools/cisa_sweep.py --ncu-pass > $O/cisa_ncu.log 2>&1
python tools/cisa_sweep.py --merge $O/cisa_sweep.json $O/cisa_sweep_ncu.csv > $O/merge.log 2>&1
cat $O/pytest.txt | tail -8; tail -1 $O/bench_mixed.log | cut -c1-250; tail -3 $O/cisa_sweep.log; cat $O/merge.log
