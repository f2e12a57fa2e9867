#!/bin/bash
O=gpurun_out/r2b
mkdir -p $O
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > $O/pytest.txt
timeout 900 python tools/parity_report.py --out $O/parity.jsonl > $O/parity.log 2>&1
N="ncu --set full --clock-control none --import-source on"
timeout 200 $N -k regex:conv_gemm --launch-skip 3 -c 1 -o $O/ncu_l1c2 python tools/gemm_bench.py --only 1 --precision bf16x3 --iters 1 > $O/ncu1.log 2>&1
timeout 200 $N -k regex:conv_gemm --launch-skip 3 -c 1 -o $O/ncu_l1c3 python tools/gemm_bench.py --only 2 --precision bf16x3 --iters 1 > $O/ncu2.log 2>&1
timeout 200 python tools/roi_bench.py > $O/roi_bench.txt 2>&1
ls -la $O
cat $O/pytest.txt | tail -8; tail -12 $O/parity.log | cut -c1-400
