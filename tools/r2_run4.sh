#!/bin/bash
O=gpurun_out/r2d
mkdir -p $O
(timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "test_conv or test_linear" 2>&1 | tail -15) > $O/pytest_conv.txt
cat $O/pytest_conv.txt | tail -5
timeout 300 python tools/gemm_bench.py --precision bf16x3 > $O/gemm_bench_dx3.txt 2>&1
DANA_DX3=0 timeout 300 python tools/gemm_bench.py --precision bf16x3 > $O/gemm_bench_nodx3.txt 2>&1
paste $O/gemm_bench_dx3.txt $O/gemm_bench_nodx3.txt | cut -c1-200
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > $O/pytest.txt
cat $O/pytest.txt
timeout 300 python bench.py --precision mixed --no-cpu-baseline > $O/bench_mixed.log 2>&1
tail -1 $O/bench_mixed.log | cut -c1-200
