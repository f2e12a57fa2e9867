#!/bin/bash
O=gpurun_out/r2e
mkdir -p $O
(timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "test_conv" 2>&1 | tail -3) > $O/pytest_conv.txt
cat $O/pytest_conv.txt | tail -3
for only in 1 5 11; do timeout 100 python tools/gemm_bench.py --precision bf16x3 --only $only | tail -1; DANA_DX3=0 timeout 100 python tools/gemm_bench.py --precision bf16x3 --only $only | tail -1; done
for v in 256 64 0; do DANA_DX3=$v timeout 300 python bench.py --precision mixed --no-cpu-baseline > $O/bench_mixed_dx$v.log 2>&1; echo dx3=$v; tail -1 $O/bench_mixed_dx$v.log | cut -c1-140; done
