#!/bin/bash
O=gpurun_out/r2g
mkdir -p $O
(timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -x -q 2>&1 | tail -40) > $O/pytest_train.txt
tail -40 $O/pytest_train.txt
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > $O/pytest.txt
tail -8 $O/pytest.txt
