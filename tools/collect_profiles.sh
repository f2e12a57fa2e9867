#!/bin/bash
# Copies the summaries of a round-2 capture (tools/round_capture.sh -> gpurun_out/cap2) into profiles/r02_* (tracked).
S=${1:-gpurun_out/cap2}
D=profiles
mkdir -p $D/r02_ncu
: > $D/r02_bench_lines.jsonl
for f in bench_mixed bench_mixed_sustained bench_x3 bench_bf16 bench_ref; do
  [ -f $S/$f.log ] && tail -1 $S/$f.log >> $D/r02_bench_lines.jsonl
done
cp $S/parity.jsonl $D/r02_parity.jsonl
cp $S/launches_mixed.csv $D/r02_launches_mixed.csv
cp $S/launches_bf16x3.csv $D/r02_launches_bf16x3.csv
cp $S/launches_mixed_summary.txt $S/launches_bf16x3_summary.txt $S/launches_bench_summary.txt $D/ 2>/dev/null
for f in launches_mixed_summary launches_bf16x3_summary launches_bench_summary; do mv $D/$f.txt $D/r02_$f.txt; done
cp $S/launches_bench.csv.gz $D/r02_launches_bench.csv.gz
cp $S/gemm_bench.txt $D/r02_gemm_bench.txt
cp $S/roi_bench.txt $D/r02_roi_bench.txt
cp $S/cisa_sweep.json $D/r02_cisa_sweep.json
cp $S/episode_bench.txt $D/r02_episode_bench.txt
cp $S/gpu_state.txt $D/r02_gpu_state.txt
cp $S/pytest_gpu.txt $D/r02_pytest_gpu.txt
: > $D/r02_train_bench_lines.jsonl
for f in train_bench train_bench_eager train_bench_res50_bs4; do
  [ -f $S/$f.log ] && tail -1 $S/$f.log >> $D/r02_train_bench_lines.jsonl
done
grep -A40 "GPU kernels" $S/train_profile.txt > $D/r02_train_profile.txt 2>/dev/null
cp $S/bwd_bench.txt $D/r02_bwd_bench.txt 2>/dev/null
cp $S/ncu_summary.md $D/r02_ncu_summary.md
cp $S/ncu_*.metrics.txt $S/ncu_*.hot.txt $D/r02_ncu/ 2>/dev/null
python tools/gemm_traffic.py $S/launches_mixed.csv mixed $D/r02_gemm_traffic_mixed.json > /dev/null
python tools/gemm_traffic.py $S/launches_bf16x3.csv bf16x3 $D/r02_gemm_traffic_bf16x3.json > /dev/null
ls $D | head -80
