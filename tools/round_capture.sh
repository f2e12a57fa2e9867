#!/bin/bash
# One GPU session that regenerates every measured artefact of profiles/ for round 2 (run through gpurun; outputs under
# gpurun_out/cap2; tools/collect_profiles.sh copies the summaries into profiles/r02_*).
O=gpurun_out/cap2
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,memory.total,power.draw,clocks_event_reasons.active --format=csv > $O/gpu_state.txt 2>&1
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6) > $O/pytest_gpu.txt
timeout 900 python tools/parity_report.py --out $O/parity.jsonl > $O/parity.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench_mixed.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --precision bf16x3 --no-cpu-baseline > $O/bench_x3.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --precision bf16 --no-cpu-baseline > $O/bench_bf16.log 2>&1
timeout 300 python bench.py --steps 1000 --warmup 20 --no-cpu-baseline > $O/bench_mixed_sustained.log 2>&1
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.log 2>&1
for prec in mixed bf16x3; do
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/launches_$prec.csv python tools/profile_step.py --precision $prec > $O/prof_$prec.log 2>&1
python tools/summarize_launches.py $O/launches_$prec.csv 40 > $O/launches_${prec}_summary.txt 2>&1
done
# the launch list of the bench command itself (graph replays are profiled node by node)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
python tools/summarize_launches.py $O/launches_bench.csv 14 > $O/launches_bench_summary.txt 2>&1
gzip -f $O/launches_bench.csv
timeout 300 python tools/gemm_bench.py --precision all > $O/gemm_bench.txt 2>&1
timeout 200 python tools/roi_bench.py --sweep > $O/roi_bench.txt 2>&1
timeout 600 python tools/cisa_sweep.py --out $O/cisa_sweep.json > $O/cisa_sweep.log 2>&1
timeout 900 ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum --clock-control none -k regex:conv_gemm --csv --log-file $O/cisa_sweep_ncu.csv python tools/cisa_sweep.py --ncu-pass > $O/cisa_ncu.log 2>&1
python tools/cisa_sweep.py --merge $O/cisa_sweep.json $O/cisa_sweep_ncu.csv > $O/cisa_merge.log 2>&1
timeout 100 python tools/episode_bench.py > $O/episode_bench.txt 2>&1
# training step (BASELINE configs[3]): captured and eager lines, kernel profile of one captured step, backward kernels alone
timeout 300 python tools/train_bench.py --steps 20 > $O/train_bench.log 2>&1
timeout 300 python tools/train_bench.py --steps 10 --eager > $O/train_bench_eager.log 2>&1
timeout 300 python tools/train_bench.py --steps 10 --layers 50 --batch 4 --height 600 --width 1000 --shots 3 > $O/train_bench_res50_bs4.log 2>&1
timeout 300 python tools/train_profile.py > $O/train_profile.txt 2>&1
timeout 300 python tools/bwd_bench.py > $O/bwd_bench.txt 2>&1
N="ncu --set full --clock-control none --import-source on"
timeout 200 $N -k regex:roi_align7_kernel --launch-skip 2 -c 1 -o $O/ncu_roi_head16_mix python tools/roi_bench.py --iters 1 --only head16 > $O/ncu1.log 2>&1
timeout 200 $N -k regex:roi_align7_bwd --launch-skip 1 -c 1 -o $O/ncu_roi_bwd python tools/roi_bench.py --iters 1 > $O/ncu2.log 2>&1
for u in 3 6 100; do
timeout 300 $N -k regex:conv_gemm --launch-skip 4 -c 4 -o $O/ncu_cisa_ns400_u$u python tools/cisa_bench.py --eager --iters 1 --units $u --ns 400 --precision bf16x3 > $O/ncu3_$u.log 2>&1
done
timeout 200 $N -k regex:conv_gemm --launch-skip 3 -c 1 -o $O/ncu_gemm_rpn_f16 python tools/gemm_bench.py --only 7 --precision f16 --iters 1 > $O/ncu4.log 2>&1
timeout 200 $N -k regex:conv_gemm --launch-skip 3 -c 1 -o $O/ncu_gemm_l1c2_dx3 python tools/gemm_bench.py --only 1 --precision bf16x3 --iters 1 > $O/ncu5.log 2>&1
timeout 200 $N -k regex:conv_gemm --launch-skip 3 -c 1 -o $O/ncu_gemm_l1c3_x3 python tools/gemm_bench.py --only 2 --precision bf16x3 --iters 1 > $O/ncu6.log 2>&1
timeout 200 $N -k regex:grad_prepare --launch-skip 2 -c 1 -o $O/ncu_bwd_grad_prepare python tools/bwd_bench.py --iters 1 --only 2 > $O/ncu7.log 2>&1
timeout 200 $N -k regex:im2col_t --launch-skip 2 -c 1 -o $O/ncu_bwd_im2col_t python tools/bwd_bench.py --iters 1 --only 1 > $O/ncu8.log 2>&1
timeout 200 $N -k regex:conv_gemm --launch-skip 3 -c 1 -o $O/ncu_bwd_wgrad_l3c2 python tools/bwd_bench.py --iters 1 --only 1 > $O/ncu9.log 2>&1
# the .ncu-rep files (20 MB each with sources) do not fit the 64 MiB return channel: extract what profiles/ keeps
python tools/ncu_summary.py $O/ncu_summary.md $O/*.ncu-rep > /dev/null 2>&1
for r in $O/*.ncu-rep; do
  b=${r%.ncu-rep}
  ncu -i $r --page raw --csv > $b.raw.csv 2>/dev/null
  python tools/ncu_metrics_extract.py $b.raw.csv $b.metrics.txt > /dev/null 2>&1
  rm -f $b.raw.csv
  python tools/ncu_sass_hot.py $r > $b.hot.txt 2>&1
done
for u in 3 6 100; do for i in 0 1 2 3; do python tools/ncu_sass_hot.py $O/ncu_cisa_ns400_u$u.ncu-rep conv_gemm $i > $O/ncu_cisa_ns400_u$u.launch$i.hot.txt 2>&1; done; done
rm -f $O/*.ncu-rep
cat $O/pytest_gpu.txt; for f in $O/bench_*.log; do echo $f; tail -1 $f | cut -c1-220; done; ls $O | head -80
