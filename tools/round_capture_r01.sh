#!/bin/bash
# One GPU session that regenerates every measured artefact of profiles/ (run through gpurun; outputs under gpurun_out/cap).
O=gpurun_out/cap
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,memory.total,power.draw,clocks_event_reasons.active --format=csv > $O/gpu_state.txt 2>&1
(timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5) > $O/pytest_gpu.txt
timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench_x3.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --precision bf16 --no-cpu-baseline > $O/bench_bf16.log 2>&1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.log 2>&1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/launches_bf16x3.csv python tools/profile_step.py > $O/prof.log 2>&1
# the launch list of the bench command itself (graph replays are profiled node by node): the GEMM kernel's share of
# the launches of `bench.py` must agree with its share of the step
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
python tools/summarize_launches.py $O/launches_bench.csv 12 > $O/launches_bench_summary.txt 2>&1
gzip -f $O/launches_bench.csv
timeout 300 python tools/gemm_bench.py > $O/gemm_bench.txt 2>&1
timeout 200 python tools/roi_bench.py --sweep > $O/roi_bench.txt 2>&1
timeout 300 python tools/cisa_bench.py --iters 10 > $O/cisa_bench.txt 2>&1
timeout 100 python tools/episode_bench.py > $O/episode_bench.txt 2>&1
N="ncu --set full --clock-control none --import-source on"
timeout 200 $N -k regex:roi_align7 --launch-skip 2 -c 1 -o $O/ncu_roi_f32_mix python tools/roi_bench.py --iters 1 --only f32 > $O/ncu1.log 2>&1
timeout 200 $N -k regex:roi_align7 --launch-skip 2 -c 1 -o $O/ncu_roi_f32_64px python tools/roi_bench.py --iters 1 --only f32 --side 64 > $O/ncu2.log 2>&1
timeout 300 $N -k regex:conv_gemm --launch-skip 12 -c 4 -o $O/ncu_cisa_u100 python tools/cisa_bench.py --eager --iters 1 --units 100 --ns 196 --precision bf16x3 > $O/ncu3.log 2>&1
timeout 200 $N -k regex:conv_gemm --launch-skip 3 -c 1 -o $O/ncu_gemm_rpn_x3 python tools/gemm_bench.py --only 7 --precision bf16x3 --iters 1 > $O/ncu4.log 2>&1
timeout 200 $N -k regex:conv_gemm --launch-skip 3 -c 1 -o $O/ncu_gemm_l1c3_x3 python tools/gemm_bench.py --only 2 --precision bf16x3 --iters 1 > $O/ncu5.log 2>&1
timeout 200 $N -k regex:proposals_nms --launch-skip 2 -c 1 -o $O/ncu_prop_nms python tools/profile_step.py > $O/ncu6.log 2>&1
# the .ncu-rep files (20 MB each with sources) do not fit the 64 MiB return channel: extract what profiles/ keeps
python tools/ncu_summary.py $O/ncu_summary.md $O/*.ncu-rep > /dev/null 2>&1
for r in $O/*.ncu-rep; do
  b=${r%.ncu-rep}
  ncu -i $r --page raw --csv > $b.raw.csv 2>/dev/null
  python tools/ncu_sass_hot.py $r > $b.hot.txt 2>&1
done
# the CISA capture holds four GEMM launches: k-proj, q-proj, logits+softmax, P.V
for i in 0 1 2 3; do python tools/ncu_sass_hot.py $O/ncu_cisa_u100.ncu-rep conv_gemm $i > $O/ncu_cisa_u100.launch$i.hot.txt 2>&1; done
rm -f $O/*.ncu-rep
cat $O/pytest_gpu.txt; tail -1 $O/bench_x3.log | cut -c1-160; tail -1 $O/bench_bf16.log | cut -c1-160; tail -1 $O/bench_ref.log | cut -c1-160; ls $O
