#!/bin/bash
# GPU tool: the ncu evidence kept under profiles/ -- `--set full` of every launch of one batch-1 and one batch-64 forward
# (condensed by tools/ncu_summary.py) and the time-only launch list of a short bench run.  Run from the repo root under gpurun.
mkdir -p gpurun_out
set -x
# launch 0 of the process is conv1_1_pack_kernel (weight tile of the tensor-core conv1_1), launches 1-29 the first forward
OPS="c1_2 c2_1 c2_2 c3_1 c3_2 c3_3 c4_1 c4_2 c4_3 c5_1 c5_2 c5_3 c6_1 c6_2 c6_3 c7_1 c7_2 c7_3 up8 c8_2 c8_3"
ncu --set full --clock-control none --launch-skip 30 --launch-count 29 -o gpurun_out/r02s2_b1 -f python tools/one_forward.py 1 2 > gpurun_out/r02s2_b1.log 2>&1
ncu -i gpurun_out/r02s2_b1.ncu-rep --page raw --csv > gpurun_out/r02s2_b1_raw.csv 2>> gpurun_out/r02s2_b1.log
python tools/ncu_summary.py gpurun_out/r02s2_b1_raw.csv gpurun_out/r02s2_b1_summary.csv "ncu --set full --clock-control none --launch-skip 30 --launch-count 29 python tools/one_forward.py 1 2   (HEAD of round 2, session 2: conv1_1 on the tensor cores, 128-column split-K CTA pairs; batch-1 interactive plan, dist head on its side stream + rgb)" pack+conv1_1 $OPS class softmax529 up9 c9_2 up10 c10_2 lab2rgb
ncu --set full --clock-control none --launch-skip 30 --launch-count 29 -o gpurun_out/r02s2_b64 -f python tools/one_forward.py 64 2 > gpurun_out/r02s2_b64.log 2>&1
ncu -i gpurun_out/r02s2_b64.ncu-rep --page raw --csv > gpurun_out/r02s2_b64_raw.csv 2>> gpurun_out/r02s2_b64.log
python tools/ncu_summary.py gpurun_out/r02s2_b64_raw.csv gpurun_out/r02s2_b64_summary.csv "ncu --set full --clock-control none --launch-skip 30 --launch-count 29 python tools/one_forward.py 64 2   (HEAD of round 2, session 2: 64 x 256^2, dist head + rgb; conv1_1 on the tensor cores; the throughput plan: halo tiles, CTA pairs, MT=2)" pack+conv1_1 $OPS class up9 c9_2 up10 c10_2 softmax529 lab2rgb
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02s2_launches.csv python bench.py --steps 2 --warmup 3 --skip-e2e > gpurun_out/r02s2_launch_bench.log 2>&1
rm -f gpurun_out/r02s2_b1.ncu-rep gpurun_out/r02s2_b64.ncu-rep
ls -la gpurun_out | tail -12
