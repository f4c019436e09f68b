#!/bin/bash
# ncu evidence for profiles/: launch lists of the bench steps and --set full captures of the dominant kernels.
# Usage (on the GPU box): bash tools/profile_round.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_cfg2_$tag.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_cfg2_$tag.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_joint1024_$tag.csv \
    python bench.py --workload joint1024 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_joint_$tag.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:gmm_fwd_tc -s 3 -c 1 -f -o gpurun_out/prof_gmm_fwd_$tag \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_gmm_$tag.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:conv3 -s 2 -c 2 -f -o gpurun_out/prof_conv3_$tag \
    python tools/conv_one.py > gpurun_out/ncu_conv3_$tag.log 2>&1
ls -la gpurun_out/*$tag*
