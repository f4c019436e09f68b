#!/bin/bash
# round 2, GPU call I: micro-benchmarks (TMEM read, bulk-TMA ingest per SM), GPU suite, ncu evidence of the north-star step
mkdir -p gpurun_out
echo "== 1. ubench"
timeout 120 tools/ubench 2>&1 | tee gpurun_out/i_ubench.txt
echo "== 2. GPU suite"
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider 2>&1 | grep -v "^$" | tail -30 > gpurun_out/i_pytest.log
tail -5 gpurun_out/i_pytest.log
echo "== 3. ncu launch list + full captures (joint1024)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/i_launches_joint1024.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-parity-check --no-graph > gpurun_out/i_ncu_bench_joint.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"gmm_fwd_tcm|lik_kernel|gmm_bwd|joint_grad" -s 10 -c 5 -f -o gpurun_out/prof_joint1024_r02 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-parity-check --no-graph > gpurun_out/i_ncu_full.log 2>&1
tail -3 gpurun_out/i_ncu_full.log
ls -la gpurun_out/prof_joint1024_r02* gpurun_out/i_*
