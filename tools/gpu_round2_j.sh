#!/bin/bash
# round 2, GPU call J: micro-benchmarks v2 (mbarrier wake-up, TMEM read shapes), spin-wait build of the prior kernel
mkdir -p gpurun_out
echo "== 1. ubench"
timeout 120 tools/ubench 2>&1 | tee gpurun_out/j_ubench.txt
echo "== 2. prior forward alone: product library vs spin-wait build"
for dbg in 0 1 9; do
  JD_TC_DEBUG=$dbg timeout 120 python tools/tcm_exp.py 512 1024 2>&1 | grep "backend 3"
  JD_LIB_PATH=$PWD/jolideco_b200/libjolideco_b200_spin.so JD_TC_DEBUG=$dbg timeout 120 python tools/tcm_exp.py 512 1024 2>&1 | grep "backend 3" | sed 's/^/   spin: /'
done
echo "== 3. step with the spin build"
JD_LIB_PATH=$PWD/jolideco_b200/libjolideco_b200_spin.so timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-parity-check > gpurun_out/j_bench_joint1024_spin.json 2>/dev/null
python - <<'PY'
import json
d = json.loads(open("gpurun_out/j_bench_joint1024_spin.json").read().strip().splitlines()[-1])
print("spin build: value=%.1f ms/step=%.4f" % (d["value"], d["ms_per_step"]))
for k in (d.get("roofline_kernels") or [])[:7]:
    print("   ", k["kernel"], "us/step %.1f" % k["us_per_step"], "frac", k.get("frac"))
PY
echo "== 4. the test fixed in this commit"
timeout 600 python -m pytest tests/test_gpu_e2e.py -q -m gpu -k batched_independent -p no:cacheprovider 2>&1 | tail -3
