#!/bin/bash
# round 2, GPU call K: first run of the two-tiles-per-CTA prior forward (backend 4)
mkdir -p gpurun_out
echo "== 1. parity of backend 4"
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -q -m gpu --tb=short -p no:cacheprovider -x \
   -k "two_tiles or value_and_grad or prior_golden or tensor_core or benchmark_size" 2>&1 | grep -v "^$" | tail -25 > gpurun_out/k_pytest.log
tail -25 gpurun_out/k_pytest.log
echo "== 2. prior forward alone"
for dbg in 0 1 9; do
  JD_TC_DEBUG=$dbg timeout 120 python tools/tcm_exp.py 512 1024 2>&1 | grep "backend"
done
echo "== 3. step"
for b in 3 4; do
JD_PRIOR_BACKEND=$b timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-parity-check > gpurun_out/k_bench_joint1024_b$b.json 2>/dev/null
JD_PRIOR_BACKEND=$b timeout 300 python bench.py --workload cfg2 --steps 50 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-parity-check > gpurun_out/k_bench_cfg2_b$b.json 2>/dev/null
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/k_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value=%.1f ms/step=%.4f" % (d["value"], d["ms_per_step"]))
        for k in (d.get("roofline_kernels") or [])[:7]:
            print("   ", k["kernel"], "us/step %.1f" % k["us_per_step"], "frac", k.get("frac"))
    except Exception as exc:
        print(f, "ERR", exc)
PY
