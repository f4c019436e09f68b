#!/bin/bash
# round 2, GPU call Y: row-wise fold in the joint update kernel, smoke(), bench
mkdir -p gpurun_out
echo "== 1. tests"
timeout 900 python -m pytest tests/test_gpu_likelihood.py tests/test_gpu_e2e.py tests/test_gpu_fullsize.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | grep -v "^$" | tail -8
echo "== 2. smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== 3. bench"
B="--steps 30 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-parity-check"
timeout 300 python bench.py $B > gpurun_out/y_bench_joint1024.json 2>/dev/null
timeout 300 python bench.py $B --workload cfg4 --steps 10 > gpurun_out/y_bench_cfg4.json 2>/dev/null
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/y_bench_*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split("/")[-1], "value=%.1f ms/step=%.4f" % (d["value"], d["ms_per_step"]))
    for k in (d.get("roofline_kernels") or [])[:7]:
        print("   ", k["kernel"], "us/step %.1f" % k["us_per_step"], "frac", k.get("frac"), k.get("bound"))
PY
