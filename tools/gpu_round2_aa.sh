#!/bin/bash
# round 2, GPU call AA: packed (fma.rn.f32x2) 4 x 8 likelihood kernels
mkdir -p gpurun_out
echo "== 1. GPU suite"
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider 2>&1 | grep -v "^$" | tail -30 > gpurun_out/aa_pytest.log
tail -5 gpurun_out/aa_pytest.log
echo "== 2. bench"
B="--steps 30 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-parity-check"
timeout 300 python bench.py $B > gpurun_out/aa_bench_joint1024.json 2>/dev/null
JD_LIK_RT=8 timeout 300 python bench.py $B > gpurun_out/aa_bench_joint1024_rt8.json 2>/dev/null
timeout 300 python bench.py $B --datasets 1 > gpurun_out/aa_bench_d1.json 2>/dev/null
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/aa_bench_*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split("/")[-1], "value=%.1f ms/step=%.4f" % (d["value"], d["ms_per_step"]))
    for k in (d.get("roofline_kernels") or [])[:7]:
        print("   ", k["kernel"], "us/step %.1f" % k["us_per_step"], "frac", k.get("frac"), k.get("bound"))
PY
