#!/bin/bash
# round 2, GPU call T: bucketed max-mode backward (rewritten), 4 x 8 likelihood tiles, full suite
mkdir -p gpurun_out
echo "== 1. GPU suite"
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider 2>&1 | grep -v "^$" | tail -30 > gpurun_out/t_pytest.log
tail -6 gpurun_out/t_pytest.log
echo "== 2. A/B"
B="--steps 30 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-parity-check"
run() { name=$1; shift; env "$@" timeout 300 python bench.py $B $EXTRA > gpurun_out/t_$name.json 2>/dev/null; }
EXTRA="" run joint_default X=1
EXTRA="" run joint_nobucket JD_BWD_BUCKETED=0
EXTRA="" run joint_rt4 JD_LIK_RT=4
EXTRA="--datasets 1" run d1_rt8 JD_LIK_RT=8
EXTRA="--datasets 1" run d1_rt4 JD_LIK_RT=4
EXTRA="--datasets 2" run d2_rt8 JD_LIK_RT=8
EXTRA="--datasets 2" run d2_rt4 JD_LIK_RT=4
EXTRA="--workload cfg2" run cfg2 X=1
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/t_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value=%.1f ms/step=%.4f" % (d["value"], d["ms_per_step"]))
        for k in (d.get("roofline_kernels") or [])[:7]:
            print("   ", k["kernel"], "us/step %.1f" % k["us_per_step"], "frac", k.get("frac"))
    except Exception as exc:
        print(f, "ERR", exc)
PY
