#!/bin/bash
# round 2, GPU call AM (1 GPU): auto-tuned split of one-dataset steps (prior forward on part of the SMs): GPU suite, bench lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider 2>&1 | grep -v "^$" | tail -15 > gpurun_out/am_pytest.log
tail -3 gpurun_out/am_pytest.log
B="--no-cpu-baseline --no-gpu-baseline --no-parity-check --steps 100 --warmup 5"
timeout 300 python bench.py --workload cfg2 $B > gpurun_out/am_cfg2.json 2>/dev/null
JD_SPLIT_CLUSTERS=0 timeout 300 python bench.py --workload cfg2 $B > gpurun_out/am_cfg2_nosplit.json 2>/dev/null
timeout 300 python bench.py --datasets 1 $B --no-e2e > gpurun_out/am_d1.json 2>/dev/null
timeout 300 python bench.py $B > gpurun_out/am_joint1024.json 2>/dev/null
timeout 300 python bench.py --workload cfg5 $B > gpurun_out/am_cfg5.json 2>/dev/null
JD_SPLIT_CLUSTERS=0 timeout 300 python bench.py --workload cfg5 $B > gpurun_out/am_cfg5_nosplit.json 2>/dev/null
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/am_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        c = d["config"]
        print(f.split("/")[-1], "value=%.1f ms/step=%.4f e2e=%s pairs=%s tuning=%s" % (
            d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value"), c.get("prior_forward_sm_pairs"), c.get("split_tuning_ms")))
    except Exception as exc:
        print(f, "ERR", exc)
PY
