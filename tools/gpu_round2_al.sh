#!/bin/bash
# round 2, GPU call AL (1 GPU): single-dataset steps with the prior forward on part of the SMs beside the likelihood chain
mkdir -p gpurun_out
B="--no-cpu-baseline --no-gpu-baseline --no-e2e --no-parity-check --steps 100 --warmup 5"
for wl in "cfg2" "joint1024 --datasets 1"; do
  tag=$(echo $wl | tr -d ' -')
  timeout 300 python bench.py --workload $wl $B > gpurun_out/al_${tag}_base.json 2>/dev/null
  for c in 26 37 48; do
    JD_OVERLAP=2 JD_TCM_CLUSTERS=$c timeout 300 python bench.py --workload $wl $B > gpurun_out/al_${tag}_c$c.json 2>/dev/null
  done
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/al_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value=%.1f ms/step=%.4f" % (d["value"], d["ms_per_step"]))
    except Exception as exc:
        print(f, "ERR", exc)
PY
