#!/bin/bash
# round 2, GPU call Z: trace reuse after joint steps (suite), e2e profile of the north-star run
mkdir -p gpurun_out
echo "== 1. GPU suite"
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider 2>&1 | grep -v "^$" | tail -30 > gpurun_out/z_pytest.log
tail -5 gpurun_out/z_pytest.log
echo "== 2. e2e profile joint1024, 30 epochs"
timeout 300 python tools/e2e_profile.py joint1024 30 2>&1 | head -70 | cut -c1-180 > gpurun_out/z_e2e_profile.txt
head -60 gpurun_out/z_e2e_profile.txt
echo "== 3. bench e2e"
timeout 600 python bench.py --steps 100 --no-cpu-baseline --no-gpu-baseline --no-parity-check > gpurun_out/z_bench_joint1024.json 2>/dev/null
python - <<'PY'
import json
d = json.loads(open("gpurun_out/z_bench_joint1024.json").read().strip().splitlines()[-1])
print("value=%.1f ms/step=%.4f e2e=%s" % (d["value"], d["ms_per_step"], d.get("e2e")))
PY
