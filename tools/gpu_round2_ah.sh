#!/bin/bash
# round 2, GPU call AH: second-generation bucket kernel (8 x 4 tiles, cp.async gather, device work counter): tests and A/B (JD_BWD_BUCKET_V1=1 = first kernel)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -x -q -m gpu -k "bucket or backward or value_and_grad or fullsize" 2>&1 | tail -3
B="--no-cpu-baseline --no-gpu-baseline --no-e2e --no-parity-check --steps 100 --warmup 5"
JD_BWD_BUCKET_V1=1 timeout 300 python bench.py $B > gpurun_out/ah_bench_v1.json 2>/dev/null
timeout 300 python bench.py $B > gpurun_out/ah_bench_v2.json 2>/dev/null
python - <<'PY'
import json
for t in ("v1", "v2"):
    d = json.loads(open(f"gpurun_out/ah_bench_{t}.json").read().strip().splitlines()[-1])
    k = {x["kernel"]: round(x["us_per_step"], 1) for x in d.get("roofline_kernels") or []}
    print(t, "value=%.1f ms/step=%.4f" % (d["value"], d["ms_per_step"]), k)
PY
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"gmm_bwd_bucket8" -s 4 -c 1 -f -o gpurun_out/prof_bwd_bucket8 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-parity-check --no-graph > /dev/null 2>&1
