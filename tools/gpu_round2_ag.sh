#!/bin/bash
# round 2, GPU call AG: 8 x 8-tile bucket kernel of the max-mode backward (correctness, A/B against the first one),
# memcheck of the two-tile prior forward
mkdir -p gpurun_out
echo "== 1. tests"
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -x -q -m gpu -k "bucket or backward or value_and_grad or fullsize or full_size" 2>&1 | tail -3
echo "== 2. A/B"
B="--no-cpu-baseline --no-gpu-baseline --no-e2e --no-parity-check --steps 100 --warmup 5"
JD_BWD_BUCKET_V1=1 timeout 300 python bench.py $B > gpurun_out/ag_bench_v1.json 2>/dev/null
timeout 300 python bench.py $B > gpurun_out/ag_bench_v2.json 2>/dev/null
python - <<'PY'
import json
for t in ("v1", "v2"):
    d = json.loads(open(f"gpurun_out/ag_bench_{t}.json").read().strip().splitlines()[-1])
    k = {x["kernel"]: round(x["us_per_step"], 1) for x in d.get("roofline_kernels") or []}
    print(t, "value=%.1f ms/step=%.4f" % (d["value"], d["ms_per_step"]), k)
PY
echo "== 3. memcheck of the two-tile forward (backends 4 and 5) and the new bucket kernel"
S="compute-sanitizer --launch-timeout 0 --print-limit 20 --tool memcheck"
timeout 400 $S python -m pytest -x -q -m gpu -p no:cacheprovider \
  "tests/test_gpu_kernels.py::test_gmm_prior_value_and_grad[shape0-shift0-False-4]" \
  "tests/test_gpu_kernels.py::test_gmm_prior_value_and_grad[shape2-shift2-True-4]" \
  "tests/test_gpu_kernels.py::test_gmm_prior_value_and_grad[shape1-shift1-False-5]" \
  "tests/test_gpu_kernels.py::test_gmm_prior_value_and_grad[shape1-shift1-False-2]" > gpurun_out/ag_tcm2_mem.log 2>&1
echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|error" gpurun_out/ag_tcm2_mem.log | tail -4
timeout 300 compute-sanitizer --launch-timeout 0 --print-limit 20 --tool racecheck python -m pytest -x -q -m gpu -p no:cacheprovider \
  "tests/test_gpu_kernels.py::test_gmm_prior_value_and_grad[shape1-shift1-False-2]" > gpurun_out/ag_bwd8_race.log 2>&1
echo "rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|error" gpurun_out/ag_bwd8_race.log | tail -4
