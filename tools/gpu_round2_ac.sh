#!/bin/bash
# round 2, GPU call AC: mw prefetch in the two-tile epilogue (non-zero-mean mixtures)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -q -m gpu --tb=short -p no:cacheprovider -x 2>&1 | tail -3
B="--steps 30 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-parity-check"
timeout 300 python bench.py $B --gmm-mean-scale 0.01 > gpurun_out/ac_n1_nonzero_mean.json 2>/dev/null
timeout 300 python bench.py $B --gmm-mean-scale 0.01 --backend 5 > gpurun_out/ac_n1_nonzero_mean_b5.json 2>/dev/null
timeout 300 python bench.py $B --gmm-mean-scale 0.01 --backend 3 > gpurun_out/ac_n1_nonzero_mean_b3.json 2>/dev/null
timeout 300 python bench.py $B > gpurun_out/ac_n1.json 2>/dev/null
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/ac_n*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split("/")[-1], "value=%.1f ms/step=%.4f" % (d["value"], d["ms_per_step"]))
    for k in (d.get("roofline_kernels") or [])[:1]:
        print("   ", k["kernel"], "us/step %.1f" % k["us_per_step"], "frac", k.get("frac"), k.get("bound"))
PY
