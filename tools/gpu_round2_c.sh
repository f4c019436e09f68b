#!/bin/bash
# round 2, GPU call C: the mixed TF32/FP16 persistent prior kernel (backend 3): parity, then speed
mkdir -p gpurun_out
echo "== 1. backend 3 parity (small goldens, then benchmark sizes)"
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu --tb=short -p no:cacheprovider -x -k "golden or value_and_grad or vs_cuda_core" 2>&1 | tail -15
timeout 600 python -m pytest tests/test_gpu_fullsize.py -q -m gpu --tb=short -p no:cacheprovider -s -k "tcm or joint" 2>&1 | tail -25
echo "== 2. speed: backend 1 vs 3, cfg2 and joint1024 (no overlap, to read the kernel alone)"
for b in 1 3; do
  for w in cfg2 joint1024; do
    JD_OVERLAP=0 timeout 300 python bench.py --workload $w --backend $b --steps 30 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-parity-check > gpurun_out/c_bench_${w}_b$b.json 2> gpurun_out/c_bench_${w}_b$b.err || tail -5 gpurun_out/c_bench_${w}_b$b.err
  done
done
timeout 300 python bench.py --backend 3 --steps 30 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-parity-check > gpurun_out/c_bench_joint1024_b3_overlap.json 2>/dev/null
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/c_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f.split("/")[-1], "value=%.1f ms/step=%.4f frac=%s issued=%s" % (d["value"], d["ms_per_step"], r.get("frac"), r.get("issued_frac")))
        for k in (d.get("roofline_kernels") or [])[:5]:
            print("   ", k["kernel"], "us/step %.1f" % k["us_per_step"], "frac", k.get("frac"), k.get("bound"))
    except Exception as exc:
        print(f, "ERR", exc)
PY
