#!/bin/bash
# round 2, GPU call G: e2e suite after the launch-size heuristic, trimmed-B-copy A/B inside the step, cfg2 back on FFT
mkdir -p gpurun_out
echo "== 1. GPU suite"
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider 2>&1 | grep -v "^$" | tail -30 > gpurun_out/g_pytest.log
tail -12 gpurun_out/g_pytest.log
echo "== 2. A/B trimmed copy (no overlap)"
for d in 0 32; do
  JD_TC_DEBUG=$d JD_OVERLAP=0 timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-parity-check > gpurun_out/g_bench_joint1024_dbg$d.json 2>/dev/null
done
echo "== 3. cfg2, cfg5"
timeout 300 python bench.py --workload cfg2 --steps 50 --no-cpu-baseline --no-gpu-baseline --no-parity-check > gpurun_out/g_bench_cfg2.json 2>/dev/null
timeout 300 python bench.py --workload cfg5 --steps 20 > gpurun_out/g_bench_cfg5.json 2>/dev/null
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/g_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f.split("/")[-1], "value=%.1f ms/step=%.4f e2e=%s frac=%s" % (
            d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value"), r.get("frac")))
        for k in (d.get("roofline_kernels") or [])[:7]:
            print("   ", k["kernel"], "us/step %.1f" % k["us_per_step"], "frac", k.get("frac"), k.get("bound"))
    except Exception as exc:
        print(f, "ERR", exc)
PY
