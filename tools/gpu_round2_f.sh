#!/bin/bash
# round 2, GPU call F: suite, prior-kernel A/B after the trimmed B copy, likelihood staging change, benches of all configs
mkdir -p gpurun_out
echo "== 1. GPU suite"
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider 2>&1 | grep -v "^$" | tail -40 > gpurun_out/f_pytest.log
tail -25 gpurun_out/f_pytest.log
echo "== 2. prior kernel alone"
timeout 120 python tools/tcm_exp.py 512 1024 2>&1 | grep backend
echo "== 3. likelihood kernels alone"
timeout 200 python tools/lik_one.py
echo "== 4. benches"
timeout 900 python bench.py --steps 30 --breakdown > gpurun_out/f_bench_joint1024.json 2> gpurun_out/f_bench_joint1024.err
JD_OVERLAP=0 timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-parity-check > gpurun_out/f_bench_joint1024_nooverlap.json 2>/dev/null
for w in cfg2 cfg3 cfg4; do
  timeout 400 python bench.py --workload $w --steps 20 --no-cpu-baseline --no-gpu-baseline --no-parity-check > gpurun_out/f_bench_$w.json 2>/dev/null
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/f_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f.split("/")[-1], "value=%.1f ms/step=%.4f e2e=%s frac=%s" % (
            d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value"), r.get("frac")))
        if d.get("parity_check"): print("   parity:", d["parity_check"]["status"])
        if d.get("cpu_baseline"): print("   cpu:", d["cpu_baseline"]["value"], "gpu:", (d.get("gpu_baseline") or {}).get("value"))
        for k in (d.get("roofline_kernels") or [])[:7]:
            print("   ", k["kernel"], "us/step %.1f" % k["us_per_step"], "frac", k.get("frac"), k.get("bound"))
    except Exception as exc:
        print(f, "ERR", exc)
PY
