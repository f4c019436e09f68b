#!/bin/bash
# round 2, GPU call S: full GPU suite with backend 4 as default, north-star bench (all legs), cfg2
mkdir -p gpurun_out
echo "== 1. GPU suite"
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider 2>&1 | grep -v "^$" | tail -30 > gpurun_out/s_pytest.log
tail -6 gpurun_out/s_pytest.log
echo "== 2. north-star bench, default arguments"
timeout 900 python bench.py > gpurun_out/s_bench_default.json 2> gpurun_out/s_bench_default.err
tail -c 300 gpurun_out/s_bench_default.err
echo "== 3. backend 5, cfg2"
timeout 300 python bench.py --steps 30 --backend 5 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-parity-check > gpurun_out/s_bench_joint1024_b5.json 2>/dev/null
timeout 300 python bench.py --workload cfg2 --steps 50 --no-cpu-baseline --no-gpu-baseline > gpurun_out/s_bench_cfg2.json 2>/dev/null
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/s_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f.split("/")[-1], "value=%.1f ms/step=%.4f e2e=%s frac=%s" % (
            d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value"), r.get("frac")))
        print("   parity:", d.get("parity_check"))
        print("   cpu:", d.get("cpu_baseline"), "gpu:", d.get("gpu_baseline"))
        for k in (d.get("roofline_kernels") or [])[:7]:
            print("   ", k["kernel"], "us/step %.1f" % k["us_per_step"], "frac", k.get("frac"), k.get("bound"))
    except Exception as exc:
        print(f, "ERR", exc)
PY
