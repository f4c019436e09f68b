#!/bin/bash
# round 2, GPU call A: full GPU suite (new batched likelihood kernels, full-size oracle parity, un-skipped shift tests),
# then the north-star bench with the per-entry breakdown, old vs new likelihood path.
mkdir -p gpurun_out
echo "== 1. GPU suite"
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -s 2>&1 | grep -v "^$" | tail -120 > gpurun_out/a_pytest.log
tail -40 gpurun_out/a_pytest.log
echo "== 2. bench joint1024 (default) with breakdown"
timeout 600 python bench.py --steps 30 --breakdown > gpurun_out/a_bench_joint1024.json 2> gpurun_out/a_bench_joint1024.err
tail -c 600 gpurun_out/a_bench_joint1024.err
echo "== 3. old likelihood path, overlap off/on"
JD_LIK_BATCHED=0 JD_OVERLAP=0 timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-parity-check > gpurun_out/a_bench_joint1024_old.json 2>/dev/null
JD_OVERLAP=0 timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-parity-check --breakdown > gpurun_out/a_bench_joint1024_nooverlap.json 2>/dev/null
echo "== 4. cfg2"
timeout 300 python bench.py --workload cfg2 --steps 50 --no-cpu-baseline --no-gpu-baseline --breakdown > gpurun_out/a_bench_cfg2.json 2>/dev/null
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/a_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f.split("/")[-1], "value=%.1f ms/step=%.4f e2e=%s frac=%s" % (
            d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value"), r.get("frac")))
        print("   parity:", d.get("parity_check"))
        print("   cpu:", d.get("cpu_baseline"), "gpu:", d.get("gpu_baseline"))
        for k in (d.get("roofline_kernels") or [])[:6]:
            print("   ", k["kernel"], "us/step %.1f" % k["us_per_step"], "frac", k.get("frac"), k.get("bound"))
        if d.get("breakdown_us_per_step"):
            print("   ", d["breakdown_us_per_step"])
    except Exception as exc:
        print(f, "ERR", exc)
PY
