#!/bin/bash
# round 2, GPU call B: GPU suite again (capture fix), north-star bench, ncu captures of the new likelihood kernels
mkdir -p gpurun_out
echo "== 1. GPU suite"
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -s 2>&1 | grep -v "^$" | tail -150 > gpurun_out/b_pytest.log
tail -30 gpurun_out/b_pytest.log
echo "== 2. bench joint1024 (default) with breakdown"
timeout 900 python bench.py --steps 30 --breakdown > gpurun_out/b_bench_joint1024.json 2> gpurun_out/b_bench_joint1024.err
tail -c 400 gpurun_out/b_bench_joint1024.err
echo "== 3. likelihood kernels alone + ncu"
timeout 200 python tools/lik_one.py
timeout 600 ncu --set full --import-source on --clock-control none -k regex:lik_kernel -s 6 -c 2 -f -o gpurun_out/prof_lik_r02 python tools/lik_one.py > gpurun_out/b_ncu_lik.log 2>&1
tail -3 gpurun_out/b_ncu_lik.log
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/b_bench_*.json")):
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
