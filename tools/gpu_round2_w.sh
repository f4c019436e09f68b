#!/bin/bash
# round 2, GPU call W: suite + bench after the stride-4 fold gather / scan changes; per-kernel rooflines on one stream
mkdir -p gpurun_out
echo "== 1. GPU suite"
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider 2>&1 | grep -v "^$" | tail -30 > gpurun_out/w_pytest.log
tail -5 gpurun_out/w_pytest.log
echo "== 2. bench"
B="--steps 30 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-parity-check"
timeout 300 python bench.py $B > gpurun_out/w_bench_joint1024.json 2>/dev/null
timeout 300 python bench.py $B --workload cfg2 --steps 50 > gpurun_out/w_bench_cfg2.json 2>/dev/null
timeout 300 python bench.py $B --workload cfg3 > gpurun_out/w_bench_cfg3.json 2>/dev/null
timeout 300 python bench.py $B --workload cfg4 --steps 10 > gpurun_out/w_bench_cfg4.json 2>/dev/null
timeout 300 python bench.py --workload cfg5 --steps 20 > gpurun_out/w_bench_cfg5.json 2>/dev/null
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/w_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value=%.1f ms/step=%.4f" % (d["value"], d["ms_per_step"]))
        for k in (d.get("roofline_kernels") or [])[:9]:
            print("   ", k["kernel"], "us/step %.1f" % k["us_per_step"], "frac", k.get("frac"), k.get("bound"))
    except Exception as exc:
        print(f, "ERR", exc)
PY
echo "== 3. launch list cfg4 (one step)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/w_launches_cfg4.csv \
    python bench.py --workload cfg4 --steps 1 --warmup 3 $B --no-graph > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/w_launches_cfg4.csv 2>&1 | tail -16
