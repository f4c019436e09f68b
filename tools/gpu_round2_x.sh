#!/bin/bash
# round 2, GPU call X: batched FFT likelihood path (parity + cfg3 / cfg4 / cfg2 timings)
mkdir -p gpurun_out
echo "== 1. GPU suite"
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider 2>&1 | grep -v "^$" | tail -40 > gpurun_out/x_pytest.log
tail -12 gpurun_out/x_pytest.log
echo "== 2. bench"
B="--steps 30 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-parity-check"
for wl in cfg3 cfg4 cfg2; do
  timeout 300 python bench.py $B --workload $wl > gpurun_out/x_bench_$wl.json 2>/dev/null
  JD_FFT_BATCHED=0 timeout 300 python bench.py $B --workload $wl > gpurun_out/x_bench_${wl}_unbatched.json 2>/dev/null
done
timeout 300 python bench.py $B > gpurun_out/x_bench_joint1024.json 2>/dev/null
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/x_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value=%.1f ms/step=%.4f" % (d["value"], d["ms_per_step"]))
        for k in (d.get("roofline_kernels") or [])[:6]:
            print("   ", k["kernel"], "us/step %.1f" % k["us_per_step"], "frac", k.get("frac"), k.get("bound"))
    except Exception as exc:
        print(f, "ERR", exc)
PY
