#!/bin/bash
# First GPU call of the next session: validate what was written after the round-1 GPU minutes ran out, then measure
# the configurations that have no number yet.  ~4 min on one B200.  Usage: gpurun --timeout 900 -- 'bash tools/next_gpu_run.sh'
mkdir -p gpurun_out
echo "== 1. full GPU suite with the pending tests enabled (shift kernels, fused shift engine path, shift runs)"
JD_TEST_PENDING=1 timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "== 1b. the same e2e parity tests with the likelihood / prior chains on two streams (JD_OVERLAP=1), then its speed"
JD_OVERLAP=1 timeout 300 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3
for v in 0 1; do JD_OVERLAP=$v timeout 200 python bench.py --steps 100 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg2 JD_OVERLAP=$v ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'])"; done
for v in 0 1; do JD_OVERLAP=$v timeout 200 python bench.py --workload joint1024 --steps 30 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('joint1024 JD_OVERLAP=$v ms/step', d['ms_per_step'])"; done
echo "== 2. default bench (headline) with the per-entry breakdown"
timeout 300 python bench.py --breakdown > gpurun_out/next_bench_cfg2.json 2> gpurun_out/next_bench_cfg2.err
echo "== 3. BASELINE configs[2] and [3] (cfg3: 8 x 512^2, 64^2 PSFs; cfg4: 20 x 1024^2, 201^2 PSFs on 1280^2 FFTs), one GPU"
timeout 300 python bench.py --workload cfg3 --steps 30 --no-cpu-baseline --breakdown > gpurun_out/next_bench_cfg3.json 2>/dev/null
timeout 400 python bench.py --workload cfg4 --steps 10 --no-cpu-baseline --breakdown > gpurun_out/next_bench_cfg4.json 2>/dev/null
JD_FFT_MIXED=0 timeout 400 python bench.py --workload cfg4 --steps 10 --no-cpu-baseline --no-e2e > gpurun_out/next_bench_cfg4_pow2.json 2>/dev/null
echo "== 3b. max-mode backward at 65 025 patches: Lam kernel (default there) vs triangular vs bucketed"
for v in "" "JD_BWD_BUCKETED=1"; do env $v timeout 200 python bench.py --workload joint1024 --steps 20 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('joint1024 $v ms/step', d['ms_per_step'])"; done
echo "== 4. split-FP16 prior kernel and the batched bootstrap runs"
timeout 200 python bench.py --steps 50 --no-cpu-baseline --backend 2 > gpurun_out/next_bench_cfg2_fp16.json 2>/dev/null
timeout 300 python bench.py --workload cfg5 --steps 20 > gpurun_out/next_bench_cfg5.json 2>/dev/null
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/next_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f.split("/")[-1], "value=%.1f ms/step=%.4f e2e=%s frac=%s" % (
            d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value"), r.get("frac")))
        if d.get("breakdown_us_per_step"):
            print("   ", d["breakdown_us_per_step"])
    except Exception as exc:
        print(f, "ERR", exc)
PY
