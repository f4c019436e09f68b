set -x
JD_TC_TRIM8=1 timeout 200 python -m pytest tests/test_gpu_kernels.py -x -q -k "tensor_core or stream_k or gmm_prior_golden" 2>&1 | tail -4
JD_TC_TRIM8=1 timeout 120 python bench.py --steps 50 --no-cpu-baseline > gpurun_out/bench_cfg2_trim8.json 2>/dev/null
timeout 120 python bench.py --steps 50 --no-cpu-baseline > gpurun_out/bench_cfg2_e.json 2>/dev/null
timeout 120 python bench.py --steps 50 --no-cpu-baseline --backend 2 > gpurun_out/bench_cfg2_b2.json 2>/dev/null
timeout 120 python bench.py --workload joint1024 --steps 30 --no-cpu-baseline > gpurun_out/bench_joint_e.json 2>/dev/null
python - <<'PY'
import json
for f in ["bench_cfg2_trim8","bench_cfg2_e","bench_cfg2_b2","bench_joint_e"]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value=%.1f ms=%.4f kern_ms=%.4f frac=%.3f e2e=%.1f"%(d["value"],d["ms_per_step"],d["roofline"]["avg_launch_ms"],d["roofline"]["frac"],d["e2e"]["value"]))
    except Exception as e: print(f, "ERR", e)
PY
