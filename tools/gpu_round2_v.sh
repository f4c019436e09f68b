#!/bin/bash
# round 2, GPU call V: ncu evidence of the final kernels (launch lists + --set full), default bench lines
mkdir -p gpurun_out
B="--no-cpu-baseline --no-gpu-baseline --no-e2e --no-parity-check --no-graph"
echo "== 1. launch lists"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/v_launches_joint1024.csv \
    python bench.py --steps 2 --warmup 3 $B > gpurun_out/v_ncu_bench_joint.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/v_launches_cfg2.csv \
    python bench.py --workload cfg2 --steps 3 --warmup 3 $B > gpurun_out/v_ncu_bench_cfg2.log 2>&1
echo "== 2. --set full of the step's kernels (joint1024)"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"gmm_fwd_tcx2|lik_kernel|gmm_bwd_bucket|joint_grad|bwd_hist|bwd_scan|bwd_scatter|step_begin_flux" -s 16 -c 8 -f -o gpurun_out/prof_joint1024_r02_final \
    python bench.py --steps 2 --warmup 3 $B > gpurun_out/v_ncu_full.log 2>&1
tail -2 gpurun_out/v_ncu_full.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"gmm_fwd_tcx2|gmm_bwd_max_tri" -s 4 -c 2 -f -o gpurun_out/prof_cfg2_r02_final \
    python bench.py --workload cfg2 --steps 2 --warmup 3 $B > gpurun_out/v_ncu_full_cfg2.log 2>&1
echo "== 3. bench lines (default arguments = what the driver runs; cfg2)"
timeout 900 python bench.py > gpurun_out/v_bench_default.json 2> gpurun_out/v_bench_default.err
timeout 300 python bench.py --workload cfg2 --steps 50 > gpurun_out/v_bench_cfg2.json 2>/dev/null
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/v_bench_reference.json 2>/dev/null
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/v_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f.split("/")[-1], "value=%.2f ms/step=%.4f e2e=%s frac=%s" % (
            d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value"), r.get("frac")))
        print("   cpu:", (d.get("cpu_baseline") or {}).get("value"), "gpu:", (d.get("gpu_baseline") or {}).get("value"), "parity:", (d.get("parity_check") or {}).get("status"))
        for k in (d.get("roofline_kernels") or [])[:8]:
            print("   ", k["kernel"], "us/step %.1f" % k["us_per_step"], "frac", k.get("frac"), k.get("bound"))
    except Exception as exc:
        print(f, "ERR", exc)
PY
ls -la gpurun_out/prof_*final* gpurun_out/v_launches*
