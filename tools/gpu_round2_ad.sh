#!/bin/bash
# round 2, GPU call AD: final ncu evidence (launch lists + --set full) and bench lines of the final code
mkdir -p gpurun_out
B="--no-cpu-baseline --no-gpu-baseline --no-e2e --no-parity-check --no-graph"
echo "== 1. launch lists"
for wl in joint1024 cfg2 cfg3; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/ad_launches_$wl.csv \
    python bench.py --workload $wl --steps 2 --warmup 3 $B > /dev/null 2>&1
done
echo "== 2. --set full"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"gmm_fwd_tcx2|lik_kernel|gmm_bwd_bucket|joint_grad|bwd_hist|bwd_scan|bwd_scatter|step_begin_flux" -s 16 -c 8 -f -o gpurun_out/prof_joint1024_r02_final2 \
    python bench.py --steps 2 --warmup 3 $B > gpurun_out/ad_ncu_full.log 2>&1
tail -1 gpurun_out/ad_ncu_full.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"rows_fwd|cols_kernel|rows_inv" -s 60 -c 6 -f -o gpurun_out/prof_cfg3_fft_r02_final \
    python bench.py --workload cfg3 --steps 2 --warmup 3 $B > gpurun_out/ad_ncu_full_cfg3.log 2>&1
tail -1 gpurun_out/ad_ncu_full_cfg3.log
echo "== 3. bench lines"
timeout 900 python bench.py > gpurun_out/ad_bench_default.json 2> gpurun_out/ad_bench_default.err
timeout 600 python bench.py --workload cfg2 --steps 100 > gpurun_out/ad_bench_cfg2.json 2>/dev/null
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/ad_bench_*.json")):
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
