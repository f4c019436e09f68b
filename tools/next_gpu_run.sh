#!/bin/bash
# First GPU calls of the next session: what was written after the round-2 GPU minutes ran out has CPU tests only
# (the collective split decision and the pooled peer buffers ran on 2 ranks, not on 8).
#   gpurun --timeout 900 -- 'bash tools/next_gpu_run.sh one'            (~3 min on one B200)
#   gpurun --gpus 8 --timeout 600 -- 'bash tools/next_gpu_run.sh eight' (~2 min on 8 GPUs)
mkdir -p gpurun_out
B="--no-cpu-baseline --no-gpu-baseline --no-e2e --no-parity-check --no-graph"
if [ "$1" = "eight" ]; then
  echo "== 8 ranks, default settings: split of the one-dataset steps decided by all ranks together, e2e with pooled peer buffers"
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 8 --steps 30 --warmup 5 --breakdown > gpurun_out/next_n8.json 2> gpurun_out/next_n8.err
  python - <<'PY'
import json
d = json.loads(open("gpurun_out/next_n8.json").read().strip().splitlines()[-1])
c = d["config"]
print("value=%.1f ms/step=%.4f e2e=%s pairs=%s tuning=%s" % (d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value"),
      c.get("prior_forward_sm_pairs"), c.get("split_tuning_ms")))
print("parity:", d.get("parity_check"))
print("peer:", d.get("peer_kernel_us_per_rank_last_step"))
PY
  exit 0
fi
echo "== 1. full GPU suite"
timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -4
echo "== 2. racecheck / memcheck of the second-generation bucket kernel (persistent, device work counter)"
S="compute-sanitizer --launch-timeout 0 --print-limit 20"
for tool in memcheck racecheck; do
  timeout 300 $S --tool $tool python -m pytest -x -q -m gpu -p no:cacheprovider tests/test_gpu_kernels.py -k "bucketed" > gpurun_out/next_bwd8_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/next_bwd8_$tool.log | tail -2
done
echo "== 3. launch list + --set full of the north-star step with the final kernels"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/next_launches_joint1024.csv \
    python bench.py --steps 2 --warmup 3 $B > /dev/null 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"gmm_fwd_tcx2|lik_kernel|gmm_bwd_bucket8|joint_grad|bwd_hist|bwd_scan|bwd_scatter|step_begin_flux" \
    -s 16 -c 8 -f -o gpurun_out/prof_joint1024_next python bench.py --steps 2 --warmup 3 $B > /dev/null 2>&1
echo "== 4. default bench line, cfg2 (split of one-dataset steps)"
timeout 900 python bench.py > gpurun_out/next_bench_default.json 2>/dev/null
timeout 600 python bench.py --workload cfg2 --steps 100 > gpurun_out/next_bench_cfg2.json 2>/dev/null
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/next_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f.split("/")[-1], "value=%.1f ms/step=%.4f e2e=%s frac=%s pairs=%s" % (
            d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value"), r.get("frac"), d["config"].get("prior_forward_sm_pairs")))
    except Exception as exc:
        print(f, "ERR", exc)
PY
