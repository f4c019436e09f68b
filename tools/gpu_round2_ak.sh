#!/bin/bash
# round 2, GPU call AK (8 GPUs): strong scaling of the final code 8 / 4 / 2 / 1 ranks (the smaller jobs side by side on
# disjoint GPUs), and at 8 ranks the single-dataset overlap experiment (JD_OVERLAP=2: prior forward on part of the SMs,
# enqueued first, likelihood chain beside it)
mkdir -p gpurun_out
run() {  # name, nproc, extra args...
  name=$1; n=$2; shift 2
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 "$@" > gpurun_out/ak_$name.json 2> gpurun_out/ak_$name.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus $n --steps 30 --warmup 5 "$@" > gpurun_out/ak_$name.json 2> gpurun_out/ak_$name.err
  fi
  tail -c 300 gpurun_out/ak_$name.err | grep -v "OMP_NUM\|\*\*\*" | tail -3
}
run n8 8 --breakdown --no-e2e
JD_OVERLAP=2 JD_TCM_CLUSTERS=37 run n8_split37 8 --no-e2e --no-parity-check
JD_OVERLAP=2 JD_TCM_CLUSTERS=52 run n8_split52 8 --no-e2e --no-parity-check
CUDA_VISIBLE_DEVICES=0,1,2,3 run n4 4 --no-e2e &
CUDA_VISIBLE_DEVICES=4,5 run n2 2 --no-e2e &
CUDA_VISIBLE_DEVICES=6 run n1 1 --no-cpu-baseline --no-gpu-baseline --no-parity-check --no-e2e &
wait
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/ak_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value=%.1f ms/step=%.4f" % (d["value"], d["ms_per_step"]))
        if d.get("parity_check"): print("   parity:", d["parity_check"].get("status"), d["parity_check"].get("n_rank_vs_1_rank_gradient_max_rel_err"))
        if d.get("peer_kernel_us_per_rank_last_step"): print("   peer kernel:", d["peer_kernel_us_per_rank_last_step"])
        for k in (d.get("roofline_kernels") or [])[:8]:
            print("   ", k["kernel"], "us/step %.1f" % k["us_per_step"])
    except Exception as exc:
        print(f, "ERR", exc)
PY
