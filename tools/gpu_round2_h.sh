#!/bin/bash
# round 2, GPU call H (8 GPUs): strong scaling of the north-star joint deconvolution, 1 -> 2 -> 4 -> 8 ranks
mkdir -p gpurun_out
run() {  # name, nproc, extra args...
  name=$1; n=$2; shift 2
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 "$@" > gpurun_out/h_$name.json 2> gpurun_out/h_$name.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus $n --steps 30 --warmup 5 "$@" > gpurun_out/h_$name.json 2> gpurun_out/h_$name.err
  fi
  tail -c 300 gpurun_out/h_$name.err | grep -v "OMP_NUM\|\*\*\*" | tail -3
}
run n8 8 --breakdown
run n1 1 --no-cpu-baseline --no-gpu-baseline --no-parity-check --no-e2e
run n2 2 --no-e2e
run n4 4 --no-e2e
JD_OVERLAP=0 run n8_nooverlap 8 --no-e2e --no-parity-check
run n8_nccl 8 --no-e2e --collective nccl --no-parity-check
run n8_cfg3 8 --workload cfg3 --no-e2e
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/h_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f.split("/")[-1], "value=%.1f ms/step=%.4f e2e=%s frac=%s" % (
            d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value"), r.get("frac")))
        if d.get("parity_check"): print("   parity:", d["parity_check"])
        for k in (d.get("roofline_kernels") or [])[:8]:
            print("   ", k["kernel"], "us/step %.1f" % k["us_per_step"], "frac", k.get("frac"), k.get("bound"))
    except Exception as exc:
        print(f, "ERR", exc)
PY
