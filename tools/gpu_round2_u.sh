#!/bin/bash
# round 2, GPU call U (8 GPUs): multi-rank tests, strong scaling of the north-star joint deconvolution 1 -> 2 -> 4 -> 8
mkdir -p gpurun_out
echo "== 1. GPU suite (incl. the 2-rank tests)"
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider 2>&1 | grep -v "^$" | tail -30 > gpurun_out/u_pytest.log
tail -4 gpurun_out/u_pytest.log
run() {  # name, nproc, extra args...
  name=$1; n=$2; shift 2
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 "$@" > gpurun_out/u_$name.json 2> gpurun_out/u_$name.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus $n --steps 30 --warmup 5 "$@" > gpurun_out/u_$name.json 2> gpurun_out/u_$name.err
  fi
  tail -c 300 gpurun_out/u_$name.err | grep -v "OMP_NUM\|\*\*\*" | tail -3
}
echo "== 2. scaling"
run n8 8 --breakdown --no-e2e
run n1 1 --no-cpu-baseline --no-gpu-baseline --no-parity-check --no-e2e
run n2 2 --no-e2e
run n4 4 --no-e2e
run n8_nccl 8 --no-e2e --collective nccl --no-parity-check
run n8_b5 8 --no-e2e --backend 5 --no-parity-check
run n8_cfg3 8 --workload cfg3 --no-e2e
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/u_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f.split("/")[-1], "value=%.1f ms/step=%.4f frac=%s" % (d["value"], d["ms_per_step"], r.get("frac")))
        if d.get("parity_check"): print("   parity:", d["parity_check"].get("status"), d["parity_check"].get("n_rank_vs_1_rank_gradient_max_rel_err"))
        if d.get("peer_kernel_us_per_rank_last_step"): print("   peer kernel:", d["peer_kernel_us_per_rank_last_step"])
        for k in (d.get("roofline_kernels") or [])[:8]:
            print("   ", k["kernel"], "us/step %.1f" % k["us_per_step"], "frac", k.get("frac"), k.get("bound"))
    except Exception as exc:
        print(f, "ERR", exc)
PY
