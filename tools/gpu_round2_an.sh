#!/bin/bash
# round 2, GPU call AN (8 GPUs): 8-rank north-star run with the auto-tuned split of one-dataset steps
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 8 --steps 30 --warmup 5 --breakdown --no-e2e > gpurun_out/an_n8.json 2> gpurun_out/an_n8.err
tail -c 300 gpurun_out/an_n8.err | grep -v "OMP_NUM\|\*\*\*" | tail -3
python - <<'PY'
import json
d = json.loads(open("gpurun_out/an_n8.json").read().strip().splitlines()[-1])
c = d["config"]
print("value=%.1f ms/step=%.4f pairs=%s tuning=%s" % (d["value"], d["ms_per_step"], c.get("prior_forward_sm_pairs"), c.get("split_tuning_ms")))
print("parity:", d["parity_check"])
print("peer:", d.get("peer_kernel_us_per_rank_last_step"))
for k in (d.get("roofline_kernels") or [])[:8]:
    print("   ", k["kernel"], "us/step %.1f" % k["us_per_step"])
PY
