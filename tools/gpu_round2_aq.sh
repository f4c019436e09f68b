#!/bin/bash
# round 2, GPU call AQ (2 GPUs): end-to-end leg of a 2-rank run (symmetric buffers pooled across engines), with a profile
mkdir -p gpurun_out
JD_E2E_PROFILE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 100 --warmup 5 --no-parity-check > gpurun_out/aq_n2.json 2> gpurun_out/aq_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/aq_n2.json").read().strip().splitlines()[-1])
print("value=%.1f e2e=%s" % (d["value"], d.get("e2e")))
PY
grep -A 45 "Ordered by" gpurun_out/aq_n2.err | cut -c1-150 | head -50
