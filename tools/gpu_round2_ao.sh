#!/bin/bash
# round 2, GPU call AO (2 GPUs): collective split decision: 2-rank tests, 2 ranks x 1 dataset bench (split eligible), 2 ranks x 4 datasets
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | grep -v "^$" | tail -5 > gpurun_out/ao_pytest.log
tail -2 gpurun_out/ao_pytest.log
for v in "d2 --datasets 2" "d8"; do
  set -- $v; tag=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus 2 --steps 30 --warmup 5 --no-e2e "$@" > gpurun_out/ao_n2_$tag.json 2> gpurun_out/ao_n2_$tag.err
  tail -c 300 gpurun_out/ao_n2_$tag.err | grep -v "OMP_NUM\|\*\*\*" | tail -3
done
JD_SPLIT_CLUSTERS=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29477 \
    bench.py --gpus 2 --steps 30 --warmup 5 --no-e2e --no-parity-check --datasets 2 > gpurun_out/ao_n2_d2_nosplit.json 2>/dev/null
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/ao_n2_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        c = d["config"]
        print(f.split("/")[-1], "value=%.1f ms/step=%.4f pairs=%s tuning=%s" % (d["value"], d["ms_per_step"], c.get("prior_forward_sm_pairs"), c.get("split_tuning_ms")))
        if d.get("parity_check"): print("   parity:", d["parity_check"])
        print("   peer:", d.get("peer_kernel_us_per_rank_last_step"))
    except Exception as exc:
        print(f, "ERR", exc)
PY
