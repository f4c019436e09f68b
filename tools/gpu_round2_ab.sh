#!/bin/bash
# round 2, GPU call AB (2 GPUs): 2-rank tests after the trace-reuse / row-fold changes, N = 2 bench, non-zero-mean mixture
mkdir -p gpurun_out
echo "== 1. GPU suite (incl. 2-rank tests)"
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider 2>&1 | grep -v "^$" | tail -30 > gpurun_out/ab_pytest.log
tail -4 gpurun_out/ab_pytest.log
echo "== 2. N = 2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/ab_n2.json 2> gpurun_out/ab_n2.err
tail -c 300 gpurun_out/ab_n2.err | grep -v "OMP_NUM\|\*\*\*" | tail -3
echo "== 3. N = 1: zero-mean vs non-zero-mean mixture; e2e repeats"
B="--steps 30 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-parity-check"
timeout 300 python bench.py $B > gpurun_out/ab_n1.json 2>/dev/null
timeout 300 python bench.py $B --gmm-mean-scale 0.01 > gpurun_out/ab_n1_nonzero_mean.json 2>/dev/null
timeout 300 python bench.py $B --gmm-mean-scale 0.01 --backend 5 > gpurun_out/ab_n1_nonzero_mean_b5.json 2>/dev/null
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/ab_n*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split("/")[-1], "value=%.1f ms/step=%.4f e2e=%s" % (d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value")))
    if d.get("parity_check"): print("   parity:", d["parity_check"].get("status"))
    for k in (d.get("roofline_kernels") or [])[:3]:
        print("   ", k["kernel"], "us/step %.1f" % k["us_per_step"], "frac", k.get("frac"), k.get("bound"))
PY
REPS=6 timeout 300 python tools/e2e_profile.py joint1024 100 2>&1 | grep "^run"
