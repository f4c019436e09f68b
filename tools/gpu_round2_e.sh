#!/bin/bash
# round 2, GPU call E (2 GPUs): whole GPU suite incl. the 2-rank tests (NCCL and the self-synchronising peer kernel),
# then the north-star bench at N = 1 and N = 2 with the pre-timing self-check
mkdir -p gpurun_out
echo "== 1. GPU suite (2 GPUs visible)"
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider 2>&1 | grep -v "^$" | tail -60 > gpurun_out/e_pytest.log
tail -40 gpurun_out/e_pytest.log
echo "== 2. bench N=1"
timeout 900 python bench.py --steps 30 --breakdown > gpurun_out/e_bench_n1.json 2> gpurun_out/e_bench_n1.err
tail -c 300 gpurun_out/e_bench_n1.err
echo "== 3. bench N=2 (peer, then nccl)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/e_bench_n2.json 2> gpurun_out/e_bench_n2.err
tail -c 600 gpurun_out/e_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 --collective nccl --no-e2e > gpurun_out/e_bench_n2_nccl.json 2> gpurun_out/e_bench_n2_nccl.err
tail -c 300 gpurun_out/e_bench_n2_nccl.err
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/e_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f.split("/")[-1], "value=%.1f ms/step=%.4f e2e=%s frac=%s" % (
            d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value"), r.get("frac")))
        print("   parity:", d.get("parity_check"))
        print("   cpu:", d.get("cpu_baseline"), "gpu:", d.get("gpu_baseline"))
        print("   config:", d["config"].get("parallelism"))
        for k in (d.get("roofline_kernels") or [])[:7]:
            print("   ", k["kernel"], "us/step %.1f" % k["us_per_step"], "frac", k.get("frac"), k.get("bound"))
        if d.get("breakdown_us_per_step"):
            print("   ", d["breakdown_us_per_step"])
    except Exception as exc:
        print(f, "ERR", exc)
PY
