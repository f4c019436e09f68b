#!/bin/bash
# round 2, GPU call AI (1 GPU): full GPU suite, smoke, default bench line of the final code
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider 2>&1 | grep -v "^$" | tail -15 > gpurun_out/ai_pytest.log
tail -3 gpurun_out/ai_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/ai_bench_default.json 2> gpurun_out/ai_bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/ai_bench_reference.json 2>/dev/null
python - <<'PY'
import json
d = json.loads(open("gpurun_out/ai_bench_default.json").read().strip().splitlines()[-1])
print("value=%.1f ms/step=%.4f e2e=%.1f frac=%.3f cpu=%.3f gpu=%.2f parity=%s launches=%s" % (
    d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["cpu_baseline"]["value"],
    d["gpu_baseline"]["value"], d["parity_check"]["status"], d["gpu_launches"]))
for k in d.get("roofline_kernels") or []:
    print("   ", k["kernel"], "us/step %.1f" % k["us_per_step"], "frac", round(k.get("frac") or 0, 3), k.get("bound"))
print(open("gpurun_out/ai_bench_reference.json").read()[:400])
PY
