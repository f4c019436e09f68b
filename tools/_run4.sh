set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 300 gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
timeout 200 python bench.py --impl reference --steps 4 --warmup 1 2>/dev/null | cut -c1-700
JD_BWD_TRI=0 timeout 120 python bench.py --steps 50 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('BWD_TRI=0 ms', d['ms_per_step'])"
timeout 200 python bench.py --workload joint1024 --steps 30 > gpurun_out/bench_joint_f.json 2>/dev/null; cut -c1-400 gpurun_out/bench_joint_f.json
