set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --breakdown > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 300 gpurun_out/bench_default.err; cut -c1-300 gpurun_out/bench_default.json
timeout 300 python bench.py --workload joint1024 --steps 30 --breakdown > gpurun_out/bench_joint1024_n1.json 2>/dev/null; python -c "
import json
for f in ['bench_default','bench_joint1024_n1']:
    d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline'], d['breakdown_us_per_step'])
"
timeout 200 python bench.py --steps 50 --no-cpu-baseline --backend 2 > gpurun_out/bench_cfg2_backend2.json 2>/dev/null; python -c "
import json
d=json.loads(open('gpurun_out/bench_cfg2_backend2.json').read().strip().splitlines()[-1]); print('backend2', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['avg_launch_ms'])"
