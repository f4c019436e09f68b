set -x
( time timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 ) 2>&1 | tail -9
timeout 120 python tools/e2e_profile.py cfg2 20 > gpurun_out/e2e_profile.txt 2>&1; head -75 gpurun_out/e2e_profile.txt | cut -c1-150
