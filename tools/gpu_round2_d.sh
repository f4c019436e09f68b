#!/bin/bash
# round 2, GPU call D: where does the backend-3 prior kernel spend its ~600 clk per component?  + 2-GPU tests
mkdir -p gpurun_out
for d in 0 1 2 4 8 9 3 16; do JD_TC_DEBUG=$d timeout 120 python tools/tcm_exp.py 512 1024 2>&1 | grep "backend 3"; done
for c in 37 74; do JD_TCM_CLUSTERS=$c timeout 120 python tools/tcm_exp.py 1024 2>&1 | grep "backend 3" | sed "s/^/clusters $c: /"; done
timeout 120 python tools/tcm_exp.py 512 1024 2>&1 | grep "backend 1"
