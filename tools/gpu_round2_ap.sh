#!/bin/bash
# round 2, GPU call AP (1 GPU): tests of the prior forward on part of the SM pairs (kernel level, goldens end to end)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py -q -m gpu --tb=short -p no:cacheprovider \
   -k "part_of_the_sm_pairs or beside_the_likelihood_chain" 2>&1 | grep -v "^$" | tail -25 > gpurun_out/ap_pytest.log
tail -25 gpurun_out/ap_pytest.log
