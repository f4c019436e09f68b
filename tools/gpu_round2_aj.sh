#!/bin/bash
# round 2, GPU call AJ (2 GPUs): the 2-rank GPU tests of the final code
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -k "rank or peer or joint or dist or shard" 2>&1 | grep -v "^$" | tail -8 > gpurun_out/aj_pytest.log
tail -3 gpurun_out/aj_pytest.log
