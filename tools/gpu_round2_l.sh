#!/bin/bash
# round 2, GPU call L/M: parity + timing + hand-over timeline of the two-tile kernels (backends 4, 5)
mkdir -p gpurun_out
echo "== 1. parity"
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -q -m gpu --tb=short -p no:cacheprovider -x \
   -k "two_tiles or value_and_grad or prior_golden or tensor_core or benchmark_size" 2>&1 | grep -v "^$" | tail -25 > gpurun_out/m_pytest.log
tail -8 gpurun_out/m_pytest.log
echo "== 2. prior forward alone"
timeout 120 python tools/tcm_exp.py 512 1024 2>&1 | grep "backend"
JD_TC_DEBUG=1 timeout 120 python tools/tcm_exp.py 1024 2>&1 | grep "backend"
echo "== 3. timeline"
export JD_LIB_PATH=$PWD/jolideco_b200/libjolideco_b200_trace.so
for b in 5 4; do
  timeout 200 python tools/tcm_trace.py 1024 $b 2>&1 | tail -22
done
