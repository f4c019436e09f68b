#!/bin/bash
# round 2, GPU call AF: compute-sanitizer (memcheck, racecheck) over the kernels written or rewritten this round
mkdir -p gpurun_out
S="compute-sanitizer --launch-timeout 0 --print-limit 20"
PY="python -m pytest -x -q -m gpu -p no:cacheprovider"
run() {  # tag, tool, timeout, pytest args
  local tag=$1 tool=$2 to=$3; shift 3
  echo "== $tag ($tool)"
  timeout $to $S --tool $tool $PY "$@" > gpurun_out/af_$tag.log 2>&1
  echo "rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" gpurun_out/af_$tag.log | tail -4
}
run lik_mem memcheck 420 tests/test_gpu_likelihood.py -k "batched_likelihood_matches_oracle or joint_update"
run lik_race racecheck 420 tests/test_gpu_likelihood.py -k "joint_update or batched_likelihood_equals"
run bwd_mem memcheck 300 tests/test_gpu_kernels.py -k "bucketed or row_blocks"
run bwd_race racecheck 300 tests/test_gpu_kernels.py -k "bucketed"
run tcm2_mem memcheck 300 tests/test_gpu_kernels.py -k "test_gmm_prior_value_and_grad and (4- or 5-)"
