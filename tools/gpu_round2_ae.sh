#!/bin/bash
# round 2, GPU call AE: MAPDeconvolver.run fixed costs with pageable vs page-locked host arrays
mkdir -p gpurun_out
for p in 0 1; do
  echo "== PINNED=$p"
  PINNED=$p REPS=5 timeout 600 python tools/e2e_profile.py joint1024 100 2>&1 | head -40 | cut -c1-170
done
