#!/bin/bash
# compute-sanitizer evidence for the dependency-driven kernels (run on a GPU
# box through gpurun): memcheck (out-of-bounds / misaligned accesses), racecheck
# (shared-memory hazards inside a block) and synccheck, each over smoke() and a
# cart-pole N=300 solve. Logs land in gpurun_out/ and are copied to profiles/.
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
for tool in memcheck racecheck synccheck; do
  iters=25
  [ "$tool" = racecheck ] && iters=8
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 \
      python scripts/sanitize_target.py $iters > $OUT/${TAG}_sanitizer_${tool}.log 2>&1
  echo "[$tool] exit $?" >> $OUT/${TAG}_sanitizer_${tool}.log
  tail -4 $OUT/${TAG}_sanitizer_${tool}.log
done
