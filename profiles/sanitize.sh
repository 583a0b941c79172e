#!/usr/bin/env bash
# compute-sanitizer over the hot path (the reference has no race / memory checking at all).
# Runs on a B200 through gpurun:   gpurun --timeout 1500 -- 'bash profiles/sanitize.sh'
# Each tool drives profiles/sanitize_driver.py: a small HER buffer (ragged episodes, FIFO eviction), the stand-alone
# sampler, DDPG / TD3 updates through the row-slab kernels (in-kernel sampling) and through the tiled engines
# (batch 1100 fp32, 2048 tcgen05), the normaliser, prioritised replay -- graphs off, so every launch is checked.
set -u
mkdir -p gpurun_out
export GCRL_B200_NO_GRAPH=1
for tool in memcheck racecheck synccheck initcheck; do
  echo "== compute-sanitizer --tool $tool"
  compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python profiles/sanitize_driver.py > gpurun_out/sanitize_$tool.log 2>&1
  rc=$?
  tail -n 4 gpurun_out/sanitize_$tool.log
  echo "   exit code $rc"
done
