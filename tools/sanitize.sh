#!/bin/bash
# compute-sanitizer evidence for the kernels that use ballots, per-warp atomics, shared-memory staging and cooperative
# grid syncs.  Run on the GPU box:  bash tools/sanitize.sh  -> gpurun_out/sanitize_*.txt
# (summaries are copied to profiles/ by hand, per round).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  for tgt in smoke two_level sort build; do
    out=gpurun_out/sanitize_${tool}_${tgt}.txt
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_targets.py $tgt > $out 2>&1
    echo "rc=$?" >> $out
    echo "== $tool $tgt: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|rc=' $out | tr '\n' ' ')"
  done
done
