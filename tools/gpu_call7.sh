#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
  echo "== trie tests (branch register path)"
  timeout 600 python -m pytest tests/test_gpu_trie.py tests/test_gpu_fullsize.py tests/test_gpu_items.py -m gpu -q 2>&1 | tail -3
  echo "== C3 phases"
  python tools/c3_phases.py --reps 5 2>&1 | tail -1
  B200_PHASE_TIMING=1 python tools/c3_phases.py --reps 2 2>&1 | tail -2 | head -1
  echo "== dtrie bench"
  timeout 500 python tools/dtrie_bench.py --base 100000000 --dirty 10000 --mix 80,10,10 --compare --cpu-sample 1000000 --blocks 8 2> gpurun_out/dtrie_bench.err | tail -1
  tail -12 gpurun_out/dtrie_bench.err
} > gpurun_out/gpu_call7.log 2>&1
tail -60 gpurun_out/gpu_call7.log
