#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
  echo "== trie tests (pipelined class-0 kernel)"
  timeout 600 python -m pytest tests/test_gpu_trie.py tests/test_gpu_fullsize.py tests/test_gpu_items.py tests/test_gpu_dstate.py -m gpu -q 2>&1 | tail -3
  echo "== C3 phases"
  python tools/c3_phases.py --reps 5 2>&1 | tail -1
  B200_PHASE_TIMING=1 python tools/c3_phases.py --reps 2 2>&1 | tail -2 | head -1
  echo "== ncu: node kernels of one C3 build"
  timeout 900 ncu --set full --clock-control none -k regex:"leaf_storage_kernel|leaf_kernel|branch_kernel|branch_warp_kernel|branch3_pipelined" -s 31 -c 31 -f -o /tmp/prof_trie \
      python tools/c3_phases.py --reps 2 > gpurun_out/ncu_trie.log 2>&1
  ncu -i /tmp/prof_trie.ncu-rep --page raw --csv > gpurun_out/r02_prof_trie_raw_b.csv 2>/dev/null
  ls -la gpurun_out/r02_prof_trie_raw_b.csv
} > gpurun_out/gpu_call8.log 2>&1
tail -40 gpurun_out/gpu_call8.log
