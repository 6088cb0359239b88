#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
  echo "== new GPU tests (stream, changesets) + whole suite"
  timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
  echo "== dtrie bench, instrumented"
  timeout 900 python tools/dtrie_bench.py --base 100000000 --dirty 10000 --mix 80,10,10 --compare --cpu-sample 1000000 --blocks 8 2> gpurun_out/dtrie_bench.err | tail -1
  tail -30 gpurun_out/dtrie_bench.err
  echo "== ncu --set full: node kernels of one C3 build"
  timeout 1200 ncu --set full --clock-control none -k regex:"leaf_storage_kernel|leaf_kernel|branch_kernel|branch_warp_kernel" -s 31 -c 31 -f -o /tmp/prof_trie \
      python tools/c3_phases.py --reps 2 > gpurun_out/ncu_trie.log 2>&1
  ncu -i /tmp/prof_trie.ncu-rep --page raw --csv > gpurun_out/r02_prof_trie_raw.csv 2>/dev/null
  ls -la /tmp/prof_trie.ncu-rep gpurun_out/r02_prof_trie_raw.csv
  tail -3 gpurun_out/ncu_trie.log
} > gpurun_out/gpu_call5.log 2>&1
tail -80 gpurun_out/gpu_call5.log
