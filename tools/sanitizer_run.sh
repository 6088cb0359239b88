#!/usr/bin/env bash
# compute-sanitizer memcheck / racecheck over the round-2 kernels (register-path leaves and branches, pipelined class-0 kernel,
# items, changesets, stream) on small inputs
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
  echo "== memcheck: from-scratch builds (leaf_storage_kernel, branch3_pipelined_kernel, strip classes)"
  timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_trie.py -m gpu -q -x -k "golden or deep_shared or forest_random or rejects or sticky" 2>&1 | tail -6
  echo "== memcheck: items / changesets / stream"
  timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_items.py tests/test_gpu_changesets.py tests/test_gpu_stream.py -m gpu -q -x -k "not threshold_mirror and not 3000 and not 60000" 2>&1 | tail -6
  echo "== racecheck: shared-memory strips next to the register paths"
  timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_trie.py tests/test_gpu_items.py -m gpu -q -x -k "deep_shared or malformed or untouched" 2>&1 | tail -6
} > gpurun_out/r02_sanitizer.log 2>&1
tail -30 gpurun_out/r02_sanitizer.log
