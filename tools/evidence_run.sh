#!/usr/bin/env bash
# round-2 evidence run: the whole GPU suite, the default bench line, the ncu launch list of the same command and — with
# FULL_CAPTURES=1 — the --set full captures of the key-hash kernel and the node kernels of one C3 build
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
  echo "== whole GPU suite"
  timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6
  echo "== C3 phases"
  python tools/c3_phases.py --reps 5 2>&1 | tail -1
  B200_PHASE_TIMING=1 python tools/c3_phases.py --reps 2 2>&1 | tail -2 | head -1
  echo "== bench (defaults)"
  timeout 1500 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "rc=$?"
  tail -c 600 gpurun_out/r02_bench_1gpu.err
  echo "== bench --impl reference"
  timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null; echo "rc=$?"
  echo "== ncu launch list of a short bench run"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches.csv \
      python bench.py --steps 2 --warmup 3 --skip-cpu --skip-incremental --skip-dynamic > gpurun_out/ncu_bench.log 2>&1
  if [ "${FULL_CAPTURES:-0}" = 1 ]; then
  echo "== ncu --set full: key-hash kernel"
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:keccak256_fixed32 -s 3 -c 1 -f -o gpurun_out/r02_prof_keccak32 \
        python bench.py --steps 1 --warmup 3 --skip-cpu --skip-state-root --skip-incremental --skip-dynamic > gpurun_out/ncu_keccak.log 2>&1
    ncu -i gpurun_out/r02_prof_keccak32.ncu-rep --page raw --csv > gpurun_out/r02_prof_keccak32_raw.csv 2>/dev/null
    echo "== ncu --set full: node kernels of one C3 build"
    timeout 900 ncu --set full --clock-control none -k regex:"leaf_storage_kernel|leaf_kernel|branch_kernel|branch_warp_kernel|branch3_pipelined" -s 31 -c 31 -f -o /tmp/prof_trie \
        python tools/c3_phases.py --reps 2 > gpurun_out/ncu_trie.log 2>&1
    ncu -i /tmp/prof_trie.ncu-rep --page raw --csv > gpurun_out/r02_prof_trie_raw.csv 2>/dev/null
  fi
  ls -la gpurun_out/ | tail -12
} > gpurun_out/gpu_call_final.log 2>&1
tail -60 gpurun_out/gpu_call_final.log
