#!/bin/bash
# tools/profile.sh — the ncu passes of /opt/skills/guides/B200_PROFILING.md for this repo (run under gpurun, 1 GPU).
#   launch list   : every kernel of a short bench run with its device time (cold-cache, serialised: compare shares)
#   full captures : the keccak key-hash kernel and the trie leaf / branch kernels
set -u
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
BENCH="python bench.py --steps 2 --warmup 3 --skip-cpu"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/launches.csv" $BENCH > "$OUT/ncu_bench.log" 2>&1
ncu --set full --clock-control none --import-source on -k regex:keccak256_fixed32 -s 3 -c 1 -f -o "$OUT/prof_keccak32" \
    python bench.py --steps 1 --warmup 3 --skip-cpu --skip-state-root > "$OUT/ncu_keccak.log" 2>&1
ncu --set full --clock-control none --import-source on -k regex:leaf_kernel -c 2 -f -o "$OUT/prof_leaf" \
    python bench.py --steps 1 --warmup 3 --skip-cpu --keys 100000 > "$OUT/ncu_leaf.log" 2>&1
ncu --set full --clock-control none --import-source on -k regex:branch_kernel -c 14 -f -o "$OUT/prof_branch" \
    python bench.py --steps 1 --warmup 3 --skip-cpu --keys 100000 > "$OUT/ncu_branch.log" 2>&1
ls -la "$OUT"
