#!/bin/bash
# tools/profile.sh — the ncu passes of /opt/skills/guides/B200_PROFILING.md for this repo (run under gpurun, 1 GPU).
#   launch list   : every kernel of a short bench run with its device time (cold-cache, serialised: compare shares)
#   full captures : the keccak key-hash kernel and the trie leaf / branch kernels; the raw pages are exported to CSV
#                   on the box and the (large) .ncu-rep of the trie kernels is dropped to stay inside gpurun's 64 MiB
set -u
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 2 --warmup 3 --skip-cpu --skip-incremental > "$OUT/ncu_bench.log" 2>&1
ncu --set full --clock-control none --import-source on -k regex:keccak256_fixed32 -s 3 -c 1 -f -o "$OUT/prof_keccak32" \
    python bench.py --steps 1 --warmup 3 --skip-cpu --skip-state-root --skip-incremental > "$OUT/ncu_keccak.log" 2>&1
ncu --set full --clock-control none -k regex:"leaf_kernel|branch_kernel|branch_warp_kernel" -c 44 -f -o /tmp/prof_trie \
    python bench.py --steps 1 --warmup 3 --skip-cpu --keys 100000 --skip-incremental > "$OUT/ncu_trie.log" 2>&1
ncu -i /tmp/prof_trie.ncu-rep --page raw --csv > "$OUT/prof_trie_raw.csv" 2>/dev/null
ls -la "$OUT"
