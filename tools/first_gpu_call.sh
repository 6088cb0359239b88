#!/usr/bin/env bash
# The run that first put the dynamic trie / state / proofs on a B200 (profiles/r02_first_gpu_call.log): tests, racecheck /
# memcheck on them and their latency; everything under timeouts so a misbehaving kernel costs minutes, not the budget.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/first_gpu_call.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
  echo "== gated GPU tests"
  timeout 600 python -m pytest tests/test_gpu_dtrie.py tests/test_gpu_dstate.py tests/test_gpu_proofs.py tests/test_gpu_host_mirror.py -m gpu -q 2>&1 | tail -25
  echo "== the same on the large-block paths (multi-launch restructure, two-stage re-hash) forced"
  B200_DT_TWO_STAGE_MIN=0 B200_DT_FUSED_MAX=0 timeout 600 python -m pytest tests/test_gpu_dtrie.py tests/test_gpu_dstate.py -m gpu -q 2>&1 | tail -15
  echo "== C++ host mirror incl. DynamicTrie"
  timeout 300 python -m pytest tests/test_cpp_host.py -m gpu -q 2>&1 | tail -3
  echo "== compute-sanitizer racecheck / memcheck on small dynamic tests"
  timeout 500 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_dtrie.py -m gpu -q -x -k "n0-50 or 50-10-20" 2>&1 | tail -15
  timeout 500 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_dstate.py -m gpu -q -x -k "3-6-5" 2>&1 | tail -15
  echo "== dynamic trie latency vs merge+rebuild (C5 shape)"
  timeout 400 python tools/dtrie_bench.py --base 100000000 --dirty 10000 --mix 100,0,0 2>&1 | tail -2
  timeout 400 python tools/dtrie_bench.py --base 100000000 --dirty 10000 --mix 80,10,10 2>&1 | tail -2
  timeout 400 python tools/dtrie_bench.py --base 10000000 --dirty 10000 --mix 80,10,10 --compare 2>&1 | tail -2
  echo "== dynamic state (C3 shape, 2000 touched accounts per block)"
  timeout 600 python tools/dstate_bench.py --accounts 1000000 --slots 16 --touch 2000 --slot-writes 10 --device-resident 2>&1 | tail -2
  echo "== f2/f3/f4 throughput"
  timeout 300 python tools/ordered_bench.py --blocks 2000 --items 200 --shape receipts 2>&1 | tail -1
  timeout 300 python tools/ordered_bench.py --blocks 2000 --items 200 --shape transactions 2>&1 | tail -1
  timeout 300 python tools/rows_bench.py --accounts 1000000 --slots 16 2>&1 | tail -1
} > gpurun_out/first_gpu_call.log 2>&1
tail -80 gpurun_out/first_gpu_call.log
