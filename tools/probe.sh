#!/bin/bash
# One gpurun call of measurements that steer the next change (phase timings, host-path walls).  Output: gpurun_out/probe/.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/probe; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_keccak.py -x -q -m gpu > $O/pytest_keccak.log 2>&1; echo "pytest rc=$?" >> $O/pytest_keccak.log
B200_PHASE_TIMING=1 B200_OVERLAP_STRUCTURE=0 timeout 300 python tools/c3_phases.py > $O/c3_serial.log 2>&1
B200_PHASE_TIMING=1 timeout 300 python tools/c3_phases.py > $O/c3_overlap.log 2>&1
B200_PHASE_TIMING=1 timeout 600 python tools/dstate_bench.py --blocks 4 > $O/dstate_host.log 2>&1
B200_PHASE_TIMING=1 timeout 600 python tools/dstate_bench.py --blocks 4 --device-resident > $O/dstate_dev.log 2>&1
timeout 600 python tools/hash_sort_bench.py > $O/hash_sort.json 2> $O/hash_sort.err
timeout 600 python tools/ordered_bench.py > $O/ordered.json 2> $O/ordered.err
timeout 900 python tools/rows_bench.py > $O/rows.json 2> $O/rows.err
tail -n 3 $O/pytest_keccak.log; cat $O/hash_sort.json $O/ordered.json $O/rows.json
