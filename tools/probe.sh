#!/bin/bash
# One gpurun call of measurements that steer the next change (phase timings, host-path walls).  Output: gpurun_out/probe/.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/probe; mkdir -p $O
B200_PHASE_TIMING=1 timeout 600 python tools/dstate_bench.py --blocks 8 > $O/dstate_host.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dt_ --csv --log-file $O/dstate_ncu.csv python tools/dstate_bench.py --blocks 3 > $O/dstate_ncu.log 2>&1
tail -n 1 $O/dstate_host.log | cut -c1-600
