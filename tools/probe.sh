#!/bin/bash
# One gpurun call of measurements that steer the next change (phase timings, host-path walls).  Output: gpurun_out/probe/.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/probe; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_dtrie.py tests/test_gpu_dstate.py tests/test_gpu_zz_table_rows_device.py -x -q -m gpu > $O/pytest_dyn.log 2>&1; echo "pytest rc=$?" >> $O/pytest_dyn.log
B200_PHASE_TIMING=1 timeout 600 python tools/dstate_bench.py --blocks 8 --device-resident > $O/dstate.log 2>&1
timeout 900 python tools/rows_bench.py > $O/rows.json 2> $O/rows.err
tail -n 3 $O/pytest_dyn.log; tail -n 1 $O/dstate.log | cut -c1-700; cat $O/rows.json
