#!/usr/bin/env bash
# Round-2 evidence: launch lists and one `--set full` capture of the dominant kernel of every part that was only
# emulation-validated in round 1 (dynamic state, ordered roots, device table rows).  Run AFTER tools/first_gpu_call.sh
# is green.  One GPU, short commands (ncu replays every kernel ~40 times).
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/ncu_round2.sh'
# Afterwards, here: ncu -i gpurun_out/<name>.ncu-rep --page raw --csv | grep -E 'dram__bytes_(read|write)\.sum|gpu__time_duration|sm__inst_executed_pipe_alu|launch__registers_per_thread'
# and summarise under profiles/r02_*.md.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export B200_DTRIE_ON_GPU=1
NCU_LIST="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
NCU_FULL="ncu --set full --clock-control none --import-source on"
{
  echo "== dynamic state: launch list of a few blocks, then the wavefront kernel"
  timeout 600 $NCU_LIST -c 600 --log-file gpurun_out/r02_dstate_launches.csv \
      python tools/dstate_bench.py --accounts 200000 --slots 16 --touch 2000 --slot-writes 10 --blocks 3 2>&1 | tail -1
  timeout 600 $NCU_FULL -k regex:dt_wavefront -s 4 -c 3 -o gpurun_out/r02_dt_wavefront \
      python tools/dstate_bench.py --accounts 200000 --slots 16 --touch 2000 --slot-writes 10 --blocks 3 2>&1 | tail -1
  timeout 600 $NCU_FULL -k regex:dt_restructure_fused -s 2 -c 2 -o gpurun_out/r02_dt_fused \
      python tools/dstate_bench.py --accounts 200000 --slots 16 --touch 2000 --slot-writes 10 --blocks 3 2>&1 | tail -1
  echo "== ordered roots: launch list, then the leaf kernel (receipt-shaped batch)"
  timeout 600 $NCU_LIST -c 400 --log-file gpurun_out/r02_ordered_launches.csv \
      python tools/ordered_bench.py --blocks 2000 --items 200 --shape receipts --reps 1 2>&1 | tail -1
  timeout 600 $NCU_FULL -k regex:ordered_leaf -s 1 -c 1 -o gpurun_out/r02_ordered_leaf \
      python tools/ordered_bench.py --blocks 2000 --items 200 --shape receipts --reps 1 2>&1 | tail -1
  echo "== device table rows: launch list, then the row encoder"
  timeout 600 $NCU_LIST -c 400 --log-file gpurun_out/r02_rows_launches.csv \
      python tools/rows_bench.py --accounts 200000 --slots 16 --reps 1 2>&1 | tail -1
  timeout 600 $NCU_FULL -k regex:encode_rows -s 1 -c 1 -o gpurun_out/r02_encode_rows \
      python tools/rows_bench.py --accounts 200000 --slots 16 --reps 1 2>&1 | tail -1
} > gpurun_out/ncu_round2.log 2>&1
tail -30 gpurun_out/ncu_round2.log
