#!/usr/bin/env bash
# full GPU suite (dynamic paths ungated) + the forced large-block variants + one default bench run
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
  echo "== full GPU suite"
  timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15
  echo "== dynamic paths, large-block forms forced"
  B200_DT_TWO_STAGE_MIN=0 B200_DT_FUSED_MAX=0 timeout 600 python -m pytest tests/test_gpu_dtrie.py tests/test_gpu_dstate.py -m gpu -q 2>&1 | tail -5
  echo "== racecheck on the split-run shape"
  timeout 500 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_dtrie.py -m gpu -q -x -k "split_runs" 2>&1 | tail -8
  echo "== bench (defaults)"
  timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_call2.json 2> gpurun_out/bench_call2.err; echo "rc=$?"
  tail -c 3000 gpurun_out/bench_call2.err
  python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_call2.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline'])
    print('state_root', d['state_root']['value'], d['state_root']['ms_per_step'], d['state_root'].get('roofline'))
    print('incremental', {k:v for k,v in d['incremental'].items() if k not in ('config',)})
except Exception as e:
    print('parse failed', e)
PY
} > gpurun_out/gpu_call2.log 2>&1
tail -70 gpurun_out/gpu_call2.log
