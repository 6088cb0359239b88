#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
  echo "== C3 phases"
  python tools/c3_phases.py 2>&1 | tail -2
  B200_PHASE_TIMING=1 python tools/c3_phases.py --reps 3 2>&1 | tail -3
  echo "== GPU tests touching the leaf kernel"
  timeout 900 python -m pytest tests/test_gpu_trie.py tests/test_gpu_fullsize.py tests/test_gpu_host_mirror.py tests/test_gpu_dstate.py -m gpu -q 2>&1 | tail -4
  echo "== bench (defaults, dynamic legs on)"
  timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_call4.json 2> gpurun_out/bench_call4.err; echo "rc=$?"
  tail -c 1500 gpurun_out/bench_call4.err
  python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_call4.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e'])
    print('state_root', d['state_root']['value'], d['state_root']['ms_per_step'], d['state_root']['roofline']['alu_frac'], d['state_root'].get('e2e'))
    print('incremental', d['incremental']['value'], d['incremental']['device_us'])
    print(json.dumps(d.get('dynamic'), indent=1))
except Exception as e:
    print('parse failed', e)
PY
} > gpurun_out/gpu_call4.log 2>&1
tail -120 gpurun_out/gpu_call4.log
