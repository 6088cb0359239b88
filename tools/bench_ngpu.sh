#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
{
  nvidia-smi -L | head -8
  NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "rc=$?"
  python - <<PY
import json
try:
    lines=[l for l in open('gpurun_out/bench_${N}gpu.json').read().strip().splitlines() if l.startswith('{')]
    d=json.loads(lines[-1])
    print(len(lines), {k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], d['e2e'].get('numa_node_bound'))
    print('state_root', d['state_root']['value'], d['state_root']['ms_per_step'])
    print('c4', d['mainnet_shape']['value'], d['mainnet_shape']['ms_per_step'], d['mainnet_shape']['config']['workload'][:80])
    hp=d.get('hash_partition'); print('hash_partition', hp['value'], hp['ms_per_step'], hp['sorted_and_owned_rank0'])
except Exception as e:
    print('parse failed', e)
PY
  tail -5 gpurun_out/bench_${N}gpu.err | cut -c1-300
} > gpurun_out/gpu_call_${N}gpu.log 2>&1
tail -30 gpurun_out/gpu_call_${N}gpu.log
