#!/usr/bin/env bash
# gpurun --gpus 2: the exchange steps behind the C ABI on two GPUs, then the bench under torchrun at N=2
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
  nvidia-smi -L
  echo "== b200_comm tests (2 ranks, one thread per GPU)"
  timeout 600 python -m pytest tests/test_gpu_comm.py -m gpu -q 2>&1 | tail -8
  echo "== bench at N=2 (torchrun)"
  NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "rc=$?"
  grep -c "NCCL INFO" gpurun_out/bench_2gpu.err; grep -m2 "nranks" gpurun_out/bench_2gpu.err | cut -c1-200
  wc -l gpurun_out/bench_2gpu.json
  python - <<'PY'
import json
try:
    lines=[l for l in open('gpurun_out/bench_2gpu.json').read().strip().splitlines() if l.startswith('{')]
    d=json.loads(lines[-1])
    print(len(lines), {k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], d['e2e'].get('numa_node_bound'))
    print('state_root', d['state_root']['value'], d['state_root']['ms_per_step'], d['state_root']['root'][:16])
    print('c4', d['mainnet_shape']['value'] if d.get('mainnet_shape') else None)
    print('hash_partition', d.get('hash_partition'))
except Exception as e:
    print('parse failed', e)
PY
  tail -5 gpurun_out/bench_2gpu.err | cut -c1-300
} > gpurun_out/gpu_call_2gpu.log 2>&1
tail -60 gpurun_out/gpu_call_2gpu.log
