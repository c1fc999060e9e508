#!/bin/bash
# bash scripts/gpu_scale.sh N : weak- and strong-scaling bench lines on N GPUs of one box (gpurun --gpus N)
N=$1
mkdir -p gpurun_out
for mode in weak strong; do
  timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
    bench.py --gpus $N --steps 20 --warmup 5 --scaling $mode --no-cpu-baseline > gpurun_out/r02_bench_${mode}_n$N.json 2> gpurun_out/r02_bench_${mode}_n$N.err
  echo "$mode n=$N exit $?"
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r02_bench_${mode}_n$N.json').read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'n_gpus', 'scaling')}, d['config']['per_gpu_batch'], d['config']['global_batch'], d['e2e']['value'])
except Exception as e:
    print('parse failed', e)
PY
  tail -2 gpurun_out/r02_bench_${mode}_n$N.err
done
