#!/bin/bash
# Session-4 gpurun call pieces.  Usage: bash scripts/gpu_s4.sh [update|benchab|tests ...]
mkdir -p gpurun_out
TAG="${TAG:-s4}"
for w in "$@"; do
case $w in
update)
  timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "broyden" --timeout 300 > gpurun_out/${TAG}_tests_broyden.log 2>&1
  echo "broyden tests exit $?"; tail -5 gpurun_out/${TAG}_tests_broyden.log
  timeout -s KILL 600 python scripts/update_bench.py > gpurun_out/${TAG}_update_bench.txt 2>&1
  echo "update bench exit $?"; cat gpurun_out/${TAG}_update_bench.txt ;;
benchab)
  for v in a b; do
    for smp in 0 1; do
      IMPFLOW_BENCH_NOSAMPLER=$smp timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_ab_${v}_nosampler$smp.json 2> gpurun_out/${TAG}_bench_ab_${v}_$smp.err
      echo "bench nosampler=$smp exit $?"; python - <<PY
import json
d = json.loads(open('gpurun_out/${TAG}_bench_ab_${v}_nosampler$smp.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'instrumented_pass', 'gpu_launches', 'clocks')}, d['e2e']['ms_per_step'])
PY
    done
  done ;;
persteps)
  for a in "--steps 10 --warmup 3" "--steps 20 --warmup 5" "--steps 20 --warmup 5"; do
    timeout -s KILL 600 python bench.py $a --no-cpu-baseline > gpurun_out/${TAG}_bench_ps.json 2> gpurun_out/${TAG}_bench_ps.err
    echo "bench $a exit $?"; python - <<PY
import json
d = json.loads(open('gpurun_out/${TAG}_bench_ps.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'per_step_ms')}, d['instrumented_pass']['ms_per_step'], d['e2e']['ms_per_step'])
PY
  done ;;
tests)
  timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread -x > gpurun_out/${TAG}_tests_gpu.log 2>&1
  echo "tests exit $?"; tail -5 gpurun_out/${TAG}_tests_gpu.log ;;
esac
done
