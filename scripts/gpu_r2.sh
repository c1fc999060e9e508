#!/bin/bash
# Round-2 gpurun call: GPU parity tests, bench lines of every BASELINE.json configuration, whole-step ncu launch list.
# Usage (repo root, GPU box):  bash scripts/gpu_r2.sh [tests|bench|wl|ref|ncu|strong ...]   TAG=<suffix for the logs>
mkdir -p gpurun_out
TAG="${TAG:-r2}"
WHAT="${@:-tests bench wl ncu}"
for w in $WHAT; do
case $w in
tests)
  timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread -s > gpurun_out/${TAG}_tests_gpu.log 2>&1
  echo "tests exit $?"; grep -E "passed|failed|error" gpurun_out/${TAG}_tests_gpu.log | tail -5; grep -E "nstep" gpurun_out/${TAG}_tests_gpu.log | head -40 ;;
bench)
  timeout -s KILL 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_cifar.json 2> gpurun_out/${TAG}_bench_cifar.err
  echo "bench exit $?"; python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${TAG}_bench_cifar.json').read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'e2e', 'gpu_launches', 'broyden_solves_per_sec', 'solver_phase', 'solver_iterations_fwd_last_step', 'solver_iterations_bwd_last_step', 'cpu_baseline')})
    for k in ('roofline', 'roofline_secondary', 'roofline_tertiary'):
        r = d.get(k)
        if r: print(k, r['kernel'][:24], 'achieved %.1f frac %.3f of-ceiling %.3f share %.3f' % (r['achieved'], r['frac'], r['frac_of_3xtf32_ceiling'], r['share_of_step']))
    print([ (r['shape'], round(r['frac'],3), round(r['us_per_iteration'],1)) for r in d.get('roofline_solver', [])])
except Exception as e:
    print('parse failed', e)
PY
  tail -3 gpurun_out/${TAG}_bench_cifar.err ;;
wl)
  for wl in toy tabular-power tabular-miniboone tabular-bsds300 classifier; do
    timeout -s KILL 900 python bench.py --workload $wl --steps 5 --warmup 2 > gpurun_out/${TAG}_bench_$wl.json 2> gpurun_out/${TAG}_bench_$wl.err
    echo "bench $wl exit $?"; python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${TAG}_bench_$wl.json').read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'gpu_launches', 'broyden_solves_per_sec', 'solver_phase', 'solver_iterations_fwd_last_step', 'solver_iterations_bwd_last_step', 'cpu_baseline')})
except Exception as e:
    print('parse failed', e)
PY
    tail -3 gpurun_out/${TAG}_bench_$wl.err
  done ;;
ref)
  timeout -s KILL 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
  echo "ref exit $?"; tail -c 1500 gpurun_out/${TAG}_bench_ref.json ;;
ncu)
  # whole step, every thread (the autograd thread launches the implicit-backward solves): cudaProfilerStart/Stop
  # around the timed step instead of a thread-local NVTX filter
  CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
  IMPFLOW_PROFILER_API=1 timeout -s KILL 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
      -c 20000 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
  echo "ncu launches exit $?"; wc -l gpurun_out/${TAG}_launches.csv
  python scripts/summarise_launches.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launch_summary.txt 2>&1; head -40 gpurun_out/${TAG}_launch_summary.txt ;;
ncufull)
  # one --set full capture per tile kernel on its micro-benchmark (each after the same command exited 0 without ncu)
  for k in chain23 branch3; do
    CMD="python scripts/${k}_bench.py"
    timeout -s KILL 300 $CMD > gpurun_out/${TAG}_${k}_bench.txt 2>&1 &&
    timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:k_${k} -s 4 -c 3 \
        -o gpurun_out/${TAG}_prof_${k} -f $CMD > gpurun_out/${TAG}_ncu_full_${k}.log 2>&1
    echo "ncu full $k exit $?"; cat gpurun_out/${TAG}_${k}_bench.txt
  done ;;
insitu)
  timeout -s KILL 900 python scripts/kernel_time.py > gpurun_out/${TAG}_insitu.txt 2>&1; head -45 gpurun_out/${TAG}_insitu.txt ;;
strong)
  for n in 2 4; do
    timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --steps 10 --warmup 3 --scaling strong --no-cpu-baseline > gpurun_out/${TAG}_bench_strong_n$n.json 2> gpurun_out/${TAG}_bench_strong_n$n.err
    echo "strong n=$n exit $?"; tail -c 600 gpurun_out/${TAG}_bench_strong_n$n.json
  done ;;
esac
done
