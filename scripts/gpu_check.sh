#!/bin/bash
# One gpurun call: GPU parity tests, bench lines, ncu launch list + one full capture of the top kernel.
# Usage (from the repo root, on the GPU box):  bash scripts/gpu_check.sh [tests|bench|ncu ...]
mkdir -p gpurun_out
WHAT="${@:-tests bench ncu}"
for w in $WHAT; do
case $w in
tests)
  timeout -s KILL 1200 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread > gpurun_out/tests_gpu.log 2>&1
  echo "tests exit $?"; tail -8 gpurun_out/tests_gpu.log ;;
small)
  timeout -s KILL 600 python bench.py --workload cifar-small --steps 2 --warmup 1 --cpu-batch 2 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err
  echo "bench small exit $?"; tail -c 1500 gpurun_out/bench_small.json; tail -5 gpurun_out/bench_small.err ;;
bench)
  timeout -s KILL 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench exit $?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err ;;
ref)
  timeout -s KILL 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
  echo "ref exit $?"; tail -c 1500 gpurun_out/bench_ref.json ;;
ncu)
  CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --probe-mode device"
  timeout -s KILL 900 $CMD > gpurun_out/ncu_plain.log 2>&1 &&
  timeout -s KILL 1500 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" \
      -c 15000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches exit $?"; wc -l gpurun_out/launches.csv
  timeout -s KILL 1500 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "timed/" \
      -k regex:k_gemm_tc3 -s 10 -c 4 \
      -o gpurun_out/prof_gemm_tc3 -f $CMD > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?"; ls -la gpurun_out/*.ncu-rep ;;
esac
done
