#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (+ reference arm), launch list.  Outputs in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "== pytest gpu" 
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -150 | tee gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench c3"
timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_c3.json
echo "== bench c2"
timeout 600 python bench.py --steps 10 --warmup 3 --workload c2 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_c2.json
if [ "${1:-}" = "ncu" ]; then
  echo "== ncu launch list"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  tail -2 gpurun_out/ncu_bench.log
fi
