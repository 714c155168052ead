#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (+ reference arm).  Outputs in gpurun_out/.
#   bash scripts/gpu_check.sh [ref]
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "== pytest gpu"
timeout 2400 python -m pytest tests -m gpu -q --tb=short -x -s 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench c3"
timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_c3.json | cut -c1-600
echo "== bench c2"
timeout 600 python bench.py --steps 10 --warmup 3 --workload c2 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_c2.json | cut -c1-600
for a in "$@"; do
  if [ "$a" = "ref" ]; then
    echo "== bench reference arm"
    MCCNN_REF_BUDGET_S=${MCCNN_REF_BUDGET_S:-60} timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.json | cut -c1-1500
  fi
done
