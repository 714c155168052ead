#!/bin/bash
# One GPU-box visit: all GPU tests, bench c3 (no CPU leg), ncu --set full of the kernels named in $NCU_KERNELS
set -u
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 2400 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== bench c3"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_c3_quick.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step']); print({k: round(v,3) for k,v in d['stages_ms'].items()}); print(d.get('cbca_ms_per_round_per_volume'))"
for k in ${NCU_KERNELS:-}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -o gpurun_out/r2b_$k -f python scripts/profile_step.py 2 > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log
done
