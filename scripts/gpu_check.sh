#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (+ reference arm), optional ncu.  Outputs in gpurun_out/.
#   bash scripts/gpu_check.sh [ncu] [full]
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench c3"
timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_c3.json
echo "== bench c2"
timeout 600 python bench.py --steps 10 --warmup 3 --workload c2 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_c2.json
echo "== bench reference arm"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -2 | tee gpurun_out/bench_ref.json
for a in "$@"; do
  if [ "$a" = "ncu" ]; then
    echo "== ncu launch list"
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python scripts/profile_step.py 2 > gpurun_out/ncu_list.log 2>&1
    tail -2 gpurun_out/ncu_list.log
  fi
  if [ "$a" = "full" ]; then
    echo "== ncu full captures"
    for k in k_cbca_pass k_sgm_pass k_cost_volume_tc k_conv64_tc k_wta; do
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 2 \
          -o gpurun_out/r1z_$k -f python scripts/profile_step.py 2 > gpurun_out/ncu_$k.log 2>&1
      tail -1 gpurun_out/ncu_$k.log
    done
  fi
done
