#!/bin/bash
# Round-2 evidence: ncu launch list of bench.py itself + ncu --set full captures of every kernel of the hot path.
set -u
mkdir -p gpurun_out
echo "== launch list (bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-flat-stage)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-flat-stage > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/ncu_list.log | cut -c1-200
for k in k_cbca_colrow_g k_cbca_close_g k_cbca_pass k_sgm_pass k_conv64_h; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 2 -o gpurun_out/r2f_$k -f python scripts/profile_step.py 2 > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log
done
for k in k_cost_volume_tc; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/r2f_$k -f python scripts/profile_step.py 2 > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log
done
timeout 900 ncu --set full --clock-control none -k regex:'k_conv1|k_conv_prep|k_cost_fill|k_cross_arms|k_cross_count|k_sgm_flags|k_wta_decode|k_lr_labels|k_lr_fill|k_subpixel|k_median|k_bilateral' -c 24 -o gpurun_out/r2f_small_kernels -f python scripts/profile_step.py 1 > gpurun_out/ncu_small.log 2>&1
tail -1 gpurun_out/ncu_small.log
