set -u
mkdir -p gpurun_out
for k in k_cbca_pass k_sgm_pass k_cost_volume_tc k_conv64_tc k_wta; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 2 -o gpurun_out/r1z_$k -f python scripts/profile_step.py 2 > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cbca_march -s 1 -c 1 -o gpurun_out/r1z_k_cbca_march -f python scripts/cbca_one.py 3 > gpurun_out/ncu_k_cbca_march.log 2>&1
echo == c4; timeout 600 python bench.py --steps 5 --warmup 3 --workload c4 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_c4.json | cut -c1-400
echo == c5 one gpu; timeout 600 python bench.py --steps 3 --warmup 3 --workload c5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_c5_n1.json | cut -c1-400
echo == modes; timeout 300 python scripts/cbca_modes.py 192 --all 2>&1 | tail -12 | tee gpurun_out/cbca_modes.txt
