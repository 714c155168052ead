#!/bin/bash
# CBCA chained-round variants: parity tests + timings (MCCNN_CBCA_CHAIN = 0 cp.async narrow, 1 TMA narrow, 2 TMA wide, 3 cp.async wide)
set -u
mkdir -p gpurun_out
for v in 1 2; do
  echo "== variant $v tests"
  MCCNN_CBCA_CHAIN=$v timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cbca or flat_image or pipeline_object" 2>&1 | tail -5
done
for v in 0 1 2 3; do
  echo "== variant $v timing"
  MCCNN_CBCA_CHAIN=$v timeout 600 python scripts/cbca_modes.py 192 2>&1 | grep -v "^$"
done 2>&1 | tee gpurun_out/exp_chain.txt
