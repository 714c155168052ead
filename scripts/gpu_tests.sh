#!/bin/bash
# GPU box: the whole gpu-marked suite + smoke.  Outputs in gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --tb=short -x -s 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
