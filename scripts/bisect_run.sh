timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slab.py -m gpu -q -x --tb=short -k "features or slab or pipeline or config1" 2>&1 | tail -5
for s in "500 3000 400 2" "2000 768 400 3 p2p"; do echo "== $s"; timeout 600 python scripts/slab_bisect.py $s 2>&1 | tail -16; done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stages_ms']['features'], d.get('cbca_ms_per_round_per_volume'))"
