python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "collide or iterate" 2>&1 | tail -2
HCG_TAU1=0 python bench.py --steps 60 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('generic', round(d['value']), d['kernel_ms_per_step'].get('kernel:k_collide_stream'), d['roofline']['frac'])"
python bench.py --steps 60 --warmup 5 --cadence 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c5', round(d['value']), d['ms_per_step'])"
