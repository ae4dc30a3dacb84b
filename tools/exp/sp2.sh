python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -5
for t in 1 0; do HCG_SPREAD_BULK=$t python bench.py --steps 60 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bulk=$t', round(d['value']), d['kernel_ms_per_step']['spreadParticleForce'])"; done
