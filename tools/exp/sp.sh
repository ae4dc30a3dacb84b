for t in 256 128 352 672; do HCG_SPREAD_THREADS=$t python bench.py --steps 60 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$t', round(d['value']), d['kernel_ms_per_step']['spreadParticleForce'])"; done
