python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -15
for t in 1 0; do HCG_TAU1=$t python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tau1=$t', round(d['value']), d['ms_per_step'], d['roofline']['frac'], d['kernel_ms_per_step'])"; done
