python bench.py --workload cube --steps 200 --warmup 20 > gpurun_out/b_r1i_cube.json 2>gpurun_out/b.err; tail -c 1800 gpurun_out/b_r1i_cube.json; tail -3 gpurun_out/b.err
python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('main', round(d['value']), 'e2e', round(d['e2e']['value']), d['ms_per_step'])"
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
