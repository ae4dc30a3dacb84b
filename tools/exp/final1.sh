python -m pytest tests -m gpu -x -q --durations=6 2>&1 | tail -12
python bench.py --steps 100 --warmup 10 > gpurun_out/b_r1k_c1.json 2>gpurun_out/b.err; tail -c 600 gpurun_out/b_r1k_c1.json
python bench.py --steps 100 --warmup 10 --cadence 5 --no-cpu-baseline > gpurun_out/b_r1k_c5.json 2>gpurun_out/b.err; python - <<EOF
import json
for f in ("gpurun_out/b_r1k_c1.json","gpurun_out/b_r1k_c5.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["roofline"]["kernel"], round(d["roofline"]["frac"],3), {k:round(v,3) for k,v in d["kernel_ms_per_step"].items()})
EOF
