python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 100 --warmup 10 > gpurun_out/b_r1h_c1.json 2>gpurun_out/b.err; tail -c 2500 gpurun_out/b_r1h_c1.json
python bench.py --steps 100 --warmup 10 --cadence 5 --no-cpu-baseline > gpurun_out/b_r1h_c5.json 2>gpurun_out/b.err; tail -c 600 gpurun_out/b_r1h_c5.json
