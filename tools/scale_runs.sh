set -u
mkdir -p gpurun_out
run() { n=$1; shift; out=$1; shift; if [ $n -eq 1 ]; then timeout 400 python bench.py --gpus 1 "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err; else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $n "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err; fi; tail -c 300 gpurun_out/$out.json | head -c 10 >/dev/null; }
run 8 r2m_bench_n8_c1 --steps 100 --warmup 10
run 4 r2m_bench_n4_c1 --steps 100 --warmup 10
run 1 r2m_bench_n1_c1 --steps 100 --warmup 10 --no-cpu-baseline
run 4 r2m_bench_stenosis_n4 --workload stenosis --steps 60 --warmup 10
run 2 r2m_bench_pipeflow_n2 --workload pipeflow --steps 100 --warmup 10
run 8 r2m_bench_n8_c5 --steps 100 --warmup 10 --cadence 5
ls -la gpurun_out/r2m_*
