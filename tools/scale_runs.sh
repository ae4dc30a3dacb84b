# weak-scaling lines of the benchmark on one 8-GPU box (run under `gpurun --gpus 8`): N = 8 and N = 1 back to back
set -u
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/r2p_bench_n8_c1.json 2> gpurun_out/r2p_bench_n8_c1.err
timeout 400 python bench.py --gpus 1 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2p_bench_n1_c1.json 2> gpurun_out/r2p_bench_n1_c1.err
ls -la gpurun_out/r2p_*
