set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r2g_pytest_2gpu.log; tail -8 gpurun_out/r2g_pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config 5 --steps 5 --warmup 3 > gpurun_out/r2g_bench_cfg5_n2.json 2> gpurun_out/r2g_bench_cfg5_n2.err; tail -c 500 gpurun_out/r2g_bench_cfg5_n2.err; cut -c1-300 gpurun_out/r2g_bench_cfg5_n2.json
timeout 600 python bench.py --config 5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench_cfg5_n1.json 2> gpurun_out/r2g_bench_cfg5_n1.err; cut -c1-300 gpurun_out/r2g_bench_cfg5_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2g_bench_cfg2_n2.json 2> gpurun_out/r2g_bench_cfg2_n2.err; tail -c 300 gpurun_out/r2g_bench_cfg2_n2.err; cut -c1-300 gpurun_out/r2g_bench_cfg2_n2.json
