# Round-end evidence on TWO B200s of one box (tag r2y): smoke(), the NCCL tests and a short 2-rank bench of config 2.
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r2y_smoke.log
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -q 2>&1 | tail -8 > gpurun_out/r2y_pytest_2gpu.log; tail -5 gpurun_out/r2y_pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --soak 1 > gpurun_out/r2y_bench_cfg2_n2.json 2> gpurun_out/r2y_bench_cfg2_n2.err; tail -c 300 gpurun_out/r2y_bench_cfg2_n2.err; cut -c1-300 gpurun_out/r2y_bench_cfg2_n2.json
