# Round-end evidence on TWO B200s of one box (tag r2y): the NCCL tests and the 2-rank bench arms of configs 2 and 5.
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -q 2>&1 | tail -8 > gpurun_out/r2y_pytest_2gpu.log; tail -5 gpurun_out/r2y_pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2y_bench_cfg2_n2.json 2> gpurun_out/r2y_bench_cfg2_n2.err; tail -c 300 gpurun_out/r2y_bench_cfg2_n2.err; cut -c1-300 gpurun_out/r2y_bench_cfg2_n2.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config 5 --steps 10 --warmup 3 > gpurun_out/r2y_bench_cfg5_n2.json 2> gpurun_out/r2y_bench_cfg5_n2.err; tail -c 300 gpurun_out/r2y_bench_cfg5_n2.err; cut -c1-400 gpurun_out/r2y_bench_cfg5_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-200
