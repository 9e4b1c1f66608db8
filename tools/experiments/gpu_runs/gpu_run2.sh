set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r2b_pytest.log; tail -15 gpurun_out/r2b_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2b_profile_cfg2.txt > gpurun_out/r2b_bench_cfg2.json 2> gpurun_out/r2b_bench_cfg2.err; tail -c 600 gpurun_out/r2b_bench_cfg2.err; cut -c1-300 gpurun_out/r2b_bench_cfg2.json
timeout 600 python bench.py --config 3 --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2b_profile_cfg3.txt > gpurun_out/r2b_bench_cfg3.json 2> gpurun_out/r2b_bench_cfg3.err; tail -c 600 gpurun_out/r2b_bench_cfg3.err; cut -c1-300 gpurun_out/r2b_bench_cfg3.json
AIRFE_SINKHORN_V1=1 AIRFE_NMS_V1=1 timeout 600 python -m pytest tests/test_match_gpu.py tests/test_detect_gpu.py -m gpu -q 2>&1 | tail -5
