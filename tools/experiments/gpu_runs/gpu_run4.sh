set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2d_pytest.log; tail -12 gpurun_out/r2d_pytest.log
timeout 600 python bench.py --pairs 1 --units-per-step 16 --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2d_profile_cfg2_p1.txt > gpurun_out/r2d_bench_cfg2_p1.json 2> gpurun_out/r2d_bench_cfg2_p1.err; tail -c 300 gpurun_out/r2d_bench_cfg2_p1.err; cut -c1-250 gpurun_out/r2d_bench_cfg2_p1.json
