set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2i_pytest.log; tail -8 gpurun_out/r2i_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2i_profile_cfg2.txt > gpurun_out/r2i_bench_cfg2.json 2> gpurun_out/r2i_bench_cfg2.err; tail -c 300 gpurun_out/r2i_bench_cfg2.err; cat gpurun_out/r2i_bench_cfg2.json
timeout 600 python bench.py --config 3 --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2i_profile_cfg3.txt 2>/dev/null | tee gpurun_out/r2i_bench_cfg3.json | cut -c1-400
timeout 600 python bench.py --config 4 --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2i_profile_cfg4.txt 2>/dev/null | tee gpurun_out/r2i_bench_cfg4.json | cut -c1-400
timeout 600 python bench.py --config 5 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/r2i_bench_cfg5.json | cut -c1-400
timeout 600 python bench.py --mode class --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/r2i_bench_class.json | cut -c1-600
