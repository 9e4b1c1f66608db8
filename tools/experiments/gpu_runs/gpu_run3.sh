set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2c_pytest.log; tail -12 gpurun_out/r2c_pytest.log
timeout 600 python bench.py --mode class --steps 20 --warmup 5 > gpurun_out/r2c_bench_class.json 2> gpurun_out/r2c_bench_class.err; tail -c 400 gpurun_out/r2c_bench_class.err; cut -c1-200 gpurun_out/r2c_bench_class.json; grep -o '"latency_ms": {[^}]*}' gpurun_out/r2c_bench_class.json
AIRFE_NO_GRAPH=1 timeout 600 python bench.py --mode class --steps 20 --warmup 5 2>/dev/null | grep -o '"latency_ms": {[^}]*}'
timeout 600 python bench.py --config 3 --mode class --steps 20 --warmup 5 2>/dev/null | grep -o '"value": [0-9.]*\|"latency_ms": {[^}]*}' | head -3
timeout 600 python bench.py --config 3 --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2c_profile_cfg3.txt > gpurun_out/r2c_bench_cfg3.json 2> gpurun_out/r2c_bench_cfg3.err; tail -c 400 gpurun_out/r2c_bench_cfg3.err; cut -c1-200 gpurun_out/r2c_bench_cfg3.json
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2c_profile_cfg2.txt > gpurun_out/r2c_bench_cfg2.json 2> gpurun_out/r2c_bench_cfg2.err; tail -c 400 gpurun_out/r2c_bench_cfg2.err; cut -c1-200 gpurun_out/r2c_bench_cfg2.json
