set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2h_pytest.log; tail -8 gpurun_out/r2h_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2h_profile_cfg2.txt > gpurun_out/r2h_bench_cfg2.json 2> gpurun_out/r2h_bench_cfg2.err; tail -c 300 gpurun_out/r2h_bench_cfg2.err; cut -c1-200 gpurun_out/r2h_bench_cfg2.json
AIRFE_LG_PADDED=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-200
for P in 23 40 47; do timeout 600 python bench.py --pairs $P --units-per-step $((P*12)) --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2h_profile_cfg2_p$P.txt 2>/dev/null | cut -c1-200; done
timeout 600 python bench.py --config 5 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-200
