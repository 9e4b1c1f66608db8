set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2f_pytest.log; tail -6 gpurun_out/r2f_pytest.log
timeout 300 python tools/trace_conv.py c32 c6432 c64 > gpurun_out/r2f_conv_trace.txt 2>&1; grep "TFLOP\|steady" gpurun_out/r2f_conv_trace.txt
AIRFE_NO_PREWAIT=1 timeout 300 python tools/prof_conv.py c32 c6432 c64 c128 c256 c96 2>&1 | grep TFLOP
timeout 300 python tools/prof_conv.py c32 c6432 c64 c128 c256 c96 2>&1 | grep TFLOP
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2f_profile_cfg2.txt > gpurun_out/r2f_bench_cfg2.json 2> gpurun_out/r2f_bench_cfg2.err; tail -c 300 gpurun_out/r2f_bench_cfg2.err; cut -c1-200 gpurun_out/r2f_bench_cfg2.json
AIRFE_NO_PREWAIT=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2f_profile_cfg2_noprewait.txt 2>/dev/null | cut -c1-200
timeout 600 python bench.py --config 3 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-200
