set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gemm.py -m gpu -q -x -k "conv3x3" 2>&1 | tail -15 > gpurun_out/r2j_pytest_conv.log; tail -12 gpurun_out/r2j_pytest_conv.log
AIRFE_CONV_FOLD=0 timeout 300 python tools/prof_conv.py c64 c64np c64po c6432 c32 2>&1 | grep TFLOP
timeout 300 python tools/prof_conv.py c64 c64np c64po c6432 c32 2>&1 | grep TFLOP
timeout 900 python -m pytest tests/test_detect_gpu.py tests/test_batch_invariance_gpu.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2j_pytest_detect.log; tail -12 gpurun_out/r2j_pytest_detect.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2j_profile_cfg2.txt 2>/dev/null | cut -c1-250
