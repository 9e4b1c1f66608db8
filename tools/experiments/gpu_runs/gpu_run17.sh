set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gemm.py -m gpu -q -x -k "conv3x3" 2>&1 | tail -4
AIRFE_CONV_FOLD_KB2=0 timeout 300 python tools/prof_conv.py c12864 c12864s 2>&1 | grep TFLOP
timeout 300 python tools/prof_conv.py c12864 c12864s 2>&1 | grep TFLOP
timeout 900 python -m pytest tests/test_detect_gpu.py tests/test_batch_invariance_gpu.py tests/test_configs_gpu.py -m gpu -q 2>&1 | tail -6
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2q_profile_cfg2.txt 2>/dev/null | cut -c1-200
grep "128->64" gpurun_out/r2q_profile_cfg2.txt | head -12
