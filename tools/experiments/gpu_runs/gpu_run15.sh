set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gemm.py -m gpu -q -x -k "conv3x3" 2>&1 | tail -4
AIRFE_CONV_FOLD=2 timeout 300 python tools/prof_conv.py c64 c64np c64po c6432 c32 2>&1 | grep TFLOP
AIRFE_CONV_FOLD=2 timeout 300 python tools/prof_conv.py c64 c64np c64po c6432 c32 2>&1 | grep TFLOP
AIRFE_CONV_FOLD=2 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2o_profile_cfg2_foldall.txt 2>/dev/null | cut -c1-200
grep "kx-fold" gpurun_out/r2o_profile_cfg2_foldall.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2o_profile_cfg2.txt 2>/dev/null | cut -c1-200
grep "Bres" gpurun_out/r2o_profile_cfg2.txt
