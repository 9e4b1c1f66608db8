set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gemm.py -m gpu -q -x -k "conv3x3" 2>&1 | tail -4
AIRFE_FOLD_WIDE=0 timeout 600 python -m pytest tests/test_tc_gemm.py -m gpu -q -x -k "kx_fold" 2>&1 | tail -2
AIRFE_CONV_FOLD=0 timeout 300 python tools/prof_conv.py c64 c64np c64po c6432 c32 2>&1 | grep TFLOP
AIRFE_CONV_FOLD=2 AIRFE_FOLD_WIDE=0 timeout 300 python tools/prof_conv.py c64 c64np c64po c6432 c32 2>&1 | grep TFLOP
AIRFE_CONV_FOLD=2 timeout 300 python tools/prof_conv.py c64 c64np c64po c6432 c32 2>&1 | grep TFLOP
