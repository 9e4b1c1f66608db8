set -x
mkdir -p gpurun_out
NCU="ncu --set full --import-source on --clock-control none"
AIRFE_CONV_FOLD=2 timeout 300 $NCU -k regex:fold -s 2 -c 1 -f -o gpurun_out/r2n_fold6432 python tools/prof_conv.py c6432 > /dev/null 2>&1
AIRFE_CONV_FOLD=2 timeout 300 $NCU -k regex:fold -s 2 -c 1 -f -o gpurun_out/r2n_fold64 python tools/prof_conv.py c64 > /dev/null 2>&1
timeout 300 $NCU -k regex:tc_ffn -s 40 -c 1 -f -o gpurun_out/r2n_ffn python tools/trace_match.py 47 > /dev/null 2>&1
timeout 300 $NCU -k regex:tc_attn -s 40 -c 1 -f -o gpurun_out/r2n_attn python tools/trace_match.py 47 > /dev/null 2>&1
timeout 300 $NCU -k regex:tc_gemm -s 40 -c 1 -f -o gpurun_out/r2n_gemm_qkv python tools/trace_match.py 47 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 600 python -m pytest tests/test_tc_gemm.py -m gpu -q -x -k "conv3x3" 2>&1 | tail -4
