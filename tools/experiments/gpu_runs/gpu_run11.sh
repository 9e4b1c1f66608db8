set -x
mkdir -p gpurun_out
timeout 300 python tools/trace_conv.py c6432 c32 c64np > gpurun_out/r2l_trace_fold.txt 2>&1; grep -v "^ *[0-9]* " gpurun_out/r2l_trace_fold.txt | tail -12; sed -n 1,60p gpurun_out/r2l_trace_fold.txt | grep "^ *1[0-9] \|^ *2[0-4] " | head -30
timeout 900 python -m pytest tests/test_match_gpu.py tests/test_batch_invariance_gpu.py tests/test_cpp_surface.py tests/test_graphs_gpu.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2l_pytest_match.log; tail -6 gpurun_out/r2l_pytest_match.log
timeout 300 python tools/trace_match.py 47 > gpurun_out/r2l_trace_match.txt 2>&1; tail -12 gpurun_out/r2l_trace_match.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2l_profile_cfg2.txt 2>/dev/null | cut -c1-250
grep "tc_attn\|tc_ffn\|256->768\|256->512" gpurun_out/r2l_profile_cfg2.txt | head -6
