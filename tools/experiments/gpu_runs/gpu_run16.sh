set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_match_gpu.py tests/test_batch_invariance_gpu.py tests/test_cpp_surface.py tests/test_graphs_gpu.py tests/test_detect_gpu.py tests/test_configs_gpu.py -m gpu -q 2>&1 | tail -25 > gpurun_out/r2p_pytest.log; grep "Error\|passed\|failed" gpurun_out/r2p_pytest.log
timeout 300 python tools/trace_match.py 47 > gpurun_out/r2p_trace_match.txt 2>&1; tail -6 gpurun_out/r2p_trace_match.txt
AIRFE_ATTN_NP=2 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2p_profile_cfg2_np2.txt 2>/dev/null | cut -c1-200
grep "tc_attn\|256->768\|conv1a" gpurun_out/r2p_profile_cfg2_np2.txt | head -4
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2p_profile_cfg2.txt 2>/dev/null | cut -c1-200
grep "tc_attn\|256->768\|conv1a\|kx-fold" gpurun_out/r2p_profile_cfg2.txt | head -6
