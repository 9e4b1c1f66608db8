set -x
mkdir -p gpurun_out
timeout 120 tools/probe_ldtm > gpurun_out/r2m_probe_ldtm.txt 2>&1; cat gpurun_out/r2m_probe_ldtm.txt
timeout 300 python tools/trace_match.py 47 > gpurun_out/r2m_trace_match.txt 2>&1; tail -12 gpurun_out/r2m_trace_match.txt
timeout 900 python -m pytest tests/test_match_gpu.py tests/test_batch_invariance_gpu.py tests/test_cpp_surface.py tests/test_graphs_gpu.py -m gpu -q 2>&1 | tail -25 > gpurun_out/r2m_pytest_match.log; grep "Error\|passed\|failed" gpurun_out/r2m_pytest_match.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2m_profile_cfg2.txt 2>/dev/null | cut -c1-250
grep "tc_attn\|tc_ffn\|256->768\|256->512" gpurun_out/r2m_profile_cfg2.txt | head -6
