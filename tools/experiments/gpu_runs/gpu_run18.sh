set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gemm.py -m gpu -q -x 2>&1 | tail -6
timeout 900 python -m pytest tests/test_match_gpu.py tests/test_detect_gpu.py tests/test_batch_invariance_gpu.py tests/test_configs_gpu.py tests/test_cpp_surface.py tests/test_graphs_gpu.py -m gpu -q 2>&1 | tail -8
AIRFE_GEMM_TMA_STORE=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2r_profile_cfg2_stg.txt 2>/dev/null | cut -c1-200
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2r_profile_cfg2.txt 2>/dev/null | cut -c1-200
grep "tc_gemm" gpurun_out/r2r_profile_cfg2_stg.txt | awk '{print $2,$3,$4, $(NF-3)}' | sort | uniq -c | sort -rn | head -14
grep "tc_gemm" gpurun_out/r2r_profile_cfg2.txt | awk '{print $2,$3,$4, $(NF-3)}' | sort | uniq -c | sort -rn | head -14
