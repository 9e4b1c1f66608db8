# r2u: 16 epilogue warps (EW = 4) for the nine-tap halo conv kernel and for tc_gemm: parity under the wide kernels, per-op A/B
set -x
mkdir -p gpurun_out
T=r2u
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/${T}_pytest_default.log; tail -3 gpurun_out/${T}_pytest_default.log
AIRFE_CONV_WIDE_MAXN=512 AIRFE_GEMM_WIDE=1 AIRFE_PARITY_OUT=gpurun_out/${T}_parity_wide.json timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/${T}_pytest_wide.log; tail -6 gpurun_out/${T}_pytest_wide.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/${T}_profile_default.txt 2>gpurun_out/${T}.err | cut -c1-200
AIRFE_CONV_WIDE_MAXN=512 AIRFE_GEMM_WIDE=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/${T}_profile_wide.txt 2>>gpurun_out/${T}.err | cut -c1-200
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/${T}_profile_default2.txt 2>>gpurun_out/${T}.err | cut -c1-200
tail -n 3 gpurun_out/${T}.err
