# r2v: tc_ffn with 16 epilogue warps (EW = 4): parity (whole GPU suite with it on) and A/B; the adopted conv / gemm EW rules are the default in both runs
set -x
mkdir -p gpurun_out
T=r2v
AIRFE_FFN_WIDE=1 AIRFE_PARITY_OUT=gpurun_out/${T}_parity_ffnwide.json timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/${T}_pytest_ffnwide.log; tail -6 gpurun_out/${T}_pytest_ffnwide.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/${T}_pytest_default.log; tail -3 gpurun_out/${T}_pytest_default.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/${T}_profile_default.txt 2>gpurun_out/${T}.err | cut -c1-200
AIRFE_FFN_WIDE=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/${T}_profile_ffnwide.txt 2>>gpurun_out/${T}.err | cut -c1-200
AIRFE_FFN_WIDE=1 timeout 600 python bench.py --config 3 --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/${T}_profile_cfg3_ffnwide.txt 2>>gpurun_out/${T}.err | cut -c1-200
timeout 600 python bench.py --config 3 --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/${T}_profile_cfg3_default.txt 2>>gpurun_out/${T}.err | cut -c1-200
tail -n 3 gpurun_out/${T}.err
