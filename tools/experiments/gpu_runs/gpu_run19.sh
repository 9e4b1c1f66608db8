# r2s: fused nearest-2x up-sampling store in the producer conv's epilogue + fc1|fc3|fc4 as one GEMM: full GPU suite, A/B per-op profiles
set -x
mkdir -p gpurun_out
T=r2s
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/${T}_pytest.log; tail -6 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/${T}_profile_new.txt 2>gpurun_out/${T}_new.err | cut -c1-240
AIRFE_UP2_FUSE=0 AIRFE_FC134_MERGE=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/${T}_profile_old.txt 2>gpurun_out/${T}_old.err | cut -c1-240
tail -3 gpurun_out/${T}_new.err gpurun_out/${T}_old.err
grep -c . gpurun_out/${T}_profile_new.txt gpurun_out/${T}_profile_old.txt
