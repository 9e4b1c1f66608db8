# r2w: LOI endpoint features per junction (default on), conv1a PX = 4 variant: GPU suite, per-op A/B
set -x
mkdir -p gpurun_out
T=r2w
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/${T}_pytest.log; tail -4 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/${T}_profile_default.txt > gpurun_out/${T}_bench_default.json 2>gpurun_out/${T}.err; cut -c1-200 gpurun_out/${T}_bench_default.json
AIRFE_CONV1A_PX=4 AIRFE_LOI_JF=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/${T}_profile_px4_nojf.txt 2>>gpurun_out/${T}.err | cut -c1-200
AIRFE_CONV1A_PX=4 timeout 300 python -m pytest tests/test_detect_gpu.py -m gpu -q 2>&1 | tail -3
grep -h "conv1a\|loi_gather" gpurun_out/${T}_profile_default.txt gpurun_out/${T}_profile_px4_nojf.txt
tail -n 3 gpurun_out/${T}.err
