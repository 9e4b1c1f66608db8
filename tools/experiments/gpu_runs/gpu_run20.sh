# r2t: LOI map with a 160-float pixel stride, ragged / degenerate frame tests, library-call size sweep (device-only loop)
set -x
mkdir -p gpurun_out
T=r2t
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/${T}_pytest.log; tail -6 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/${T}_profile.txt 2>gpurun_out/${T}.err | cut -c1-240
for p in 47 64 74 94; do timeout 300 python bench.py --pairs $p --units-per-step $((p*8)) --steps 4 --warmup 3 --soak 2 --device-only 2>>gpurun_out/${T}.err; done
tail -n 3 gpurun_out/${T}.err
