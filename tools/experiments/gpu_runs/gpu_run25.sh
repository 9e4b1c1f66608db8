# r2z2: resize with the u8 -> half look-up table instead of a double division per pixel: GPU suite (resize / rectify bit-exactness) + one per-op profile
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r2z2_pytest.log
timeout 200 python bench.py --steps 3 --warmup 3 --soak 1 --no-cpu-baseline --profile-out gpurun_out/r2z2_profile.txt 2>/dev/null | cut -c1-160
grep -h "^resize\|conv1a" gpurun_out/r2z2_profile.txt
