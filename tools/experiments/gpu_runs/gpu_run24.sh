# r2z: GPU suite after the late host-side C-ABI hardening (create guard, argument validation, reloc table regrow, line-association overflow order)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r2z_pytest.log
