# Round-end evidence run on ONE B200 (tag r2y -> profiles/r02e_*): full GPU test suite, bench arms of every config, ncu launch list / all-kernel
# metrics of a config-2 chunk, full ncu captures of the kernels whose epilogue changed late in round 2 (EW = 4 instantiations).
set -x
mkdir -p gpurun_out
T=r2y
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/${T}_pytest.log; tail -4 gpurun_out/${T}_pytest.log
timeout 900 python bench.py --profile-out gpurun_out/${T}_profile_cfg2.txt > gpurun_out/${T}_bench_cfg2.json 2> gpurun_out/${T}_bench_cfg2.err; tail -c 300 gpurun_out/${T}_bench_cfg2.err; cut -c1-300 gpurun_out/${T}_bench_cfg2.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref_cfg2.json 2>/dev/null; cut -c1-300 gpurun_out/${T}_bench_ref_cfg2.json
for c in 3 4 5; do timeout 400 python bench.py --config $c --steps 10 --warmup 3 --profile-out gpurun_out/${T}_profile_cfg$c.txt > gpurun_out/${T}_bench_cfg$c.json 2> gpurun_out/${T}_bench_cfg$c.err; tail -c 200 gpurun_out/${T}_bench_cfg$c.err; cut -c1-300 gpurun_out/${T}_bench_cfg$c.json; done
timeout 300 python bench.py --mode class --steps 10 --warmup 3 > gpurun_out/${T}_bench_class.json 2>/dev/null; cut -c1-700 gpurun_out/${T}_bench_class.json
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_cfg2.csv python bench.py --steps 1 --warmup 1 --device-only --units-per-step 47 > gpurun_out/${T}_launches_cfg2.log 2>&1; tail -1 gpurun_out/${T}_launches_cfg2.log | cut -c1-200
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${T}_ncu_all_cfg2.csv python bench.py --steps 1 --warmup 1 --device-only --units-per-step 47 > gpurun_out/${T}_ncu_all_cfg2.log 2>&1; tail -1 gpurun_out/${T}_ncu_all_cfg2.log | cut -c1-200
NCU="ncu --set full --import-source on --clock-control none"
timeout 200 $NCU -k regex:tc_conv3x3 -s 2 -c 1 -f -o gpurun_out/${T}_full_conv_64_64_pool python tools/prof_conv.py c64 > /dev/null 2>&1
timeout 200 $NCU -k regex:tc_ffn -s 40 -c 1 -f -o gpurun_out/${T}_full_ffn python tools/trace_match.py 47 > /dev/null 2>&1
timeout 200 $NCU -k regex:tc_gemm -s 40 -c 1 -f -o gpurun_out/${T}_full_gemm_qkv python tools/trace_match.py 47 > /dev/null 2>&1
ls -la gpurun_out/${T}_*
