set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r2_pytest.log; tail -5 gpurun_out/r2_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/r2_profile_cfg2.txt > gpurun_out/r2_bench_cfg2.json 2> gpurun_out/r2_bench_cfg2.err; tail -c 600 gpurun_out/r2_bench_cfg2.err; cut -c1-300 gpurun_out/r2_bench_cfg2.json
timeout 600 python bench.py --mode class --steps 20 --warmup 5 > gpurun_out/r2_bench_class.json 2> gpurun_out/r2_bench_class.err; tail -c 600 gpurun_out/r2_bench_class.err; cat gpurun_out/r2_bench_class.json | cut -c1-1500
for c in 3 4 5; do timeout 600 python bench.py --config $c --steps 5 --warmup 3 --profile-out gpurun_out/r2_profile_cfg$c.txt > gpurun_out/r2_bench_cfg$c.json 2> gpurun_out/r2_bench_cfg$c.err; tail -c 600 gpurun_out/r2_bench_cfg$c.err; cut -c1-400 gpurun_out/r2_bench_cfg$c.json; done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/r2_ncu_all_cfg2.csv python bench.py --steps 1 --warmup 1 --device-only --units-per-step 32 > gpurun_out/r2_ncu_all_cfg2.log 2>&1; tail -2 gpurun_out/r2_ncu_all_cfg2.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum --clock-control none --csv --log-file gpurun_out/r2_ncu_all_cfg3.csv python bench.py --config 3 --steps 1 --warmup 1 --device-only --units-per-step 8 > gpurun_out/r2_ncu_all_cfg3.log 2>&1; tail -2 gpurun_out/r2_ncu_all_cfg3.log
ls -la gpurun_out | tail -20
