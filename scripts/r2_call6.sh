python -m pytest tests -m gpu -q --timeout 1500 > gpurun_out/r2_pytest6.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest6.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err; echo "bench rc=$?"
python bench.py --workload c3t --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench6_c3t.json 2>> gpurun_out/r2_bench6.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-workloads --no-cpu-baseline > gpurun_out/r2_ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r2_launches_c5_p32.csv python bench.py --workload c5 --problems 32 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_c5.log 2>&1; echo "ncu c5 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:trd_fused_kernel -s 40 -c 1 -o gpurun_out/r2_prof_symh python bench.py --workload c5 --problems 32 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_full_c5.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/*.ncu-rep
