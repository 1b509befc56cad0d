python -m pytest tests/test_scf_gpu.py -m gpu -q --timeout 900 -k "chained or rescue or hooks" > gpurun_out/r2_pytest8.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest8.log
python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 900 -k "syev" >> gpurun_out/r2_pytest8.log 2>&1; echo "pytest syev rc=$?"; tail -3 gpurun_out/r2_pytest8.log
python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b8_c4.json 2>> gpurun_out/r2_b8.err
python bench.py --workload c4 --problems 512 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b8_c4p512.json 2>> gpurun_out/r2_b8.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 2500 --csv --log-file gpurun_out/r2_launches_c4_4096.csv python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_c4_4096.log 2>&1; echo "ncu c4 rc=$?"
