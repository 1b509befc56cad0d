python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 900 -k "syev" > gpurun_out/r2_pytest23.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest23.log
python -m pytest tests/test_scf_gpu.py -m gpu -q --timeout 900 -k "c4_two or c5_mini" >> gpurun_out/r2_pytest23.log 2>&1; echo "pytest2 rc=$?"; tail -3 gpurun_out/r2_pytest23.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B --no-workloads > gpurun_out/r2_b23_c2.json 2>> gpurun_out/r2_b23.err
$B --workload c4 > gpurun_out/r2_b23_c4.json 2>> gpurun_out/r2_b23.err
$B --workload c4 --problems 512 > gpurun_out/r2_b23_c4p512.json 2>> gpurun_out/r2_b23.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches_c4_p4096.csv python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:trd_solo_reg -s 2 -c 2 -o gpurun_out/r2_prof_solo_reg python bench.py --workload c4 --problems 512 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_full_c4.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/*.ncu-rep
echo done
