python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 900 -k "syev" > gpurun_out/r2_pytest22.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest22.log
python -m pytest tests/test_scf_gpu.py -m gpu -q --timeout 900 -k "c2_size or c4_two or eigenpairs" >> gpurun_out/r2_pytest22.log 2>&1; echo "pytest2 rc=$?"; tail -3 gpurun_out/r2_pytest22.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B --workload c4 > gpurun_out/r2_b22_c4.json 2>> gpurun_out/r2_b22.err
$B --workload c4 --problems 512 > gpurun_out/r2_b22_c4p512.json 2>> gpurun_out/r2_b22.err
$B --workload c5 > gpurun_out/r2_b22_c5.json 2>> gpurun_out/r2_b22.err
echo done
