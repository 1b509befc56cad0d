python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 900 -k "syev" > gpurun_out/r2_pytest14.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest14.log
python -m pytest tests/test_scf_gpu.py -m gpu -q --timeout 900 -k "eigenpairs or c5_mini or parity_large" >> gpurun_out/r2_pytest14.log 2>&1; echo "pytest2 rc=$?"; tail -3 gpurun_out/r2_pytest14.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B --workload c3 --n-eigenpairs 512 > gpurun_out/r2_b14_c3_k512.json 2>> gpurun_out/r2_b14.err
$B --workload c3 --n-eigenpairs 1024 > gpurun_out/r2_b14_c3_k1024.json 2>> gpurun_out/r2_b14.err
$B --no-workloads --n-eigenpairs 128 > gpurun_out/r2_b14_c2_k128.json 2>> gpurun_out/r2_b14.err
echo done
