python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 900 -k "syev" > gpurun_out/r2_pytest27.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest27.log
python -m pytest tests/test_scf_gpu.py -m gpu -q --timeout 1500 -k "c3_full_size" >> gpurun_out/r2_pytest27.log 2>&1; echo "pytest2 rc=$?"; tail -3 gpurun_out/r2_pytest27.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B --workload c3 > gpurun_out/r2_b27_c3.json 2>> gpurun_out/r2_b27.err
echo done
