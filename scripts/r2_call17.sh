python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 900 -k "syev" > gpurun_out/r2_pytest17.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest17.log
python -m pytest tests/test_scf_gpu.py -m gpu -q --timeout 900 -k "parity_large or c2_full or c5_mini or c4_two or eigenpairs" >> gpurun_out/r2_pytest17.log 2>&1; echo "pytest2 rc=$?"; tail -3 gpurun_out/r2_pytest17.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
$B --no-workloads > gpurun_out/r2_b17_c2.json 2>> gpurun_out/r2_b17.err
GSCF_B200_BT_PAIR=0 $B --no-workloads > gpurun_out/r2_b17_c2_nopair.json 2>> gpurun_out/r2_b17.err
$B --workload c3 --steps 2 > gpurun_out/r2_b17_c3.json 2>> gpurun_out/r2_b17.err
$B --workload c4 --steps 2 > gpurun_out/r2_b17_c4.json 2>> gpurun_out/r2_b17.err
$B --workload c5 --steps 2 > gpurun_out/r2_b17_c5.json 2>> gpurun_out/r2_b17.err
echo done
