B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B --workload c5 > gpurun_out/r2_b10_c5.json 2>> gpurun_out/r2_b10.err
GSCF_B200_GEMM_SMALLK=0 $B --workload c5 > gpurun_out/r2_b10_c5_smallk0.json 2>> gpurun_out/r2_b10.err
$B --workload c4 --problems 512 > gpurun_out/r2_b10_c4p512.json 2>> gpurun_out/r2_b10.err
$B --workload c4 > gpurun_out/r2_b10_c4.json 2>> gpurun_out/r2_b10.err
$B --workload c5 --problems 32 > gpurun_out/r2_b10_c5p32.json 2>> gpurun_out/r2_b10.err
GSCF_B200_GEMM_SMALLK=0 $B --workload c5 --problems 32 > gpurun_out/r2_b10_c5p32_smallk0.json 2>> gpurun_out/r2_b10.err
python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 900 -k "gemm or pulay or density" > gpurun_out/r2_pytest10.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest10.log
echo done
