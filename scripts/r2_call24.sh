python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 900 -k "syev or stedc or eig" > gpurun_out/r2_pytest24.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest24.log
GSCF_B200_TF_SOLO_REG_T=512 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 900 -k "syev" >> gpurun_out/r2_pytest24.log 2>&1; echo "pytest T512 rc=$?"; tail -3 gpurun_out/r2_pytest24.log
GSCF_B200_TF_SOLO_REG_T=512 python -m pytest tests/test_scf_gpu.py -m gpu -q --timeout 900 -k "c4_two" >> gpurun_out/r2_pytest24.log 2>&1; echo "pytest2 rc=$?"; tail -3 gpurun_out/r2_pytest24.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B --no-workloads > gpurun_out/r2_b24_c2.json 2>> gpurun_out/r2_b24.err
$B --workload c4 > gpurun_out/r2_b24_c4.json 2>> gpurun_out/r2_b24.err
$B --workload c4 --problems 512 > gpurun_out/r2_b24_c4p512.json 2>> gpurun_out/r2_b24.err
GSCF_B200_TF_SOLO_REG_T=512 $B --workload c4 > gpurun_out/r2_b24_c4_t512.json 2>> gpurun_out/r2_b24.err
GSCF_B200_TF_SOLO_REG_T=512 $B --workload c4 --problems 512 > gpurun_out/r2_b24_c4p512_t512.json 2>> gpurun_out/r2_b24.err
$B --workload c5 > gpurun_out/r2_b24_c5.json 2>> gpurun_out/r2_b24.err
echo done
