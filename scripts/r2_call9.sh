python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 900 -k "syev" > gpurun_out/r2_pytest9.log 2>&1; echo "pytest syev rc=$?"; tail -3 gpurun_out/r2_pytest9.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B --workload c4 > gpurun_out/r2_b9_c4.json 2>> gpurun_out/r2_b9.err
GSCF_B200_DC_SECULAR=warp $B --workload c4 > gpurun_out/r2_b9_c4_secwarp.json 2>> gpurun_out/r2_b9.err
GSCF_B200_GEMM_SMALLK=64 $B --workload c4 > gpurun_out/r2_b9_c4_smallk64.json 2>> gpurun_out/r2_b9.err
GSCF_B200_GEMM_SMALLK=128 $B --workload c4 > gpurun_out/r2_b9_c4_smallk128.json 2>> gpurun_out/r2_b9.err
$B --workload c4 --problems 512 > gpurun_out/r2_b9_c4p512.json 2>> gpurun_out/r2_b9.err
GSCF_B200_DC_SECULAR=thread $B --workload c5 > gpurun_out/r2_b9_c5_secthread.json 2>> gpurun_out/r2_b9.err
$B --workload c5 > gpurun_out/r2_b9_c5.json 2>> gpurun_out/r2_b9.err
echo done
