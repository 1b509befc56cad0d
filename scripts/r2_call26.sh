GSCF_B200_TF_SOLO_OW=1 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 900 -k "syev" > gpurun_out/r2_pytest26.log 2>&1; echo "pytest OW rc=$?"; tail -3 gpurun_out/r2_pytest26.log
GSCF_B200_TF_SOLO_OW=1 python -m pytest tests/test_scf_gpu.py -m gpu -q --timeout 900 -k "c4_two or c2_size" >> gpurun_out/r2_pytest26.log 2>&1; echo "pytest2 rc=$?"; tail -3 gpurun_out/r2_pytest26.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B --workload c4 > gpurun_out/r2_b26_c4.json 2>> gpurun_out/r2_b26.err
GSCF_B200_TF_SOLO_OW=1 $B --workload c4 > gpurun_out/r2_b26_c4_ow.json 2>> gpurun_out/r2_b26.err
GSCF_B200_TF_SOLO_OW=1 $B --workload c4 --problems 512 > gpurun_out/r2_b26_c4p512_ow.json 2>> gpurun_out/r2_b26.err
GSCF_B200_TF_SOLO_OW=1 $B --no-workloads > gpurun_out/r2_b26_c2_ow.json 2>> gpurun_out/r2_b26.err
echo done
