python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 900 -k "syev" > gpurun_out/r2_pytest16.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest16.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
$B --no-workloads > gpurun_out/r2_b16_c2.json 2>> gpurun_out/r2_b16.err
GSCF_B200_TF_SOLO_TAIL=0 $B --no-workloads > gpurun_out/r2_b16_c2_notail.json 2>> gpurun_out/r2_b16.err
$B --workload c3 --steps 2 > gpurun_out/r2_b16_c3.json 2>> gpurun_out/r2_b16.err
echo done
