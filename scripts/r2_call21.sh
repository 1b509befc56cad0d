python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 900 -k "fused_variants" > gpurun_out/r2_pytest21.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest21.log
python -m pytest tests/test_scf_gpu.py -m gpu -q --timeout 900 -k "c5_mini or toda_default or c2_size" >> gpurun_out/r2_pytest21.log 2>&1; echo "pytest2 rc=$?"; tail -3 gpurun_out/r2_pytest21.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B --workload c5 --problems 32 > gpurun_out/r2_b21_c5p32.json 2>> gpurun_out/r2_b21.err
GSCF_B200_TF_SYMH_PF=0 $B --workload c5 --problems 32 > gpurun_out/r2_b21_c5p32_nopf.json 2>> gpurun_out/r2_b21.err
$B --workload c5 --problems 16 > gpurun_out/r2_b21_c5p16.json 2>> gpurun_out/r2_b21.err
echo done
