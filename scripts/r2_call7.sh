python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 900 -k "fused_variants or syev" > gpurun_out/r2_pytest7.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest7.log
python -m pytest tests/test_scf_gpu.py -m gpu -q --timeout 900 -k "c4_two_waves or ensemble or randomised" >> gpurun_out/r2_pytest7.log 2>&1; echo "pytest2 rc=$?"; tail -3 gpurun_out/r2_pytest7.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B --workload c4 > gpurun_out/r2_b7_c4.json 2>> gpurun_out/r2_b7.err
GSCF_B200_TF_SOLO_SPLIT=0 $B --workload c4 > gpurun_out/r2_b7_c4_nosplit.json 2>> gpurun_out/r2_b7.err
$B --workload c4 --problems 512 > gpurun_out/r2_b7_c4p512.json 2>> gpurun_out/r2_b7.err
GSCF_B200_TF_SOLO_SPLIT=0 $B --workload c4 --problems 512 > gpurun_out/r2_b7_c4p512_nosplit.json 2>> gpurun_out/r2_b7.err
GSCF_B200_TF_BATCH_GLOBAL=8 $B --workload c5 --problems 32 > gpurun_out/r2_b7_c5p32_g8.json 2>> gpurun_out/r2_b7.err
GSCF_B200_TF_BATCH_GLOBAL=2 $B --workload c5 > gpurun_out/r2_b7_c5_g2occ2.json 2>> gpurun_out/r2_b7.err
echo done
