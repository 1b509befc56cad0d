python scripts/toda_probe.py > gpurun_out/r2_toda_probe2.log 2>&1; echo "probe rc=$?"
python -m pytest tests -m gpu -q --timeout 900 -k "toda or syev or jacobi or example_config1 or functionality" > gpurun_out/r2_pytest4.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest4.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
GSCF_B200_TF_SYMH_OCC=2 $B --workload c5 > gpurun_out/r2_b4_c5_occ2.json 2>> gpurun_out/r2_b4.err
GSCF_B200_TF_BATCH_GLOBAL=2 $B --workload c5 > gpurun_out/r2_b4_c5_g2.json 2>> gpurun_out/r2_b4.err
GSCF_B200_TF_BATCH_GLOBAL=8 $B --workload c5 > gpurun_out/r2_b4_c5_g8.json 2>> gpurun_out/r2_b4.err
GSCF_B200_TF_BATCH_GLOBAL=8 GSCF_B200_TF_SYMH_OCC=2 $B --workload c5 > gpurun_out/r2_b4_c5_g8occ2.json 2>> gpurun_out/r2_b4.err
GSCF_B200_TF_SYMH_OCC=2 $B --workload c5 --problems 32 > gpurun_out/r2_b4_c5p32_occ2.json 2>> gpurun_out/r2_b4.err
GSCF_B200_TF_BATCH_GLOBAL=2 $B --workload c5 --problems 32 > gpurun_out/r2_b4_c5p32_g2.json 2>> gpurun_out/r2_b4.err
$B --workload c4 > gpurun_out/r2_b4_c4.json 2>> gpurun_out/r2_b4.err
GSCF_B200_DC_LEAF=16 $B --workload c4 > gpurun_out/r2_b4_c4_leaf16.json 2>> gpurun_out/r2_b4.err
$B --workload c4 --problems 512 > gpurun_out/r2_b4_c4p512.json 2>> gpurun_out/r2_b4.err
GSCF_B200_DC_LEAF=16 $B --workload c4 --problems 512 > gpurun_out/r2_b4_c4p512_leaf16.json 2>> gpurun_out/r2_b4.err
echo "bench done"
