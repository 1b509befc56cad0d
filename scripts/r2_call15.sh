python tests/syev_variant_main.py 600 1 > gpurun_out/r2_pw16_variant.log 2>&1; echo "variant rc=$?"
GSCF_B200_TF_GLOBAL=1 GSCF_B200_TF_UNBLOCKED_BELOW=64 python tests/syev_variant_main.py 600 1 >> gpurun_out/r2_pw16_variant.log 2>&1; echo "variant blocked rc=$?"
GSCF_B200_TF_GLOBAL=1 GSCF_B200_TF_UNBLOCKED_BELOW=100 python tests/syev_variant_main.py 257 3 >> gpurun_out/r2_pw16_variant.log 2>&1; echo "variant blocked2 rc=$?"
python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b15_c3_pw16.json 2>> gpurun_out/r2_b15.err; echo "bench rc=$?"
