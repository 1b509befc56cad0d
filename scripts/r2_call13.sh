python scripts/geneig_fuzz.py > gpurun_out/r2_geneig_fuzz.log 2>&1; echo "geneig_fuzz rc=$?"; tail -3 gpurun_out/r2_geneig_fuzz.log
python -m pytest tests -m gpu -q --timeout 900 -k "functionality or example_config1 or parity_small or c2_full or error_paths or unrestricted or cpp" > gpurun_out/r2_pytest13.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest13.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
$B --no-workloads > gpurun_out/r2_b13_c2.json 2>> gpurun_out/r2_b13.err
$B --workload c4 --steps 2 > gpurun_out/r2_b13_c4.json 2>> gpurun_out/r2_b13.err
$B --workload c3 --steps 2 > gpurun_out/r2_b13_c3.json 2>> gpurun_out/r2_b13.err
echo done
