python -m pytest tests/test_scf_gpu.py -m gpu -q --timeout 900 -k "toda" > gpurun_out/r2_pytest19.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest19.log
