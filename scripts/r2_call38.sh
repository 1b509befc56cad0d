# everything in the SCF-level suite that can reach the n <= 128 tridiagonal route, plus the C++ layer tests
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_scf_gpu.py tests/test_cpp_layer.py -m gpu -x -q --timeout 900 -k "not c3_full and not c2_full and not c2_size and not c5_miniature and not large_path" > gpurun_out/r2_pytest38.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest38.log
