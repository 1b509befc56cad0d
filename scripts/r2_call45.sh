timeout 80 python -m pytest tests/test_scf_gpu.py -m gpu -x -q --timeout 80 -k "fresh_process" 2>&1 | tail -2
