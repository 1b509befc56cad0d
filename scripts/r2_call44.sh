# last check of the round: the larger SCF parity cases with the final library
timeout 110 python -m pytest tests/test_scf_gpu.py -m gpu -x -q --timeout 100 -k "c5_miniature or large_path or c2_full or n_eigenpairs" 2>&1 | tail -2
