GSCF_B200_BT_SMALL=2 timeout 22 python scripts/eig_only.py 128 3 2 4096 1 2>&1 | tail -6
timeout 16 python scripts/eig_only.py 128 3 2 4096 0 2>&1 | tail -2
