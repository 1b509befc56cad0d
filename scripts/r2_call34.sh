# robustness sweeps at the end of round 2 (evidence, not gates): eigensolver stress, randomised GEMM and SCF sweeps
mkdir -p gpurun_out
timeout 600 python scripts/eig_stress.py quick > gpurun_out/r2_stress34_eig.txt 2>&1; echo "eig_stress rc=$?"; tail -3 gpurun_out/r2_stress34_eig.txt
timeout 600 python scripts/gemm_fuzz.py 400 2000 > gpurun_out/r2_stress34_gemm.txt 2>&1; echo "gemm_fuzz rc=$?"; tail -2 gpurun_out/r2_stress34_gemm.txt
GSCF_B200_GEMM_TMA=4 timeout 600 python scripts/gemm_fuzz.py 400 3000 > gpurun_out/r2_stress34_gemm_lvl4.txt 2>&1; echo "gemm_fuzz lvl4 rc=$?"; tail -2 gpurun_out/r2_stress34_gemm_lvl4.txt
timeout 900 python scripts/scf_fuzz.py 120 5000 > gpurun_out/r2_stress34_scf.txt 2>&1; echo "scf_fuzz rc=$?"; tail -3 gpurun_out/r2_stress34_scf.txt
