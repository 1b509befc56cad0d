# 4-D tensor maps with row groups of 4 (bank-conflict free) against row groups of 8
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --timeout 900 -k "dgemm or pulay or density" > gpurun_out/r2_pytest30.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest30.log
GSCF_B200_GEMM_RG=8 timeout 300 python scripts/gemm_bench.py > gpurun_out/r2_gemm30_rg8.txt 2>&1; echo "rc=$?"
timeout 300 python scripts/gemm_bench.py > gpurun_out/r2_gemm30_rg4.txt 2>&1; echo "rc=$?"
GSCF_B200_GEMM_TMA=4 timeout 300 python scripts/gemm_bench.py > gpurun_out/r2_gemm30_rg4_lvl4.txt 2>&1; echo "rc=$?"
paste -d'|' gpurun_out/r2_gemm30_rg8.txt gpurun_out/r2_gemm30_rg4.txt gpurun_out/r2_gemm30_rg4_lvl4.txt | sed 's/dgemm //g;s/ TFLOP\/s//g' | cut -c1-200
B="timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-workloads"
for t in 3 4; do GSCF_B200_GEMM_TMA=$t $B --workload c2 --steps 10 > gpurun_out/r2_b30_c2_tma$t.json 2>> gpurun_out/r2_b30.err; done
for t in 3 4; do GSCF_B200_GEMM_TMA=$t $B --workload c4 > gpurun_out/r2_b30_c4_tma$t.json 2>> gpurun_out/r2_b30.err; done
GSCF_B200_GEMM_RG=8 $B --workload c3 > gpurun_out/r2_b30_c3_rg8.json 2>> gpurun_out/r2_b30.err
$B --workload c3 > gpurun_out/r2_b30_c3_rg4.json 2>> gpurun_out/r2_b30.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_b30_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); c=d['config']; ph=c['phase_ms_last_step']
        print(f.split('r2_b30_')[1], round(d['value'],2), d['unit'], 'scf_ms', round(c['scf_ms_per_step'],1), 'fock', round(c['fock_ms_per_step'],1), {k:round(ph[k],1) for k in ('eig','tridiag','stedc','backtransform','error','overlap','density')})
    except Exception as ex:
        print(f, 'ERR', ex)
PY
tail -5 gpurun_out/r2_b30.err
timeout 300 ncu --metrics gpu__time_duration.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:dgemm_dmma -c 60 --csv --log-file gpurun_out/r2_gemm30_conflicts.csv python scripts/gemm_bench.py > gpurun_out/r2_ncu30.log 2>&1; echo "ncu rc=$?"
