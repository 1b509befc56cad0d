# tensor maps for both operand layouts (GSCF_B200_GEMM_TMA=3): tests, A/B, then ncu evidence for the round's final kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --timeout 900 -k "dgemm or pulay or density" > gpurun_out/r2_pytest29.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest29.log
for t in 2 3 4; do
  GSCF_B200_GEMM_TMA=$t timeout 300 python scripts/gemm_bench.py > gpurun_out/r2_gemm29_tma$t.txt 2>&1; echo "gemm tma=$t rc=$?"
done
paste -d'|' gpurun_out/r2_gemm29_tma2.txt gpurun_out/r2_gemm29_tma3.txt gpurun_out/r2_gemm29_tma4.txt | sed 's/dgemm //g;s/ TFLOP\/s//g' | cut -c1-200
B="timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-workloads"
for t in 1 3 4; do GSCF_B200_GEMM_TMA=$t $B --workload c2 --steps 10 > gpurun_out/r2_b29_c2_tma$t.json 2>> gpurun_out/r2_b29.err; done
for t in 2 3; do GSCF_B200_GEMM_TMA=$t $B --workload c3 > gpurun_out/r2_b29_c3_tma$t.json 2>> gpurun_out/r2_b29.err; done
for t in 1 3 4; do GSCF_B200_GEMM_TMA=$t $B --workload c4 > gpurun_out/r2_b29_c4_tma$t.json 2>> gpurun_out/r2_b29.err; done
for t in 2 3; do GSCF_B200_GEMM_TMA=$t $B --workload c5 > gpurun_out/r2_b29_c5_tma$t.json 2>> gpurun_out/r2_b29.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_b29_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); c=d['config']; ph=c['phase_ms_last_step']
        print(f.split('r2_b29_')[1], round(d['value'],2), d['unit'], 'scf_ms', round(c['scf_ms_per_step'],1), 'fock', round(c['fock_ms_per_step'],1), {k:round(ph[k],1) for k in ('eig','tridiag','stedc','backtransform','error','overlap','density')})
    except Exception as ex:
        print(f, 'ERR', ex)
PY
tail -5 gpurun_out/r2_b29.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2_launches_c2_final.csv python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/r2_ncu29_c2.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trd_fused_kernel -s 3 -c 1 -o gpurun_out/r2_prof_trd_rega python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/r2_ncu29_full_c2.log 2>&1; echo "ncu full c2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dgemm_dmma_tm_kernel -s 6 -c 10 -o gpurun_out/r2_prof_gemm_tm python scripts/gemm_bench.py > gpurun_out/r2_ncu29_full_gemm.log 2>&1; echo "ncu full gemm rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trd_fused_blocked -s 1 -c 1 -o gpurun_out/r2_prof_trd_blocked python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-workloads > gpurun_out/r2_ncu29_full_c3.log 2>&1; echo "ncu full c3 rc=$?"
ls -la gpurun_out/*.ncu-rep
