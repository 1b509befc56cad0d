# full GPU suite at HEAD + A/B: tensor-map TMA GEMM path, leaf-kernel shared memory sized by the leaf order
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 1200 > gpurun_out/r2_pytest28.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest28.log
for t in 1 2; do
  GSCF_B200_GEMM_TMA=$t timeout 300 python scripts/gemm_bench.py > gpurun_out/r2_gemm28_tma$t.txt 2>&1; echo "gemm tma=$t rc=$?"
done
paste -d'|' gpurun_out/r2_gemm28_tma1.txt gpurun_out/r2_gemm28_tma2.txt | cut -c1-220
B="timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-workloads"
GSCF_B200_DC_LEAF_SMEM=full $B --workload c4 --problems 512 > gpurun_out/r2_b28_c4_p512_leafsmem_full.json 2>> gpurun_out/r2_b28.err
$B --workload c4 --problems 512 > gpurun_out/r2_b28_c4_p512_leafsmem_sized.json 2>> gpurun_out/r2_b28.err
GSCF_B200_DC_LEAF=32 $B --workload c4 --problems 512 > gpurun_out/r2_b28_c4_p512_leaf32.json 2>> gpurun_out/r2_b28.err
$B --workload c4 > gpurun_out/r2_b28_c4_default.json 2>> gpurun_out/r2_b28.err
GSCF_B200_DC_LEAF=16 $B --workload c4 > gpurun_out/r2_b28_c4_leaf16.json 2>> gpurun_out/r2_b28.err
GSCF_B200_GEMM_TMA=1 $B --workload c3 > gpurun_out/r2_b28_c3_tma1.json 2>> gpurun_out/r2_b28.err
$B --workload c3 > gpurun_out/r2_b28_c3_tma2.json 2>> gpurun_out/r2_b28.err
GSCF_B200_GEMM_TMA=1 $B --workload c2 --steps 10 > gpurun_out/r2_b28_c2_tma1.json 2>> gpurun_out/r2_b28.err
$B --workload c2 --steps 10 > gpurun_out/r2_b28_c2_tma2.json 2>> gpurun_out/r2_b28.err
GSCF_B200_GEMM_TMA=1 $B --workload c5 > gpurun_out/r2_b28_c5_tma1.json 2>> gpurun_out/r2_b28.err
$B --workload c5 > gpurun_out/r2_b28_c5_tma2.json 2>> gpurun_out/r2_b28.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_b28_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); c=d['config']; ph=c['phase_ms_last_step']
        print(f.split('r2_b28_')[1], round(d['value'],2), d['unit'], 'scf_ms', round(c['scf_ms_per_step'],1), {k:round(ph[k],1) for k in ('eig','tridiag','stedc','backtransform','error','overlap')})
    except Exception as ex:
        print(f, 'ERR', ex)
PY
tail -5 gpurun_out/r2_b28.err
