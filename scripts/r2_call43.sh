# exact batch extent in the tensor maps + strict box rule: the failing sequence, GEMM tests, ensemble tests, quick benches
mkdir -p gpurun_out
K="functionality or example or hooks or diis_coeff or max_iter or synthetic_parity_small or unrestricted or ensemble_lockstep"
timeout 200 python -m pytest tests/test_scf_gpu.py -m gpu -x -q --timeout 200 -k "$K" 2>&1 | grep -E "passed|failed|Segmentation" | head -2
GSCF_B200_TMAP_STRICT=0 timeout 200 python -m pytest tests/test_scf_gpu.py -m gpu -x -q --timeout 200 -k "$K" 2>&1 | grep -E "passed|failed|Segmentation" | head -2
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --timeout 300 -k "dgemm or pulay or density" 2>&1 | tail -1
timeout 300 python -m pytest tests/test_scf_gpu.py -m gpu -x -q --timeout 300 -k "c4 or toda or ensemble or chained" 2>&1 | tail -1
B="timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-workloads"
$B --workload c2 --steps 10 > gpurun_out/r2_b43_c2.json 2>> gpurun_out/r2_b43.err
$B --workload c4 > gpurun_out/r2_b43_c4.json 2>> gpurun_out/r2_b43.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_b43_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); c=d['config']; ph=c['phase_ms_last_step']
        print(f.split('r2_b43_')[1], round(d['value'],2), d['unit'], 'pps', round(c.get('problems_converged_per_s',0),1), 'scf_ms', round(c['scf_ms_per_step'],1), {k:round(ph[k],1) for k in ('eig','tridiag','stedc','backtransform','error','overlap')})
    except Exception as ex:
        print(f, 'ERR', ex)
PY
