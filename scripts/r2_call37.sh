# fused small-matrix back-transformation v2 (staged Z, two reflectors per barrier): tests and A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --timeout 900 -k "tridiag_dc or syevx or special or nan_input" > gpurun_out/r2_pytest37.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest37.log
timeout 900 python -m pytest tests/test_scf_gpu.py -m gpu -x -q --timeout 900 -k "c4 or ensemble or randomised or example or functionality or partial" >> gpurun_out/r2_pytest37.log 2>&1; echo "pytest2 rc=$?"; tail -3 gpurun_out/r2_pytest37.log
for b in 4096 512; do
  timeout 300 python scripts/eig_only.py 128 3 2 $b 1 2>&1 | tail -6
done
B="timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-workloads"
$B --workload c4 > gpurun_out/r2_b37_c4_bt_small2.json 2>> gpurun_out/r2_b37.err
$B --workload c4 --problems 512 > gpurun_out/r2_b37_c4_p512_bt_small2.json 2>> gpurun_out/r2_b37.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_b37_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); c=d['config']; ph=c['phase_ms_last_step']
        print(f.split('r2_b37_')[1], round(d['value'],2), d['unit'], 'pps', round(c.get('problems_converged_per_s',0),1), 'scf_ms', round(c['scf_ms_per_step'],1), {k:round(ph[k],1) for k in ('eig','tridiag','stedc','backtransform','error','overlap')})
    except Exception as ex:
        print(f, 'ERR', ex)
PY
tail -3 gpurun_out/r2_b37.err
