# final state: whole GPU suite, smoke, the driver's bench command on one GPU (+ the reference arm)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 1200 > gpurun_out/r2_pytest31.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest31.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke31.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke31.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench31_default.json 2> gpurun_out/r2_bench31.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2_bench31_reference.json 2>> gpurun_out/r2_bench31.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench31_default.json').read().strip().splitlines()[-1])
print('C2', d['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['avg_launch_ms'], 'fp64', d['fp64_roofline']['frac'], 'cpu', d['cpu_baseline']['value'], 'launches', d['gpu_launches'])
for k,w in d['workloads'].items():
    print(k, w['value'], w['unit'], 'e2e', w['e2e']['value'], 'fp64', w['fp64_roofline']['frac'], 'cpu', w['cpu_baseline']['value'], w['cpu_baseline']['unit'])
r=json.loads(open('gpurun_out/r2_bench31_reference.json').read().strip().splitlines()[-1])
print('reference arm', r.get('value'), r.get('unit'), r.get('cpu_baseline',{}).get('cores'))
PY
tail -3 gpurun_out/r2_bench31.err
