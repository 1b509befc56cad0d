# the driver's 8-GPU bench command (C2 replicas + C4 / C5 sharded, C-ABI NCCL gather)
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench33_n8.json 2> gpurun_out/r2_bench33.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench33_n8.json').read().strip().splitlines()[-1])
print('C2', d['value'], 'n_gpus', d['n_gpus'])
for k,w in d['workloads'].items():
    print(k, w['value'], w['unit'], 'e2e', w['e2e']['value'], w['config'].get('problems_per_gpu'), w['config'].get('records_gathered'), w['config'].get('gather'))
PY
tail -3 gpurun_out/r2_bench33.err
