# per-kernel sequence of one batched eigensolve (n = 128; 4096 and 512 matrices)
mkdir -p gpurun_out
for b in 4096 512; do
  timeout 300 python scripts/eig_only.py 128 3 2 $b 0 > gpurun_out/r2_eig35_$b.txt 2>&1; tail -2 gpurun_out/r2_eig35_$b.txt
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_eig35_$b.csv python scripts/eig_only.py 128 2 2 $b 0 > gpurun_out/r2_ncu35_$b.log 2>&1; echo "ncu rc=$?"
done
python - <<'PY'
import csv,re
for b in (4096,512):
    lines=[l for l in open(f'gpurun_out/r2_eig35_{b}.csv') if not l.startswith('==')]
    rows=list(csv.DictReader(lines))
    seq=[]
    for r in rows:
        name=re.sub(r'^void ','',r['Kernel Name']); name=re.split(r'[<(]',name.replace('gscfb::','').replace('<unnamed>::',''))[0]
        v=float(r['Metric Value'].replace(',','')); u=r['Metric Unit']; v=v/1000 if u=='ns' else (v*1000 if u=='ms' else v)
        seq.append((name,v,r.get('Grid Size',''),r.get('Block Size','')))
    # last eigensolve = from the last dc_scale... find last 'absmax' or first kernel of syev: use last occurrence of 'scale_extreme' or 'absmax'
    idx=[i for i,s in enumerate(seq) if s[0].startswith('absmax')]
    start=idx[-1] if idx else 0
    tot=sum(s[1] for s in seq[start:])
    print(f'--- batch {b}: kernels of the last eigensolve ({len(seq)-start} launches, sum {tot/1000:.3f} ms)')
    for s in seq[start:]: print(f'   {s[0]:32s} {s[1]:9.1f} us  grid {s[2]} block {s[3]}')
PY
