mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s7_tests.log 2>&1; tail -3 gpurun_out/s7_tests.log
timeout 300 python bench.py --no-cpu-baseline --no-other-mode > gpurun_out/s7_bench.json 2> gpurun_out/s7_bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"slab_l1|colsum|rows" --csv --log-file gpurun_out/s7_small_kernels.csv python bench.py --steps 3 --warmup 3 --no-graph --no-other-mode --no-cpu-baseline > gpurun_out/s7_ncu.log 2>&1
python - <<'P'
import json,csv,collections
d=json.loads(open('gpurun_out/s7_bench.json').read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), round(d['value']), round(d['e2e']['value']))
rows=list(csv.reader(open('gpurun_out/s7_small_kernels.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; kn=h.index('Kernel Name'); mv=h.index('Metric Value'); gi=h.index('Grid Size')
dd=collections.defaultdict(list)
for r in rows[hi+1:]:
    if len(r)>mv and r[mv] and r[mv][0].isdigit(): dd[(r[kn][:50],r[gi])].append(float(r[mv].replace(',',''))/1000)
for k,v in dd.items(): print(k,len(v),round(sum(v)/len(v),1))
P
