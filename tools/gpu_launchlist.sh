#!/bin/bash
# ncu launch list (gpu__time_duration.sum) of the bench command + per-kernel share summary
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
T=${TAG:-r4s}
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_steps2.json 2> /dev/null || exit 1
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_launches.log 2>&1
python - <<'PY'
import csv, json, re, collections, os
T=os.environ.get('TAG','r4s')
rows=[r for r in csv.reader(open('gpurun_out/%s_launches.csv'%os.environ.get('TAG','r4s'))) if len(r)>5]
hdr=rows[0]
ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); iu=hdr.index('Metric Unit')
agg=collections.defaultdict(lambda:[0,0.0])
n=0
for r in rows[1:]:
    try: v=float(r[iv].replace(',',''))
    except ValueError: continue
    u=r[iu]; ms = v/1e6 if u.startswith('ns') else (v/1e3 if u.startswith('us') else v)
    name=re.sub(r'[<(].*','',r[ik].replace('void ','').replace('(anonymous namespace)::','').replace('<unnamed>::',''))
    a=agg[name]; a[0]+=1; a[1]+=ms; n+=1
tot=sum(a[1] for a in agg.values())
out={"source":"ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 python bench.py --steps 2 --warmup 3 --no-cpu-baseline (first 6000 launches: warm-up steps, graph captures and replays; cold-cache, serialised durations: compare shares)",
     "launches":n,"kernels":[{"kernel":k,"share":round(a[1]/tot,4),"ms":round(a[1],3),"launches":a[0]} for k,a in sorted(agg.items(),key=lambda kv:-kv[1][1])[:24]]}
json.dump(out,open('gpurun_out/%s_launches_summary.json'%T,'w'),indent=1)
for k in out["kernels"][:14]: print(k)
PY
