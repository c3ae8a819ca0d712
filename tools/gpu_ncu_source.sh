#!/bin/bash
# source-level ncu capture of the big kernels of one sequential step
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out /tmp/ncu gpurun_out/${TAG:-src}_src
T=${TAG:-src}
timeout 300 python tools/run_step.py --steps 2 > gpurun_out/${T}_plain.log 2>&1 || exit 1
timeout 1500 ncu --set full --import-source on --clock-control none -k regex:"tc_gemm_kernel|ffn_fused|pda_encode|group_attention_h" -s 45 -c 45 -o /tmp/ncu/step python tools/run_step.py --steps 2 > gpurun_out/${T}_ncu.log 2>&1
ncu -i /tmp/ncu/step.ncu-rep --page raw --csv > gpurun_out/${T}_step_raw.csv 2>/dev/null
ncu -i /tmp/ncu/step.ncu-rep --page source --csv --print-source sass > /tmp/ncu/all_source.csv 2>/dev/null
python - <<'PY'
import csv, re, os, sys
T=os.environ.get('TAG','src')
rows=csv.reader(open('/tmp/ncu/all_source.csv'))
secs=[]; cur=None
for r in rows:
    if r and r[0]=='Kernel Name':
        cur={'name':r[1],'rows':[r]}; secs.append(cur); continue
    if cur is not None: cur['rows'].append(r)
best={}
for s in secs:
    hdr=s['rows'][1]; 
    try: isamp=hdr.index('# Samples')
    except ValueError: continue
    tot=sum(int(r[isamp]) for r in s['rows'][2:] if len(r)>isamp and r[isamp].isdigit())
    key=re.sub(r'\(int\)|\(bool\)|void |<unnamed>::|\(.*$','',s['name'])
    if key not in best or tot>best[key][0]: best[key]=(tot,s)
keep=['ffn_fused_kernel','tc_gemm_kernel<4, 256, 2, 2, 2, 2>','tc_gemm_kernel<4, 256, 1, 2, 0, 2>','tc_gemm_kernel<4, 256, 1, 2, 3, 2>',
      'tc_gemm_kernel<4, 256, 1, 1, 1, 2>','tc_gemm_kernel<4, 256, 1, 2, 1, 2>','pda_encode_ln_kernel<128>','pda_encode_ln_kernel<64>','group_attention_h_kernel<32, 128>',
      'tc_gemm_kernel<4, 192, 1, 2, 5, 2>','tc_gemm_kernel<4, 256, 1, 2, 4, 2>']
for k,(tot,s) in best.items():
    print(k, tot)
    if any(k.startswith(x) for x in keep):
        fn='gpurun_out/%s_src/%s.csv'%(T,re.sub(r'[^A-Za-z0-9]+','_',k))
        w=csv.writer(open(fn,'w')); [w.writerow(r) for r in s['rows']]
PY
ls -la gpurun_out/${T}_src | head -20
