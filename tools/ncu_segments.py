"""Warp samples, executed instructions and top stall reasons of a kernel between landmark instructions (barriers, exits, ...),
from the same source-page CSV: a quick phase breakdown of a warp-specialised kernel.
    python tools/ncu_segments.py kernel.csv [MARK1,MARK2,...]"""
import csv,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
# split into kernel sections
secs=[]; cur=None
for r in rows:
    if r and r[0]=='Kernel Name': cur={'name':r[1],'rows':[]}; secs.append(cur); continue
    if cur is not None: cur['rows'].append(r)
for sec in secs:
    hdr=sec['rows'][0]; n=len(hdr); data=[r for r in sec['rows'][1:] if len(r)>=n]
    ia=hdr.index('Source'); isamp=hdr.index('# Samples'); iex=hdr.index('Instructions Executed')
    stalls=[i for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    tot=sum(int(r[isamp]) for r in data)
    print(sec['name'][:110]); print(' total samples',tot,'instrs',len(data),'inst exec',sum(int(r[iex]) for r in data))
    marks=sys.argv[2].split(',') if len(sys.argv)>2 else ['BAR.SYNC','EXIT']
    start=0; acc=0; accx=0; cnt=collections.Counter(); hm=0
    for k,r in enumerate(data):
        acc+=int(r[isamp]); accx+=int(r[iex])
        if 'HMMA' in r[ia]: hm+=int(r[iex])
        for i in stalls: cnt[hdr[i]]+=int(r[i])
        if any(m in r[ia] for m in marks) or k==len(data)-1:
            if acc>tot*0.004: print(' ',start,k,'samples %.1f%%'%(100*acc/tot),'inst',accx,'hmma',hm,cnt.most_common(4), r[ia].strip()[:30])
            start=k+1; acc=0; accx=0; cnt=collections.Counter(); hm=0
