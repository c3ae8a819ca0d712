"""Most-sampled SASS instructions of one kernel with their dominant stall reason, from the CSV of
`ncu -i rep --page source --csv --print-source sass` (one kernel section per file, as tools/gpu_ncu_source.sh writes them).
    python tools/ncu_top_stalls.py kernel.csv [N]"""
import csv,sys,collections
f=sys.argv[1]; N=int(sys.argv[2]) if len(sys.argv)>2 else 25
rows=list(csv.reader(open(f)))
hdr=rows[1]; n=len(hdr); data=[r for r in rows[2:] if len(r)>=n]
ia=hdr.index('Source'); isamp=hdr.index('# Samples'); iex=hdr.index('Instructions Executed')
stalls=[i for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot=sum(int(r[isamp]) for r in data)
print(rows[0][1][:100],'samples',tot,'exec',sum(int(r[iex]) for r in data))
allst=collections.Counter()
for r in data:
    for i in stalls: allst[hdr[i]]+=int(r[i])
print(' stalls:',[(k,round(100*v/tot,1)) for k,v in allst.most_common(8)])
top=sorted(range(len(data)), key=lambda k:-int(data[k][isamp]))[:N]
for k in sorted(top):
    r=data[k]; st=max(stalls,key=lambda i:int(r[i]))
    print('  %5d %5.1f%% x%-9s %-14s %s'%(k,100*int(r[isamp])/tot,r[iex],hdr[st][6:],r[ia].strip()[:80]))
