"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel."""
import csv, collections, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
hdr=None; agg=collections.defaultdict(lambda:[0,0.0]); seq=[]
for r in rows:
    if r[0]=='ID': hdr=r; continue
    if hdr is None: continue
    d=dict(zip(hdr,r))
    if 'gpu__time_duration' in d.get('Metric Name',''):
        v=float(d['Metric Value'].replace(',','')); unit=d['Metric Unit']
        v*= {'ns':1e-6,'us':1e-3,'ms':1.0,'s':1e3}.get(unit,1.0)
        k=d['Kernel Name'].split('(')[0][-48:]; agg[k][0]+=1; agg[k][1]+=v; seq.append((k,v))
skip=int(sys.argv[2]) if len(sys.argv)>2 else 0
if skip:
    agg=collections.defaultdict(lambda:[0,0.0])
    for k,v in seq[skip:]: agg[k][0]+=1; agg[k][1]+=v
tot=sum(v[1] for v in agg.values())
print(f'launches {sum(v[0] for v in agg.values())} total {tot:.3f} ms (skipped first {skip})')
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]: print(f'{v[1]:10.3f} ms {v[0]:5d}x {100*v[1]/tot:5.1f}%  {k}')
