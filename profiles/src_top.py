"""Summarise an `ncu --page source --csv` export: top SASS instructions by stall samples."""
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
h=rows[1]
data=[]
for r in rows[2:]:
    if len(r)!=len(h): 
        if data: break
        continue
    if r[0]=='Address':
        if data: break
        continue
    data.append(r)
ix={c:i for i,c in enumerate(h)}
S=lambda r,k:int(float(r[ix[k]] or 0))
tot=sum(S(r,'# Samples') for r in data)
print('total samples',tot,'instrs',len(data))
N=int(sys.argv[2]) if len(sys.argv)>2 else 40
for r in sorted(data,key=lambda r:-S(r,'# Samples'))[:N]:
    n=S(r,'# Samples')
    st={k[6:]:S(r,k) for k in h if k.startswith('stall_') and 'Not Issued' not in k and S(r,k)>0.15*max(n,1)}
    print(str(n).rjust(7), f'{100*n/tot:5.1f}%', r[ix['Instructions Executed']].rjust(11), r[ix['Source']].strip()[:64].ljust(64), st, 'xs',r[ix['L1 Wavefronts Shared Excessive']])
