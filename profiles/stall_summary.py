"""Stall-reason summary of an `ncu --page source --csv` export: whole kernel, the hottest instructions, and the
instructions executed `--exec N` times (e.g. the generator loop body)."""
import csv, sys, collections
path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[ix['# Samples']] or 0) for r in data)
print('total samples', tot)
cnt = collections.Counter(r[ix['Instructions Executed']] for r in data)
print('exec-count groups', [(k, v) for k, v in cnt.most_common(8)])
key = sys.argv[2] if len(sys.argv) > 2 else [k for k, v in cnt.most_common() if k != '0'][0]
grp = [r for r in data if r[ix['Instructions Executed']] == key]
s = sum(int(r[ix['# Samples']]) for r in grp)
print(f'group exec={key}: {len(grp)} instructions, {s} samples ({100*s/tot:.1f}% of kernel)')
agg = {st: sum(int(r[ix[st]] or 0) for r in grp) for st in stalls}
for st, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]:
    print(f'  {st:26s} {v:8d} {100*v/max(s,1):5.1f}%')
ops = collections.Counter(); opsamp = collections.Counter()
for r in grp:
    t = r[ix['Source']].split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    ops[op] += 1; opsamp[op] += int(r[ix['# Samples']])
print('  op mix:', ', '.join(f'{op}:{c}({opsamp[op]})' for op, c in ops.most_common(16)))
print('hottest instructions:')
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']] or 0))[:16]:
    st = {s_: int(r[ix[s_]] or 0) for s_ in stalls}
    top2 = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(' ', r[ix['# Samples']].rjust(6), r[ix['Instructions Executed']].rjust(9), r[ix['Source']].strip()[:58].ljust(58), top2)
