import csv,sys
from collections import defaultdict
path=sys.argv[1]
rows=list(csv.reader(open(path)))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='Line No']
h=rows[hdr[0]]; end=hdr[1]-2 if len(hdr)>1 else len(rows)
col={}
for i,n in enumerate(h): col.setdefault(n,i)
groups=[('philox',155,178),('bernoulli/quantile',179,196),('hazard_thr',197,220),('fire_index',221,228),('big_slot',229,238),('helpers',239,268),('infect_agent',269,321),('transition_agent',322,370),('stay/od',371,403),
('prologue',455,492),('household',493,521),('depwait/args',522,529),('outside thr',530,582),('draw',583,589),('infect funnel',590,595),('due funnel',596,604),('go_home',605,624),('go_out',625,687),('accum',688,762),('writeback',763,769),('epilogue',770,800)]
ins=defaultdict(float); smp=defaultdict(float); thr=defaultdict(float); ti=ts=0
for r in rows[hdr[0]+1:end]:
    if len(r)<len(h): continue
    try: line=int(r[0])
    except: continue
    n=float(r[col['Instructions Executed']] or 0); s=float(r[col['# Samples']] or 0)
    t=float(r[col['Thread Instructions Executed']] or 0) if 'Thread Instructions Executed' in col else 0
    g='other'
    for name,lo,hi in groups:
        if lo<=line<=hi: g=name;break
    ins[g]+=n; smp[g]+=s; thr[g]+=t; ti+=n; ts+=s
for name in [g[0] for g in groups]+['other']:
    if ins[name] or smp[name]:
        print(f'{name:20s} inst {100*ins[name]/ti:5.1f}%  smp {100*smp[name]/ts:5.1f}%  avg-active {thr[name]/max(ins[name],1):5.1f}')
print('total inst',ti)
