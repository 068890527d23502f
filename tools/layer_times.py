import json,sys,collections
d=json.load(open(sys.argv[1]))
g=collections.defaultdict(list)
for x in d:
    if x['kind']=='conv' and x['R']==1 and x['Cout']==128 and 'trans' not in x['name']: g[('1x1',x['H'])].append(x['ms']*1e3)
    if x['kind']=='conv' and x['R']==3: g[('3x3',x['H'])].append(x['ms']*1e3)
for k,v in sorted(g.items()): print(k, 'sum %.0f us'%sum(v), [round(t) for t in v[:6]], '...', [round(t) for t in v[-2:]])
