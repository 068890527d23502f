"""Per-layer comparison of the streaming dense-layer kernel with the conv1x1 + conv3x3 kernel pair (blocks 1-2, e4m3).
usage (under gpurun):
  B200_ENGINE_LAYERFUSE=0 PROF_TAG=pl0 python tools/prof_steps.py fp8 256 10; B200_ENGINE_LAYERFUSE=1 PROF_TAG=pl1 python tools/prof_steps.py fp8 256 10
  python tools/perlayer_fuse.py gpurun_out/pl0_steps_fp8_bs256.json gpurun_out/pl1_steps_fp8_bs256.json"""
import json, sys
a=json.load(open(sys.argv[1])); b=json.load(open(sys.argv[2]))
# a: default (pair of kernels), b: fused (time on the 1x1 step, ~0 on the 3x3)
i=0
rows=[]
while i < len(a):
    p=a[i]
    if p['kind']=='conv' and p.get('R')==1 and i+1<len(a) and a[i+1].get('R')==3 and p['H'] in (56,28):
        un=p['ms']+a[i+1]['ms']; fu=b[i]['ms']+b[i+1]['ms']
        rows.append((p['H'],p['Cin'],un*1e3,fu*1e3))
        i+=2
    else: i+=1
for r in rows: print(f"H{r[0]} Cin {r[1]:4d}: pair {r[2]:6.1f} us  fused {r[3]:6.1f} us  {'FUSED' if r[3]<r[2] else ''}")
print('sum pair', sum(r[2] for r in rows), 'fused', sum(r[3] for r in rows), 'best-of', sum(min(r[2],r[3]) for r in rows))
