import json, sys, csv, collections
d=json.load(open('gpurun_out/bench.json'))
print("clips/s", round(d['value'],2), "ms/step", round(d['ms_per_step'],1), "e2e", round(d['e2e']['value'],2), "launches", d['gpu_launches'], d['clocks'])
r=d["roofline"]; print("dominant kernel TF/s", round(r['achieved'],1), "frac", round(r['frac'],3), "e2e frac", round(r['end_to_end_frac_of_bf16_peak'],3))
for k,v in r['per_kernel'].items(): print("  ", k, v)
try:
    with open('gpurun_out/launches.csv') as f:
        lines=[l for l in f if not l.startswith('==')]
    agg=collections.OrderedDict()
    for row in csv.DictReader(lines):
        k=(row['Kernel Name'][:48],row['Grid Size']); t=float(row['Metric Value'].replace(',',''))
        a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=t
    tot=sum(a[1] for a in agg.values())
    for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:int(sys.argv[1]) if len(sys.argv)>1 else 16]:
        print(f"{a[1]/1e3:9.1f} us {100*a[1]/tot:5.1f}% n={a[0]:3d} avg={a[1]/a[0]/1e3:7.1f} us {k[0]} {k[1]}")
except Exception as e: print(e)
