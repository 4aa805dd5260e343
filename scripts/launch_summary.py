import csv, collections, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    agg[r[ki][:70]][0]+=1; agg[r[ki][:70]][1]+=v
tot=sum(v[1] for v in agg.values())
print("total %.3f ms" % (tot/1e6))
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:25]:
    print(f"{v[1]/1e6:10.3f} ms {v[0]:4d}  {100*v[1]/tot:5.1f}%  {k}")
