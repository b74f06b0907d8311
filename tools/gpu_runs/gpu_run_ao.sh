#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
for c in 1 0; do
RB_CTABLE=$c timeout 900 ncu --target-processes all --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/ao_launches_ctable$c.csv python tools/readme_bench.py 30000 > $O/ao_readme$c.json 2>> $O/ao.err
python - <<P
import csv,collections
rows=list(csv.reader(l for l in open('gpurun_out/ao_launches_ctable$c.csv') if l.startswith('"')))
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); iu=hdr.index('Metric Unit')
agg=collections.defaultdict(lambda:[0,0.0])
U={'ns':1e-6,'us':1e-3,'ms':1.0,'s':1e3,'usecond':1e-3,'msecond':1.0,'nsecond':1e-6,'second':1e3}
for r in rows[1:]:
    try: v=float(r[iv].replace(',',''))*U.get(r[iu],1e-6)
    except: continue
    a=agg[r[ik][:60]]; a[0]+=1; a[1]+=v
print("RB_CTABLE=$c")
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:12]: print("  %-60s n=%5d total %.2f ms"%(k,v[0],v[1]))
P
done
tail -n 3 $O/ao.err
