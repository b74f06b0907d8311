#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
for c in 0 1 0 1 0 1; do
  RB_CTABLE=$c timeout 600 python tools/readme_bench.py 100000 > $O/an_readme_ctable$c.json 2>> $O/an.err
  python - <<P
import json
s=json.loads(open('gpurun_out/an_readme_ctable$c.json').read().strip().splitlines()[-1])
print("RB_CTABLE=$c s_per_read %.3g driver_wall %.2f table_setup %.3f tables %s"%(s['s_per_read'], s['driver_wall_s'], s['table_setup_s'], [t.split()[2]+t.split()[3] for t in s['kmer_tables']]))
P
done
tail -n 3 $O/an.err
