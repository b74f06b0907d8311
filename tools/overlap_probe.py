#!/usr/bin/env python3
"""Do count kernels of small batches overlap across streams?  64 launches of n reads each on 1 vs 4 streams."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import readbouncer_b200 as rb
from readbouncer_b200 import synth
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
ref = [synth.random_bases(200_000, 2 + i) for i in range(100)]
plan = synth.build_plan(ref, 210_000, 13)
gf = rb.IBF.create(plan["n_bins"], 3, 13, plan["n_bits"], device=0)
gf.insert_batch(plan["bases"], plan["frag_begin"], plan["frag_end"], plan["frag_bin"])
lut = torch.from_numpy(np.stack([rb.threshold_lut(0.1, 13)]).view(np.int16)).to(dev)
gf.enable_kmer_table(0)
print("table span", gf.kmer_table_span(), file=sys.stderr)
for n in (8192, 16384, 32768, 65536):
    bases, off, _ = synth.sample_reads(plan["bases"], n, 250, seed=5)
    sets = []
    for i in range(4):
        sets.append((torch.from_numpy(bases).to(dev), torch.from_numpy(off.astype(np.int64)).to(dev),
                     torch.zeros(n, dtype=torch.int64, device=dev)))
    streams = [torch.cuda.Stream() for _ in range(4)]
    out = {"reads": n}
    for ns in (1, 2, 4):
        for rep in range(2):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(torch.cuda.current_stream())
            for s in streams[:ns]:
                s.wait_event(e0)
            for i in range(64):
                s = streams[i % ns]; b, o, k = sets[i % ns]
                gf.count_batch_dev(b, o, n, lut, 1, k, max_read_len=250, stream=s)
            for s in streams[:ns]:
                ev = torch.cuda.Event(); ev.record(s); torch.cuda.current_stream().wait_event(ev)
            e1.record(torch.cuda.current_stream()); torch.cuda.synchronize()
        out["ms_64_launches_%d_streams" % ns] = e0.elapsed_time(e1)
    out["ideal_ms_at_420M"] = 64 * n / 420e3
    print(json.dumps(out), flush=True)
