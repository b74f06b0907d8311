#!/usr/bin/env python3
"""Latency of one host-buffer classify call (rb_ibf_count_batch, pinned buffers) for live-sized batches."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import readbouncer_b200 as rb
from readbouncer_b200 import synth
torch.cuda.set_device(0)
ref = [synth.random_bases(400_000, 2 + i) for i in range(100)]
plan = synth.build_plan(ref, 410_000, 13)
gf = rb.IBF.create(plan["n_bins"], 3, 13, plan["n_bits"], device=0)
gf.insert_batch(plan["bases"], plan["frag_begin"], plan["frag_end"], plan["frag_bin"])
gf.enable_kmer_table(0)
luts = np.stack([rb.threshold_lut(0.1, 13), rb.threshold_lut(0.08, 13)])
P = rb.capi._np_ptr
pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
for n in (1, 64, 512, 4096, 32768, 262144):
    bases, off, _ = synth.sample_reads(plan["bases"], n, 250, seed=5)
    hb, ho = pin(bases), pin(off.astype(np.uint64))
    r_max = torch.empty(2 * n, dtype=torch.int16, pin_memory=True).numpy().view(np.uint16)
    r_hit = torch.empty(2 * n, dtype=torch.uint8, pin_memory=True).numpy()
    r_am = torch.empty(2 * n, dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
    r_flag = torch.empty(n, dtype=torch.uint8, pin_memory=True).numpy()
    ts = []
    for it in range(60):
        t0 = time.perf_counter()
        rb.capi._check(rb.lib().rb_ibf_count_batch(gf._h, P(hb), P(ho), n, P(luts), 2, None, None, P(r_max), P(r_hit), P(r_am), P(r_flag), None))
        ts.append(time.perf_counter() - t0)
    ts = sorted(ts[10:])
    print(json.dumps({"reads": n, "table_span": gf.kmer_table_span(), "median_us": 1e6 * ts[len(ts) // 2], "p90_us": 1e6 * ts[int(len(ts) * 0.9)],
                      "reads_per_s_at_median": n / ts[len(ts) // 2]}), flush=True)
