#!/usr/bin/env python3
"""First-call vs repeated-call time of the GPU build (column build and RED.OR) in one fresh process."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import readbouncer_b200 as rb
from readbouncer_b200 import synth
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2_100x4Mb_100bins"
w = bench.WORKLOADS[wl]
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream()
plan = synth.build_plan(bench.make_reference(w), w["fragment"], w["k"])
d_ref = torch.from_numpy(plan["bases"]).to(dev)
d_fb = torch.from_numpy(plan["frag_begin"].astype(np.int64)).to(dev)
d_fe = torch.from_numpy(plan["frag_end"].astype(np.int64)).to(dev)
d_fbin = torch.from_numpy(plan["frag_bin"].astype(np.int64)).to(dev)
n_frags = len(plan["frag_bin"]); max_frag = int((plan["frag_end"] - plan["frag_begin"]).max())
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
order = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "2,1").split(",")]
for variant in order:
    rb.set_insert_kernel(variant)
    for fresh in range(2):
        gf = rb.IBF.create(plan["n_bins"], 3, w["k"], plan["n_bits"])
        ms = []
        for rep in range(3):
            torch.cuda.synchronize(); ev0.record(stream)
            gf.insert_batch_dev(d_ref, d_fb, d_fe, d_fbin, n_frags, max_frag, stream=stream)
            ev1.record(stream); torch.cuda.synchronize(); ms.append(ev0.elapsed_time(ev1))
        print(json.dumps({"workload": wl, "variant": {1: "red_or", 2: "column"}[variant], "handle": fresh, "ms_per_call": ms}), flush=True)
        gf.close()
