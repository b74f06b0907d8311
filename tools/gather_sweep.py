#!/usr/bin/env python3
"""Random-gather ceiling sweep (rb_microbench_gather): footprint x row bytes x L2 fetch granularity.
Run on the GPU box; prints one JSON line per point.  Used to establish the random-sector roofline."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import readbouncer_b200 as rb

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
stream = torch.cuda.current_stream()
sink = torch.zeros(1, dtype=torch.int64, device=dev)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for gran in (128, 64, 32):
    rb.set_l2_fetch_granularity(gran)
    got = rb.get_l2_fetch_granularity()
    for mb in (32, 96, 256, 831, 4096):
        buf = torch.zeros(mb * 1000 * 1000 // 8, dtype=torch.int64, device=dev)
        for row_bytes in (8, 16, 32):
            n_rows = buf.numel() * 8 // row_bytes
            for blocks_per_sm in (4, 8):
                blocks, ppt = 148 * blocks_per_sm, 512
                best = 1e30
                for it in range(4):
                    torch.cuda.synchronize()
                    ev0.record(stream)
                    rb.microbench_gather(buf, n_rows, row_bytes, ppt, blocks, sink, stream=stream)
                    ev1.record(stream)
                    torch.cuda.synchronize()
                    if it:
                        best = min(best, ev0.elapsed_time(ev1))
                probes = blocks * 256 * ppt
                print(json.dumps({"l2_gran_set": gran, "l2_gran_get": got, "footprint_MB": mb, "row_bytes": row_bytes,
                                  "blocks_per_sm": blocks_per_sm, "ms": best, "Gprobes_per_s": probes / best / 1e6,
                                  "GBps_useful": probes * row_bytes / best / 1e6}))
        del buf
