#!/usr/bin/env python3
"""Random-gather ceiling vs footprint (TLB reach) and row size.  One JSON line per point."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import readbouncer_b200 as rb
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
stream = torch.cuda.current_stream()
sink = torch.zeros(1, dtype=torch.int64, device=dev)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for mb in (256, 1024, 2147, 4096, 8192, 17180, 34360, 68000):
    buf = torch.zeros(mb * 1000 * 1000 // 8, dtype=torch.int64, device=dev)
    for row_bytes in (32, 64, 128):
        n_rows = buf.numel() * 8 // row_bytes
        for blocks_per_sm in (8,):
            blocks, ppt = 148 * blocks_per_sm, 256
            best = 1e30
            for it in range(3):
                torch.cuda.synchronize(); ev0.record(stream)
                rb.microbench_gather(buf, n_rows, row_bytes, ppt, blocks, sink, stream=stream)
                ev1.record(stream); torch.cuda.synchronize()
                if it: best = min(best, ev0.elapsed_time(ev1))
            probes = blocks * 256 * ppt
            print(json.dumps({"footprint_MB": mb, "row_bytes": row_bytes, "ms": best, "Gprobes_per_s": probes / best / 1e6,
                              "GBps_useful": probes * row_bytes / best / 1e6}), flush=True)
    del buf
