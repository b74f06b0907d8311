#!/usr/bin/env python3
"""Random-gather ceiling with warp-cooperative row loads: G adjacent lanes read one row in ONE load
instruction (16 or 32 bytes per lane).  One JSON line per point; compare with gather_sweep2.py
(one thread reads the whole row with successive 16-byte loads)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import readbouncer_b200 as rb
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
stream = torch.cuda.current_stream()
sink = torch.zeros(1, dtype=torch.int64, device=dev)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for mb in (2147, 17180, 68000):
    buf = torch.zeros(mb * 1000 * 1000 // 8, dtype=torch.int64, device=dev)
    for row_bytes, lane_bytes in ((16, 16), (32, 16), (32, 32), (64, 16), (64, 32), (128, 16), (128, 32), (256, 16), (256, 32)):
        n_rows = buf.numel() * 8 // row_bytes
        g = row_bytes // lane_bytes
        for blocks_per_sm in (8,):
            blocks = 148 * blocks_per_sm
            ppg = 64 * g                      # same number of load instructions per thread for every shape
            best = 1e30
            for it in range(3):
                torch.cuda.synchronize(); ev0.record(stream)
                rb.microbench_gather_coop(buf, n_rows, row_bytes, lane_bytes, ppg, blocks, sink, stream=stream)
                ev1.record(stream); torch.cuda.synchronize()
                if it: best = min(best, ev0.elapsed_time(ev1))
            probes = blocks * 256 // g * ppg
            print(json.dumps({"footprint_MB": mb, "row_bytes": row_bytes, "lane_bytes": lane_bytes, "ms": best,
                              "Grows_per_s": probes / best / 1e6, "GBps_useful": probes * row_bytes / best / 1e6}), flush=True)
    del buf
