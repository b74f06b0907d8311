#!/usr/bin/env python3
"""Pinned host<->device copy bandwidth of this box (the bound of the ASCII host-buffer path)."""
import json, torch
dev = torch.device("cuda", 0)
out = {}
for mb in (8, 64, 256):
    h = torch.empty(mb << 20, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(mb << 20, dtype=torch.uint8, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, (dst, src) in (("h2d", (d, h)), ("d2h", (h, d))):
        best = 1e9
        for _ in range(5):
            torch.cuda.synchronize(); e0.record(); dst.copy_(src, non_blocking=True); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out["%s_%dMB_GBps" % (name, mb)] = (mb << 20) / best / 1e6
print(json.dumps(out))
