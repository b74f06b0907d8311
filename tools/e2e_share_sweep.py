#!/usr/bin/env python3
"""Host-buffer call (rb_ibf_count_batch, pinned buffers) on BASELINE config #2: sweep of the share of pieces that cross
PCIe as ASCII while the host threads pack the others (RB_ASCII_SHARE) and of the piece size (RB_PIECE_MB)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import readbouncer_b200 as rb
from readbouncer_b200 import synth
torch.cuda.set_device(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ref = [synth.random_bases(4_000_000, 2 + i) for i in range(100)]
plan = synth.build_plan(ref, 4_200_000, 13)
del ref
gf = rb.IBF.create(plan["n_bins"], 3, 13, plan["n_bits"], device=0)
gf.insert_batch(plan["bases"], plan["frag_begin"], plan["frag_end"], plan["frag_bin"])
luts = np.stack([rb.threshold_lut(0.1, 13), rb.threshold_lut(0.08, 13)])
P = rb.capi._np_ptr
pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
bases, off, _ = synth.sample_reads(plan["bases"], n, 250, seed=1234)
hb, ho = pin(bases), pin(off.astype(np.uint64))
r_max = torch.empty(2 * n, dtype=torch.int16, pin_memory=True).numpy().view(np.uint16)
r_hit = torch.empty(2 * n, dtype=torch.uint8, pin_memory=True).numpy()
r_am = torch.empty(2 * n, dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
r_flag = torch.empty(n, dtype=torch.uint8, pin_memory=True).numpy()
call = lambda: rb.capi._check(rb.lib().rb_ibf_count_batch(gf._h, P(hb), P(ho), n, P(luts), 2, None, None, P(r_max), P(r_hit), P(r_am), P(r_flag), None))
call()
ref_max = r_max.copy()
pieces = os.environ.get("SWEEP_PIECES", "8,4,16").split(",")
shares = os.environ.get("SWEEP_SHARES", "0,0.1,0.15,0.2,0.25,0.3,0.4,0.5,1").split(",")
for piece in pieces:
    os.environ["RB_PIECE_MB"] = piece
    for share in shares:
        os.environ["RB_ASCII_SHARE"] = share
        for _ in range(3):
            call()
        x0 = rb.transfer_bytes()
        ts = []
        for _ in range(int(os.environ.get('SWEEP_CALLS', '12'))):
            t0 = time.perf_counter(); call(); ts.append(time.perf_counter() - t0)
        x1 = rb.transfer_bytes()
        assert np.array_equal(r_max, ref_max)
        ts.sort()
        print(json.dumps({"piece_mb": int(piece), "ascii_share": float(share), "reads": n, "median_ms": 1e3 * ts[len(ts) // 2], "min_ms": 1e3 * ts[0],
                          "mean_ms": 1e3 * sum(ts) / len(ts), "chunks_per_s_mean": n * len(ts) / sum(ts), "h2d_bytes": (x1[0] - x0[0]) // max(1, len(ts))}), flush=True)
