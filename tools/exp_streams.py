"""Experiment: one count launch over the whole batch vs the same batch cut into pieces on concurrent streams
(device-resident; BASELINE config #2 shape at k = 13 / 15 / 17).  Prints one JSON line per variant."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import readbouncer_b200 as rb
from readbouncer_b200 import synth

dev = torch.device("cuda", 0)
for k in (13, 15, 17):
    ref = synth.HashReference([3_999_999] * 100, 2)
    plan = ref.plan(4_200_000, k)
    gf = rb.IBF.create(plan["n_bins"], 3, k, plan["n_bits"])
    d_ref = ref.to_device(dev)
    d = [torch.from_numpy(plan[key].astype(np.int64)).to(dev) for key in ("frag_begin", "frag_end", "frag_bin")]
    gf.insert_batch_dev(d_ref, d[0], d[1], d[2], 100, 4_000_000)
    torch.cuda.synchronize()
    del d_ref
    n = 1_000_000 if k < 17 else 262_144
    bases, off, _ = synth.sample_reads(ref, n, 250, seed=1234)
    d_bases = torch.from_numpy(bases).to(dev)
    d_off = torch.from_numpy(off.astype(np.int64)).to(dev)
    lut = torch.from_numpy(np.stack([rb.threshold_lut(0.1, k), rb.threshold_lut(0.08, k)]).view(np.int16)).to(dev)
    try:
        gf.enable_kmer_table(0)
    except rb.RBError:
        pass
    for pieces in (1, 2, 4, 8, 16):
        streams = [torch.cuda.Stream() for _ in range(min(pieces, 4))]
        cut = [n * i // pieces for i in range(pieces + 1)]
        keys = [torch.zeros(2 * (cut[i + 1] - cut[i]), dtype=torch.int64, device=dev) for i in range(pieces)]
        offs = [d_off[cut[i]:cut[i + 1] + 1].contiguous() for i in range(pieces)]

        def run():
            for i in range(pieces):
                st = streams[i % len(streams)]
                gf.count_batch_dev(d_bases, offs[i], cut[i + 1] - cut[i], lut, 2, keys[i], max_read_len=250, stream=st)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        main = torch.cuda.current_stream()
        e0.record(main)
        for st in streams:
            st.wait_event(e0)
        reps = 10
        for _ in range(reps):
            run()
        evs = []
        for st in streams:
            ev = torch.cuda.Event()
            ev.record(st)
            main.wait_event(ev)
        e1.record(main)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(json.dumps({"k": k, "pieces": pieces, "streams": len(streams), "ms_per_batch": ms, "chunks_per_s": n / ms * 1e3,
                          "table_kind": gf.kmer_table_kind(), "span": gf.kmer_table_span()}), flush=True)
    gf.close()
    torch.cuda.empty_cache()
