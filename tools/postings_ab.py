#!/usr/bin/env python3
"""cfg #3 (3.1 Gb, 31 008 bins) postings lookup: ids of a list ascending (RB_POSTINGS_ORDER=0) vs dealt over the
ATOMS groups by shared-memory bank (default), same process, same filter, same reads.  One JSON line per variant;
the packed keys of the two variants must be equal."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import readbouncer_b200 as rb                      # noqa: E402
from readbouncer_b200 import synth                 # noqa: E402
import bench                                       # noqa: E402


def main():
    os.environ["RB_CTABLE"] = "0"          # rows of <= 32 words would take the group-loaded k-mer table

    w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg3_3.1Gb_31kbins"]
    n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else w["reads"]
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    t0 = time.time()
    plan = synth.build_plan(bench.make_reference(w), w["fragment"], w["k"])
    gf = rb.IBF.create(plan["n_bins"], 3, w["k"], plan["n_bits"], device=0)
    gf.insert_batch(plan["bases"], plan["frag_begin"], plan["frag_end"], plan["frag_bin"])
    bases, off, _ = synth.sample_reads(plan["bases"], n_reads, w["chunk"], seed=1234)
    d_bases = torch.from_numpy(bases).to(dev)
    d_off = torch.from_numpy(off.astype(np.int64)).to(dev)
    luts = np.stack([rb.threshold_lut(0.1, w["k"]), rb.threshold_lut(0.08, w["k"])])
    d_lut = torch.from_numpy(luts.view(np.int16)).to(dev)
    setup_s = time.time() - t0
    keys = {}
    for order in ("0", "1", "0", "1"):
        os.environ["RB_POSTINGS_ORDER"] = order
        gf.disable_kmer_table()
        torch.cuda.synchronize()
        t0 = time.time()
        gf.enable_kmer_table(0)
        torch.cuda.synchronize()
        build_s = time.time() - t0
        assert gf.kmer_table_kind() == 2
        d_keys = torch.zeros(2 * n_reads, dtype=torch.int64, device=dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for it in range(8):
            if it == 3:
                torch.cuda.synchronize()
                ev[0].record(stream)
            gf.count_batch_dev(d_bases, d_off, n_reads, d_lut, 2, d_keys, max_read_len=w["chunk"], stream=stream)
        ev[1].record(stream)
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 5
        keys[order] = d_keys.cpu()
        print(json.dumps({"order": "dealt" if order == "1" else "ascending", "ms_per_launch": ms, "chunks": n_reads,
                          "chunks_per_s": n_reads / ms * 1e3, "table_bytes": gf.kmer_table_bytes(), "table_build_s": build_s,
                          "setup_s": setup_s}), flush=True)
    assert torch.equal(keys["0"], keys["1"]), "the order of a list changed a result"
    print("keys equal")


if __name__ == "__main__":
    main()
