#!/usr/bin/env python3
"""Build-kernel timing (BASELINE config #4): the RED.OR insert kernel vs the column build on the same fragments,
device-resident inputs, CUDA events; the two matrices must be identical.  Also times RED.OR into matrices of
different sizes (L2-resident vs HBM) to show what bounds the direct kernel.

  python tools/build_bench.py [--workload cfg3_3.1Gb_31kbins] [--reps 3]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (workload table and reference generator)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3_3.1Gb_31kbins")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--red-sweep", action="store_true")
    args = ap.parse_args()
    import torch
    import readbouncer_b200 as rb
    from readbouncer_b200 import synth
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    w = bench.WORKLOADS[args.workload]
    plan = synth.build_plan(bench.make_reference(w), w["fragment"], w["k"])
    d_ref = torch.from_numpy(plan["bases"]).to(dev)
    d_fb = torch.from_numpy(plan["frag_begin"].astype(np.int64)).to(dev)
    d_fe = torch.from_numpy(plan["frag_end"].astype(np.int64)).to(dev)
    d_fbin = torch.from_numpy(plan["frag_bin"].astype(np.int64)).to(dev)
    n_frags = len(plan["frag_bin"])
    max_frag = int((plan["frag_end"] - plan["frag_begin"]).max())
    n_kmers = int(np.maximum(plan["frag_end"] - plan["frag_begin"], w["k"] - 1).sum() - (w["k"] - 1) * n_frags)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = {"workload": args.workload, "bins": plan["n_bins"], "filter_bytes": plan["n_bits"] // 8, "kmers": n_kmers,
           "fragments": n_frags}
    sums = {}
    for variant, name in ((1, "red_or"), (2, "column")):
        rb.set_insert_kernel(variant)
        best = None
        for rep in range(args.reps):
            gf = rb.IBF.create(plan["n_bins"], 3, w["k"], plan["n_bits"])
            torch.cuda.synchronize()
            l0 = rb.kernel_launches()
            ev0.record(stream)
            gf.insert_batch_dev(d_ref, d_fb, d_fe, d_fbin, n_frags, max_frag, stream=stream)
            ev1.record(stream)
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1)
            best = ms if best is None else min(best, ms)
            launches = rb.kernel_launches() - l0
            if rep == args.reps - 1:
                words = gf.download()
                sums[name] = (int(np.bitwise_xor.reduce(words)), int(words.view(np.uint32).astype(np.uint64).sum()),
                              int(np.unpackbits(words[:1 << 20].view(np.uint8)).sum()))
                del words
            gf.close()
        out[name] = {"ms": best, "kmers_per_s": n_kmers / (best * 1e-3), "launches": launches,
                     "bytes_written_algorithmic_per_s": n_kmers * 24 / (best * 1e-3)}
    out["identical"] = sums["red_or"] == sums["column"]
    out["speedup"] = out["red_or"]["ms"] / out["column"]["ms"]
    print(json.dumps(out))
    if args.red_sweep:
        # RED.OR into one-word-per-row matrices of growing size: 12 fragments of 4 M bases into bins 0..11
        rb.set_insert_kernel(1)
        nb = 48_000_000
        fb = (np.arange(12) * 4_000_000).astype(np.int64)
        t_fb, t_fe = torch.from_numpy(fb).to(dev), torch.from_numpy(fb + 4_000_000).to(dev)
        t_bin = torch.arange(12, dtype=torch.int64, device=dev)
        for rows in (1_236_269, 4_000_000, 16_000_000, 51_929_353, 200_000_000):
            gf = rb.IBF.create(64, 3, 13, rows * 64)
            for rep in range(2):
                torch.cuda.synchronize()
                ev0.record(stream)
                gf.insert_batch_dev(d_ref[:nb], t_fb, t_fe, t_bin, 12, 4_000_000, stream=stream)
                ev1.record(stream)
                torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1)
            print(json.dumps({"red_or_rows": rows, "matrix_mb": rows * 8 / 1e6, "ms": ms,
                              "g_red_per_s": 12 * (4_000_000 - 12) * 3 / (ms * 1e-3) / 1e9}))
            gf.close()


if __name__ == "__main__":
    main()
