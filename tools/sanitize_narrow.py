#!/usr/bin/env python3
"""Small workloads for compute-sanitizer over the narrow-filter kernels in their final round-2 form (the hashed path of the
window-table kernels reads its filter view from shared memory): window tables of span 1..4 through the group-per-read and the
warp-per-read kernel, the packed-plane host path (dense = False), the hashed-probe and the streaming kernel, one and two
threshold tables, 1- and 2-word rows.  k = 8 keeps every table small; results are compared with the CPU oracle.

  compute-sanitizer --tool memcheck  python tools/sanitize_narrow.py [quick]
  compute-sanitizer --tool racecheck python tools/sanitize_narrow.py [quick]
(quick: 2-word rows only, window tables of span 2 and 3 through both kernels + the hashed-probe kernel)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle                                      # noqa: E402  (checker)
import readbouncer_b200 as rb                      # noqa: E402
from readbouncer_b200 import synth                 # noqa: E402


def same(got, exp, t, dense, what):
    keys = ("max_count", "hit", "argmax_bin") + (("counts_fwd", "counts_rev") if dense and t == 0 else ())
    for key in keys:
        g = got[key][t] if key in ("max_count", "hit", "argmax_bin") and got[key].ndim == 2 else got[key]
        assert np.array_equal(g, exp[key]), (what, key)


def main():
    k, n_hash = 8, 3
    luts = np.stack([rb.threshold_lut(0.1, k), rb.threshold_lut(0.08, k)])
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    plan_kernels = ((3, (2, 3)), (5, (3,)), (1, (0,))) if quick else ((3, (1, 2, 3, 4)), (5, (2, 3)), (4, (1,)), (1, (0,)), (2, (0,)))
    for n_bins in ((100,) if quick else (40, 100)):            # 1 and 2 row words
        ref = [synth.random_bases(900, 700 + i) for i in range(n_bins)]
        plan = synth.build_plan(ref, 1000, k, n_hash=n_hash)
        of = oracle.OracleIBF.create(n_bins, n_hash, k, plan["n_bits"])
        of.insert_batch(plan["bases"], plan["frag_begin"], plan["frag_end"], plan["frag_bin"], n_threads=4)
        gf = rb.IBF.create(n_bins, n_hash, k, plan["n_bits"])
        gf.insert_batch(plan["bases"], plan["frag_begin"], plan["frag_end"], plan["frag_bin"])
        assert np.array_equal(gf.download(), of.words()[:plan["n_bits"] // 64])
        batches = []
        for lengths in ([250] * 10 + [0, 1, k - 1, k, k + 2, 31, 33, 64, 100, 249, 251, 380],      # group-per-read range
                        [250] * 3 + [600, 1300]):                                                   # warp-per-read
            b, o = synth.ragged_reads(plan["bases"], lengths, seed=5 + len(lengths), frac_from_ref=0.7, n_frac=0.01, lower_frac=0.1)
            b[int(o[1]):int(o[2])] = ord("N")                 # an all-N read: every window takes the hashed path
            batches.append((b, o, [of.count_batch(b, o, luts[t], n_threads=4) for t in range(2)]))
        for which, spans in plan_kernels:
            rb.set_count_kernel(which)
            for span in spans:
                if span:
                    os.environ["RB_KMER_TABLE_SPAN"] = str(span)
                    try:
                        gf.enable_kmer_table(0)
                    finally:
                        del os.environ["RB_KMER_TABLE_SPAN"]
                    assert gf.kmer_table_span() == span, (which, span, gf.kmer_table_span())
                for b, o, exp in batches:
                    for dense in (True, False):               # dense = False: packed bit planes from the host threads
                        got = gf.count_batch(b, o, luts, dense=dense)
                        for t in range(2):
                            same(got, exp[t], t, dense, (n_bins, which, span, dense, t))
        rb.set_count_kernel(0)
        gf.close()
    print("sanitize_narrow ok, %d kernel launches" % rb.kernel_launches())


if __name__ == "__main__":
    main()
