#!/usr/bin/env python3
"""Small workloads for compute-sanitizer over the kernels added late in round 2: the group-loaded k-mer table for rows of
3..32 words (bit-sliced and atomic counters, both accumulator widths), the postings list kernel with 2 / 4 / 8 lanes per
list, and the slots of one or two lines read by lane groups.  k = 8 keeps every table small; results are compared with the
CPU oracle.

  compute-sanitizer --tool memcheck  python tools/sanitize_lane_groups.py
  compute-sanitizer --tool racecheck python tools/sanitize_lane_groups.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle                                      # noqa: E402  (checker)
import readbouncer_b200 as rb                      # noqa: E402
from readbouncer_b200 import synth                 # noqa: E402


def check(gf, of, plan, k, what):
    lut = rb.threshold_lut(0.1, k)
    for lengths in ([250] * 6 + [0, 5, k, 31, 100, 200 + k], [250] * 3 + [255 + k, 700, 1300]):
        b, o = synth.ragged_reads(plan["bases"], lengths, seed=21, frac_from_ref=0.7, n_frac=0.004)
        exp = of.count_batch(b, o, lut, n_threads=4)
        for dense in (True, False):
            got = gf.count_batch(b, o, lut, dense=dense)
            for key in ("max_count", "hit", "argmax_bin") + (("counts_fwd", "counts_rev") if dense else ()):
                assert np.array_equal(got[key], exp[key]), (what, lengths[-1], dense, key)


def main():
    k, n_hash = 8, 3
    for n_bins in (200, 300, 700, 1100):                       # 4, 5, 11, 18 row words
        ref = [synth.random_bases(600, 300 + i) for i in range(n_bins)]
        plan = synth.build_plan(ref, 1000, k, n_hash=n_hash)
        of = oracle.OracleIBF.create(n_bins, n_hash, k, plan["n_bits"])
        of.insert_batch(plan["bases"], plan["frag_begin"], plan["frag_end"], plan["frag_bin"], n_threads=4)
        gf = rb.IBF.create(n_bins, n_hash, k, plan["n_bits"])
        gf.insert_batch(plan["bases"], plan["frag_begin"], plan["frag_end"], plan["frag_bin"])
        # group-loaded table, bit-sliced and atomic counters
        os.environ["RB_CTABLE_WIDE"] = "1"
        for atomic in ("0", "1"):
            os.environ["RB_CTABLE_ATOMIC"] = atomic
            gf.enable_kmer_table(0)
            assert gf.kmer_table_kind() == 4
            check(gf, of, plan, k, ("ctable", n_bins, atomic))
        del os.environ["RB_CTABLE_ATOMIC"], os.environ["RB_CTABLE_WIDE"]
        if n_bins > 256:
            os.environ["RB_CTABLE"] = "0"
            os.environ["RB_POSTINGS_LAYOUT"] = "lists"
            gf.enable_kmer_table(0)
            assert gf.kmer_table_kind() == 2
            for sub in ("0", "2", "4", "8"):
                os.environ["RB_POSTINGS_SUB"] = sub
                check(gf, of, plan, k, ("lists", n_bins, sub))
            del os.environ["RB_POSTINGS_SUB"]
            os.environ["RB_POSTINGS_LAYOUT"] = "slots"
            for sb in ("128", "256"):
                os.environ["RB_SLOT_BYTES"] = sb
                gf.enable_kmer_table(0)
                assert gf.kmer_table_kind() == 3
                check(gf, of, plan, k, ("slots", n_bins, sb))
            del os.environ["RB_SLOT_BYTES"], os.environ["RB_POSTINGS_LAYOUT"], os.environ["RB_CTABLE"]
        gf.close()
    print("sanitize_lane_groups ok")


if __name__ == "__main__":
    main()
