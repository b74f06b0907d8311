#!/usr/bin/env python3
"""Small postings-table workload for compute-sanitizer (memcheck / racecheck): 1 100 bins, k = 9, an over-full filter so that
lists have ~160..610 ids (full rounds, further rounds, every tail width), short and long reads (8- and 16-bit counters),
dense and key-only outputs; results are compared with the CPU oracle.

  compute-sanitizer --tool memcheck  python tools/sanitize_postings.py
  compute-sanitizer --tool racecheck python tools/sanitize_postings.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle                                      # noqa: E402  (checker)
import readbouncer_b200 as rb                      # noqa: E402
from readbouncer_b200 import synth                 # noqa: E402


def main():
    os.environ["RB_CTABLE"] = "0"          # rows of <= 32 words would take the group-loaded k-mer table

    k, n_hash = 9, 3
    ref = [synth.random_bases(1500, 300 + i) for i in range(1100)]
    plan = synth.build_plan(ref, 2000, k, n_hash=n_hash)
    lut = rb.threshold_lut(0.1, k)
    for n_blocks in (2600, 6000):
        n_bits = n_blocks * 64 * 18
        of = oracle.OracleIBF.create(1100, n_hash, k, n_bits)
        of.insert_batch(plan["bases"], plan["frag_begin"], plan["frag_end"], plan["frag_bin"], n_threads=4)
        gf = rb.IBF.create(1100, n_hash, k, n_bits)
        gf.insert_batch(plan["bases"], plan["frag_begin"], plan["frag_end"], plan["frag_bin"])
        gf.enable_kmer_table(0)
        assert gf.kmer_table_kind() == 2
        for lengths in ([250] * 6 + [0, 5, k, 31, 100, 254 + k], [250] * 3 + [255 + k, 700]):
            b, o = synth.ragged_reads(plan["bases"], lengths, seed=21, frac_from_ref=0.7, n_frac=0.004)
            exp = of.count_batch(b, o, lut, n_threads=4)
            for dense in (True, False):
                got = gf.count_batch(b, o, lut, dense=dense)
                for key in ("max_count", "hit", "argmax_bin") + (("counts_fwd", "counts_rev") if dense else ()):
                    assert np.array_equal(got[key], exp[key]), (n_blocks, lengths[-1], dense, key)
        gf.close()
    print("sanitize_postings ok")


if __name__ == "__main__":
    main()
