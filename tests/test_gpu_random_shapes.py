"""Seeded random filter and batch shapes through the default (auto) paths of the C ABI against the oracle: bins 1 .. 2 500,
k 8 .. 20, 2 .. 4 hash functions, fragment sizes that make short and long postings lists, ragged reads with N / IUPAC / lower
case, one and two threshold tables, dense and summary-only outputs, tables built or not.  Whatever kernel the library picks for a
shape -- window table, k-mer table, postings, streaming, hashed probes -- the answers must be the oracle's."""
import os

import numpy as np
import pytest

import oracle
import readbouncer_b200 as rb
from readbouncer_b200 import synth

pytestmark = pytest.mark.gpu

# RB_FUZZ_BASE=<int> draws another family of shapes (default 1000: the committed, driver-run family)
FUZZ_BASE = int(os.environ.get("RB_FUZZ_BASE", "1000"))


@pytest.mark.parametrize("case", range(24))
def test_random_shape_equals_oracle(case):
    rng = np.random.default_rng(FUZZ_BASE + case)
    k = int(rng.choice([8, 10, 11, 12, 13, 13, 14, 15, 16, 17, 20]))
    n_hash = int(rng.choice([2, 3, 3, 3, 4]))
    n_seqs = int(rng.choice([1, 3, 40, 64, 65, 130, 257, 600, 2500]))
    seq_len = int(rng.integers(max(k + 5, 200), 6000))
    frag = int(rng.integers(max(k + 40, 300), 8000))
    while seq_len % frag > frag - k + 2:                  # stay out of the quirk-Q3 window (covered by its own test)
        seq_len -= 7
    ref = [synth.random_bases(seq_len, 5000 * case + i + 7 * (FUZZ_BASE - 1000)) for i in range(n_seqs)]
    if case % 5 == 0:                                      # a reference with N runs: cutOutNNNs on the way in
        for s in ref[: max(1, n_seqs // 3)]:
            a = int(rng.integers(0, max(1, len(s) - 60)))
            s[a:a + int(rng.integers(1, 50))] = ord("N")
    plan = synth.build_plan(ref, frag, k, n_hash=n_hash)
    if plan["bin_ids_consumed"] != plan["n_bins"]:
        pytest.skip("quirk Q3 shape")
    of = oracle.OracleIBF.create(plan["n_bins"], n_hash, k, plan["n_bits"])
    of.insert_batch(plan["bases"], plan["frag_begin"], plan["frag_end"], plan["frag_bin"], n_threads=4)
    gf = rb.IBF.create(plan["n_bins"], n_hash, k, plan["n_bits"])
    gf.insert_batch(plan["bases"], plan["frag_begin"], plan["frag_end"], plan["frag_bin"])
    assert np.array_equal(gf.download(), of.words()[:plan["n_bits"] // 64])
    lengths = [int(x) for x in rng.choice([0, 1, k - 1, k, k + 1, 60, 249, 250, 251, 254 + k, 255 + k, 600, 1300, 4000], size=180)]
    lengths += [250] * 60
    bases, off = synth.ragged_reads(plan["bases"], lengths, seed=77 + case, frac_from_ref=0.6, error_rate=0.06,
                                    n_frac=0.002 * (case % 3), lower_frac=0.05 * (case % 2))
    luts = np.stack([rb.threshold_lut(0.1, k), rb.threshold_lut(0.07, k)])
    exp = [of.count_batch(bases, off, luts[t], n_threads=4) for t in range(2)]
    for tables in (False, True):
        if tables:
            try:
                gf.enable_kmer_table(0)
            except rb.RBError:
                break                                      # no table applies to this shape: the first pass was the test
        else:
            gf.disable_kmer_table()
        got = gf.count_batch(bases, off, luts, dense=True)
        assert np.array_equal(got["counts_fwd"], exp[0]["counts_fwd"]) and np.array_equal(got["counts_rev"], exp[0]["counts_rev"])
        assert np.array_equal(got["read_flag"], exp[0]["short_read"])
        for t in range(2):
            for key in ("max_count", "hit", "argmax_bin"):
                assert np.array_equal(got[key][t], exp[t][key]), (tables, t, key, gf.kmer_table_kind(), gf.kmer_table_span())
        one = gf.count_batch(bases, off, luts[1])          # summary only, one table
        for key in ("max_count", "hit", "argmax_bin"):
            assert np.array_equal(one[key], exp[1][key]), (tables, key)
