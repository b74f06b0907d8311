"""Size-independent properties of the IBF path, checked on the CPU oracle (the GPU suite checks the same properties
of the CUDA path at full size): strand symmetry, OR-linearity of the build, idempotence, prefix monotonicity,
locality of a non-ACGT base, the independent definition of the hash, and the threshold wrap."""
import numpy as np
import pytest

import oracle

K, H = 13, 3
COMP = {ord("A"): ord("T"), ord("C"): ord("G"), ord("G"): ord("C"), ord("T"): ord("A"), ord("N"): ord("N")}


def rand_bases(n, seed):
    return np.frombuffer(b"ACGT", np.uint8)[np.random.default_rng(seed).integers(0, 4, size=n)]


def revcomp(a):
    return np.array([COMP[int(c)] for c in a[::-1]], np.uint8)


def small_filter(n_bins=70, frag=3000, seed=1, n_bits=None):
    seqs = [rand_bases(frag, seed + i) for i in range(n_bins)]
    n_bits = n_bits or oracle.filter_size_bits(frag, K, H, 0.01, n_bins)
    f = oracle.OracleIBF.create(n_bins, H, K, n_bits)
    for b, s in enumerate(seqs):
        f.insert(s.tobytes(), b)
    return f, seqs, n_bits


def test_reverse_strand_count_is_forward_count_of_the_reverse_complement():
    f, seqs, _ = small_filter()
    for i in range(5):
        read = seqs[3 * i][100:350].copy()
        read[::17] = ord("A")                                   # a few errors
        assert np.array_equal(f.count(read.tobytes(), revcomp=True), f.count(revcomp(read).tobytes()))
        assert np.array_equal(f.count(revcomp(read).tobytes(), revcomp=True), f.count(read.tobytes()))


def test_build_is_or_linear_and_idempotent():
    fa, seqs, n_bits = small_filter(seed=10)
    fb, seqs_b, _ = small_filter(seed=500, n_bits=n_bits)
    both = oracle.OracleIBF.create(70, H, K, n_bits)
    for b in range(70):
        both.insert(seqs[b].tobytes(), b)
        both.insert(seqs_b[b].tobytes(), b)
    n = n_bits // 64
    assert np.array_equal(both.words()[:n], fa.words()[:n] | fb.words()[:n])
    before = both.words()[:n].copy()
    for b in range(70):
        both.insert(seqs[b].tobytes(), b)                      # inserting again sets no new bit
    assert np.array_equal(both.words()[:n], before)


def test_counts_grow_with_the_prefix_and_never_exceed_the_positions():
    f, seqs, _ = small_filter()
    read = seqs[7][500:1000]
    prev = np.zeros(70, np.uint16)
    for length in (K - 1, K, 40, 250, 500):
        c = f.count(read[:length].tobytes())
        assert np.all(c >= prev) and c.max() <= max(0, length - K + 1)
        prev = c
    assert f.count(read.tobytes())[7] == 500 - K + 1            # every k-mer of an inserted stretch is found in its bin


def test_a_non_acgt_base_only_touches_the_windows_that_cover_it():
    f, seqs, _ = small_filter()
    read = seqs[11][200:450].copy()
    base = f.count(read.tobytes()).astype(np.int64)
    hit = read.copy()
    hit[100] = ord("N")
    c = f.count(hit.tobytes()).astype(np.int64)
    assert np.all(np.abs(base - c) <= K)                        # at most K windows change
    assert base[11] - c[11] in range(0, K + 1)
    lower = np.frombuffer(read.tobytes().lower(), np.uint8)
    assert np.array_equal(f.count(lower.tobytes()).astype(np.int64), base)       # case is ignored (SeqAn Dna5 translate table)


def test_hash_is_the_base5_polynomial_and_rows_follow_the_multiply_shift():
    rng = np.random.default_rng(3)
    f = oracle.OracleIBF.create(64, H, K, 64 * 1000003)
    for _ in range(50):
        kmer = rand_bases(K, int(rng.integers(1 << 30)))
        rank = {ord("A"): 0, ord("C"): 1, ord("G"): 2, ord("T"): 3}
        h = 0
        for c in kmer:
            h = (h * 5 + rank[int(c)]) % (1 << 64)
        assert oracle.kmer_hash(kmer.tobytes(), K) == h
        rows = [f.hash_row(h, i) for i in range(H)]
        assert all(0 <= r < 1000003 for r in rows) and len(set(rows)) >= 2


def test_threshold_wraps_like_the_reference_and_lut_matches_scalar():
    lut = oracle.threshold_lut(0.1, K)
    for length in (K, 50, 100, 250, 360, 1500, 65535):
        assert lut[length] == oracle.threshold(0.1, K, length)
    # short reads: readlen - k + 1 - ci.high is negative as int16 and becomes >= 32768 as uint16 -> nothing can match
    assert oracle.threshold(0.1, K, 35) >= 32768
    assert 0 < oracle.threshold(0.1, K, 250) < 250 - K + 1
    # more tolerated errors -> lower threshold ... until the interval passes the number of positions and the wrap strikes
    assert oracle.threshold(0.1, K, 250) <= oracle.threshold(0.08, K, 250) <= oracle.threshold(0.05, K, 250) < 250 - K + 1
    assert oracle.threshold(0.15, K, 250) >= 32768


@pytest.mark.parametrize("n_threads", [1, 4])
def test_batch_call_equals_per_read_calls_on_ragged_input(n_threads):
    f, seqs, _ = small_filter()
    lut = oracle.threshold_lut(0.1, K)
    lengths = [0, 1, K - 1, K, 31, 250, 251, 700]
    parts = [seqs[i % 70][50:50 + n] for i, n in enumerate(lengths)]
    off = np.zeros(len(parts) + 1, np.uint64)
    off[1:] = np.cumsum([len(p) for p in parts])
    bases = np.concatenate(parts).astype(np.uint8)
    got = f.count_batch(bases, off, lut, dense=True, n_threads=n_threads)
    for i, p in enumerate(parts):
        if len(p) >= K:
            assert np.array_equal(got["counts_fwd"][i], f.count(p.tobytes()))
            assert np.array_equal(got["counts_rev"][i], f.count(p.tobytes(), revcomp=True))
            m = np.maximum(got["counts_fwd"][i], got["counts_rev"][i])
            thr = int(lut[len(p)])
            assert got["hit"][i] == int(m.max() >= thr)
            if got["hit"][i]:
                assert got["max_count"][i] == m.max() and got["argmax_bin"][i] == int(np.argmax(m))
        else:
            assert got["short_read"][i] == 1 and got["hit"][i] == 0
