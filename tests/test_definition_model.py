"""An INDEPENDENT second restatement of SURVEY Appendix A (A.1-A.6), written in plain Python integers straight from the
specification text and sharing no code with oracle/ibf_oracle.c, compared with the C oracle bit for bit.

Why: the reference's fixtures pin the oracle only for filters of 2 and 4 bins (binWidth 1).  Rows of several words
(binWidth > 1, every benchmarked configuration), N / IUPAC bases inside k-mers, k other than 13 / 15 and a bin count that is
not a multiple of 64 are "parity unpinned" corners.  Two implementations that were written independently from
the same specification and agree on those corners do not replace a reference fixture, but they rule out a slip of one of
them (word order inside a row, bit order inside a word, the reverse strand, the digit of N).

The second half does the same for the build-side host logic: statement-level Python models of IBF::cutOutNNNs and of the fragment
loop (IBFBuild.cpp:112-132, 165-202, quirks Q1-Q3) against the oracle AND the product's host entry points on random inputs; the
third restates calculateCI and the threshold wrap (IBF.hpp:268-338, IBFClassify.cpp:102-109) with Python's math module and checks
both threshold tables for 48 (error rate, k) pairs."""
import numpy as np
import pytest

import oracle
import readbouncer_b200 as rb

SEED = 0x90B45D39FB6DA1FA       # A.1
SHIFT = 27
M64 = (1 << 64) - 1
RANK = {c: r for r, cs in enumerate(("Aa", "Cc", "Gg", "Tt")) for c in cs}      # A.3: everything else -> 4
COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "a": "t", "c": "g", "g": "c", "t": "a"}


class ModelIBF:
    """The filter as a Python set of bit addresses."""

    def __init__(self, n_bins, n_hash, k, n_bits):
        self.n_bins, self.h, self.k, self.n_bits = n_bins, n_hash, k, n_bits
        self.bin_width = -(-n_bins // 64)
        self.block_bits = 64 * self.bin_width
        self.n_blocks = n_bits // self.block_bits
        self.pre = [i ^ ((k * SEED) & M64) for i in range(n_hash)]
        self.bits = set()

    def rows(self, kmer):
        v = 0
        for c in kmer:
            v = (v * 5 + RANK.get(c, 4)) & M64
        out = []
        for p in self.pre:
            x = (p * v) & M64
            x ^= x >> SHIFT
            out.append(x % self.n_blocks)
        return out

    def insert(self, text, b):
        for j in range(len(text) - self.k + 1):
            for r in self.rows(text[j:j + self.k]):
                self.bits.add(r * self.block_bits + b)

    def count(self, text):
        counts = [0] * self.n_bins
        for j in range(len(text) - self.k + 1):
            rows = self.rows(text[j:j + self.k])
            for b in range(self.n_bins):
                if all(r * self.block_bits + b in self.bits for r in rows):
                    counts[b] += 1
        return counts

    def words(self):
        w = np.zeros(self.n_bits // 64, np.uint64)
        for p in self.bits:
            w[p >> 6] |= np.uint64(1) << np.uint64(p & 63)
        return w


def revcomp(s):
    return "".join(COMP.get(c, "N") for c in reversed(s))


def rand_text(rng, n, alphabet="ACGT"):
    return "".join(alphabet[i] for i in rng.integers(0, len(alphabet), size=n))


@pytest.mark.parametrize("n_bins,k,rows", [(2, 13, 4001), (64, 13, 3001), (65, 13, 3001), (100, 13, 2003), (130, 11, 1999),
                                           (200, 15, 1501), (70, 17, 2500), (257, 9, 997)])
def test_c_oracle_equals_the_python_definition(n_bins, k, rows):
    rng = np.random.default_rng(1000 * n_bins + k)
    bin_width = -(-n_bins // 64)
    n_bits = rows * 64 * bin_width
    model = ModelIBF(n_bins, 3, k, n_bits)
    orc = oracle.OracleIBF.create(n_bins, 3, k, n_bits)
    texts = []
    for b in range(n_bins):
        t = rand_text(rng, int(rng.integers(k - 2, 70)))          # some texts shorter than k: nothing inserted (A.5)
        if b % 7 == 3 and len(t) > k:                              # N, lower case and IUPAC inside k-mers
            t = t[:5] + "N" + t[6:9].lower() + "R" + t[10:]
        texts.append(t)
        model.insert(t, b)
        orc.insert(t.encode(), b)
    # A.2 / A.5: the same bits at the same addresses
    assert np.array_equal(orc.words()[:n_bits // 64], model.words())
    # A.6 on both strands, reads made of inserted pieces (so that counts are not all zero), with errors and an N
    for i in range(4):
        pieces = [texts[int(rng.integers(n_bins))] for _ in range(4)]
        read = "".join(pieces)[:150]
        if i == 1:
            read = read[:40] + "N" + read[41:]
        if i == 2:
            read = read.lower()
        if len(read) < k:
            continue
        exp_f, exp_r = model.count(read), model.count(revcomp(read))
        assert sum(exp_f) > 0
        assert list(orc.count(read.encode())) == exp_f
        assert list(orc.count(read.encode(), revcomp=True)) == exp_r


def test_python_definition_reproduces_a_reference_fixture_count(golden_ibf_paths, known):
    """The model itself is anchored to the reference: the 35-mer known answer of read.hpp:113,139-140 (23 per bin, 0 reverse)
    against libIBFTests/data/test.ibf, evaluated with the model's hash / row / layout arithmetic on the file's words."""
    f = oracle.OracleIBF.load(golden_ibf_paths["lib_test"])
    words = f.words()
    m = ModelIBF(2, 3, 13, 79121216)
    read = "AAAAAAACCCCCCCCCGAGAGAGGAGAGAGGAGAG"

    def count(text):
        c = [0, 0]
        for j in range(len(text) - 12):
            rows = m.rows(text[j:j + 13])
            for b in range(2):
                if all((int(words[(r * m.block_bits + b) >> 6]) >> ((r * m.block_bits + b) & 63)) & 1 for r in rows):
                    c[b] += 1
        return c

    assert count(read) == [23, 23]
    assert count(revcomp(read)) == [0, 0]


# ---- build-side host logic: behaviour models of the reference's own loops --------------------------------------------------
NPOS = 1 << 64


def model_cut_out_nnns(seq):
    """IBF::cutOutNNNs + the concatenation at IBFBuild.cpp:81-88, statement by statement with Python string methods
    (std::string::find_first_not_of / find / substr semantics; npos = 2^64 - 1 compares greater than any length)."""
    seqlen, pieces, end = len(seq), [], 0
    while True:
        start = next((i for i in range(end, seqlen) if seq[i] != "N"), NPOS)       # find_first_not_of("N", end)
        if start == NPOS:
            break
        end = seq.find("N", start)
        end = NPOS if end < 0 else end
        if end > seqlen:
            pieces.append(seq[start:start + max(0, seqlen - start - 1)])            # substr(start, seqlen - start - 1)
            break
        pieces.append(seq[start:end])
    return "".join(pieces)


def model_fragment_schedule(seqlen, F, k):
    """The fragment loop of add_sequences_to_filter (IBFBuild.cpp:165-202) with its signed arithmetic; overlap_length = 1500
    only makes the first start negative, which clamps to 0."""
    out, idx = [], 0
    start = max(0, idx * F - 1500 + 1)
    while start < seqlen - 1:
        end = min((idx + 1) * F, seqlen)
        out.append((start, end))
        idx += 1
        start = idx * F - k + 1
    return out


def test_cut_out_nnns_equals_the_statement_level_model():
    rng = np.random.default_rng(5)
    cases = ["", "N", "NN", "A", "AN", "NA", "ANNA", "ACGT", "ACGTN", "NNACGTNN", "ACGTNNNACGT", "nACGTn", "ACGNnNT"]
    for _ in range(400):
        n = int(rng.integers(0, 40))
        cases.append("".join("ACGTNNNn"[i] for i in rng.integers(0, 8, size=n)))
    for s in cases:
        exp = model_cut_out_nnns(s).encode()
        assert oracle.cut_out_nnns(s) == exp, s
        assert rb.cut_out_nnns(s) == exp, s


def test_fragment_schedule_equals_the_statement_level_model():
    rng = np.random.default_rng(6)
    cases = [(0, 100, 13), (1, 100, 13), (2, 100, 13), (99, 100, 13), (100, 100, 13), (101, 100, 13), (188, 100, 13), (189, 100, 13),
             (190, 100, 13), (199990, 100000, 13), (199989, 100000, 13), (299999, 100000, 13), (5000000, 100000, 13), (4999999, 100000, 13)]
    for _ in range(600):
        F = int(rng.integers(20, 500))
        k = int(rng.integers(2, 19))
        cases.append((int(rng.integers(0, 6 * F)), F, k))
    for seqlen, F, k in cases:
        exp = model_fragment_schedule(seqlen, F, k)
        for impl in (oracle.fragment_schedule, rb.fragment_schedule):
            b, e = impl(seqlen, F, k)
            assert [(int(x), int(y)) for x, y in zip(b, e)] == exp, (impl.__module__, seqlen, F, k)


# ---- thresholds: calculateCI / NormalCDFInverse / the int16 -> uint16 wrap, restated with Python's math module ------------------
def model_ci(r, kmer_size, readlen, confidence):
    """interleave::calculateCI (IBF.hpp:320-338) with RationalApproximation / NormalCDFInverse (:268-308); kmer_size arrives as
    uint8_t; the uint16_t casts take the low 16 bits of the truncated value (what x86-64 does for in-range doubles)."""
    import math
    k = float(kmer_size & 0xFF)
    q = 1.0 - math.pow(1.0 - r, k)
    L = float(readlen) - k + 1.0
    var_n = (L * (1.0 - q) * (q * (2.0 * k + (2.0 / r) - 1.0) - 2.0 * k) + k * (k - 1.0) * math.pow(1.0 - q, 2.0)
             + (2.0 * (1.0 - q) / math.pow(r, 2.0)) * ((1.0 + (k - 1.0) * (1.0 - q)) * r - q))
    p = 1.0 - (1 - confidence) / 2.0
    c, d = (2.515517, 0.802853, 0.010328), (1.432788, 0.189269, 0.001308)

    def rational(t):
        return t - ((c[2] * t + c[1]) * t + c[0]) / (((d[2] * t + d[1]) * t + d[0]) * t + 1.0)

    z = -rational(math.sqrt(-2.0 * math.log(p))) if p < 0.5 else rational(math.sqrt(-2.0 * math.log(1.0 - p)))
    if var_n < 0:
        return None                                           # sqrt of a negative: NaN, cast undefined (reads shorter than ~k)
    low = int(math.floor(L * q - z * math.sqrt(var_n))) & 0xFFFF
    high = int(math.ceil(L * q + z * math.sqrt(var_n))) & 0xFFFF
    return low, high


def model_threshold(r, kmer_size, readlen, confidence=0.95):
    """IBFClassify.cpp:102-109: uint16_t readlen; int16_t threshold = readlen - k + 1 - ci.second; passed on as uint16_t."""
    ci = model_ci(r, kmer_size, readlen, confidence)
    if ci is None:
        return None
    return ((readlen & 0xFFFF) - kmer_size + 1 - ci[1]) & 0xFFFF


def test_thresholds_equal_the_python_restatement():
    assert model_ci(0.1, 13, 35, 0.95) == (5, 30) and model_threshold(0.1, 13, 35) == 65529        # read.hpp:156-164
    assert model_threshold(0.1, 13, 250) == 18 and model_threshold(0.08, 13, 250) == 34 and model_threshold(0.1, 15, 360) == 22
    rng = np.random.default_rng(7)
    for r in (0.03, 0.05, 0.08, 0.1, 0.12, 0.15, 0.2, 0.3):
        for k in (9, 13, 15, 17, 20, 31):
            lut_o, lut_p = oracle.threshold_lut(r, k), rb.threshold_lut(r, k)
            lengths = list(range(k, 1200)) + [int(x) for x in rng.integers(1200, 65536, size=600)] + [65535]
            for n in lengths:
                exp = model_threshold(r, k, n)
                if exp is None:
                    continue
                assert int(lut_o[n]) == exp and int(lut_p[n]) == exp, (r, k, n, exp, int(lut_o[n]), int(lut_p[n]))
            for n in (k, 35, 250, 354, 1500, 40000):
                ci = model_ci(r, k, n, 0.95)
                if ci is not None:
                    assert oracle.calculate_ci(r, k, n, 0.95) == ci and rb.calculate_ci(r, k, n, 0.95) == ci, (r, k, n)


def test_filter_size_bits_equals_the_python_restatement():
    """IBF::calculate_filter_size_bits (IBFBuild.cpp:404-413) and the bins-per-sequence rule (:90) with Python's math module."""
    import math

    def model(F, k, h, fp, bins):
        max_kmers = F - k + 1
        opt_bins = int(math.floor(bins / 64.0 + 1)) * 64
        bin_bits = int(math.ceil(-1 / (math.pow(1 - math.pow(fp, 1.0 / h), 1.0 / float(h * max_kmers)) - 1)))
        return bin_bits * opt_bins

    assert model(100000, 13, 3, 0.01, 2) == 79121216                              # createfilter.hpp:148
    rng = np.random.default_rng(8)
    for _ in range(500):
        F = int(rng.integers(50, 5_000_000))
        k = int(rng.integers(5, 32))
        h = int(rng.integers(1, 6))
        fp = float(rng.choice([0.001, 0.01, 0.05, 0.1, 0.3]))
        bins = int(rng.choice([1, 2, 63, 64, 65, 100, 127, 128, 129, 31008, 299520]))
        if F <= k:
            continue
        exp = model(F, k, h, fp, bins)
        assert oracle.filter_size_bits(F, k, h, fp, bins) == exp, (F, k, h, fp, bins)
        assert rb.ibf_size_bits(F, k, h, fp, bins) == exp, (F, k, h, fp, bins)
    for n in (0, 1, 99999, 100000, 100001, 4999999, 5000000):
        assert oracle.bins_for_sequence(n, 100000) == n // 100000 + 1
