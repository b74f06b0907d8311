"""GPU parity tests: the CUDA path, called through the C ABI (ctypes -> librb_ibf.so), against
the CPU oracle on the same inputs and against the reference's golden fixtures.  Bit-exact."""
import hashlib
import os

import numpy as np
import pytest

import oracle
import readbouncer_b200 as rb
from readbouncer_b200 import synth
from conftest import data_path, golden_words, read_fasta

pytestmark = pytest.mark.gpu

KERNELS = {"auto": 0, "tile": 1, "stream": 2, "table": 3, "table_atomic": 4, "table_warp": 5}


@pytest.fixture(autouse=True)
def _reset_kernel_choice():
    yield
    rb.set_count_kernel(0)
    rb.set_insert_kernel(0)


def make_filter_pair(n_seqs, seq_len, fragment_length, k=13, seed=100, n_hash=3):
    """Same synthetic reference inserted by the oracle (CPU) and by the GPU library."""
    ref = [synth.random_bases(seq_len, seed + i) for i in range(n_seqs)]
    plan = synth.build_plan(ref, fragment_length, k, n_hash=n_hash)
    of = oracle.OracleIBF.create(plan["n_bins"], n_hash, k, plan["n_bits"])
    of.insert_batch(plan["bases"], plan["frag_begin"], plan["frag_end"], plan["frag_bin"], n_threads=4)
    gf = rb.IBF.create(plan["n_bins"], n_hash, k, plan["n_bits"])
    gf.insert_batch(plan["bases"], plan["frag_begin"], plan["frag_end"], plan["frag_bin"])
    return plan, of, gf


def assert_same_results(got, exp, dense=True):
    assert np.array_equal(got["read_flag"], exp["short_read"])
    if dense:
        assert np.array_equal(got["counts_fwd"], exp["counts_fwd"])
        assert np.array_equal(got["counts_rev"], exp["counts_rev"])
    assert np.array_equal(got["max_count"], exp["max_count"])
    assert np.array_equal(got["hit"], exp["hit"])
    assert np.array_equal(got["argmax_bin"], exp["argmax_bin"])


# ---- golden fixtures ----------------------------------------------------------------------------------
def test_golden_load_info_store(known, golden_ibf_paths, tmp_path):
    exp = {"lib_test": (2, 3, 13, 79121216), "lib_test1": (4, 3, 13, 79121216), "classify_test": (2, 3, 15, 79119680)}
    for name, path in golden_ibf_paths.items():
        f = rb.IBF.load(path)
        assert (f.n_bins, f.n_hash, f.kmer_size, f.n_bits) == exp[name]
        assert f.bin_width == 1 and f.n_blocks == f.n_bits // 64 and f.n_shards == 1
        out = tmp_path / (name + ".gpu.ibf")
        f.store(out)
        assert hashlib.md5(out.read_bytes()).hexdigest() == known["ibf"][name]["md5"]


def test_golden_load_rejects_non_ibf(tmp_path):
    with pytest.raises(rb.RBError) as e:
        rb.IBF.load(data_path("lib_test.fasta"))
    assert e.value.status == 4                      # ParseIBFFileException (FASTA sniffing, configReader.cpp:210-224)
    with pytest.raises(rb.RBError) as e:
        rb.IBF.load(tmp_path / "nope.ibf")
    assert e.value.status == 5                      # MissingIBFFileException


@pytest.mark.parametrize("name,fasta,k", [("lib_test", "lib_test.fasta", 13), ("lib_test1", "lib_test1.fasta", 13),
                                          ("classify_test", "classify_test.fasta", 15)])
def test_golden_rebuild_on_gpu(known, golden_sparse, name, fasta, k):
    """GPU insert kernel + host build plan reproduce the reference's .ibf files bit for bit
    (the fixtures were written by a test build that queued every sequence twice)."""
    seqs = [s for _, s in read_fasta(data_path(fasta))]
    plan = synth.build_plan(seqs + seqs, 100000, k)
    meta, words = golden_words(known, golden_sparse, name)
    assert plan["n_bits"] + 256 == meta["bit_length"] and plan["n_bins"] == meta["tail"][0]
    f = rb.IBF.create(plan["n_bins"], 3, k, plan["n_bits"])
    f.insert_batch(plan["bases"], plan["frag_begin"], plan["frag_end"], plan["frag_bin"])
    assert np.array_equal(f.download(), words[:plan["n_bits"] // 64])


@pytest.mark.parametrize("kernel", ["tile", "stream"])
def test_golden_known_answers(known, golden_ibf_paths, kernel):
    rb.set_count_kernel(KERNELS[kernel])
    ka = known["known"]
    f0 = rb.IBF.load(golden_ibf_paths["lib_test"])
    f1 = rb.IBF.load(golden_ibf_paths["lib_test1"])
    reads = [ka["read35"]["seq"].encode(), ka["read35_revcomp"]["seq"].encode(), ka["read354"]["seq"].encode()]
    off = np.cumsum([0] + [len(r) for r in reads]).astype(np.uint64)
    bases = np.frombuffer(b"".join(reads), np.uint8)
    lut = rb.threshold_lut(0.1, 13)
    assert lut[35] == 65529 and lut[354] == 36
    r0 = f0.count_batch(bases, off, lut, dense=True)
    r1 = f1.count_batch(bases, off, lut, dense=True)
    assert r0["counts_fwd"][0].tolist() == [23, 23] and r0["counts_rev"][0].tolist() == [0, 0]    # read.hpp:139-140
    assert r0["counts_fwd"][1].tolist() == [0, 0] and r0["counts_rev"][1].tolist() == [23, 23]    # read.hpp:286-327
    assert r0["max_count"].tolist() == [0, 0, 282] and r1["max_count"].tolist() == [0, 0, 182]    # read.hpp:221-229
    assert r0["hit"].tolist() == [0, 0, 1] and r0["argmax_bin"].tolist() == [0xFFFFFFFF, 0xFFFFFFFF, 0]
    # a zero threshold table reproduces the int-compare of CountMatchesTest: max 23 (read.hpp:327)
    rz = f0.count_batch(bases, off, np.zeros(65536, np.uint16))
    assert rz["max_count"].tolist()[:2] == [23, 23] and rz["hit"].tolist() == [1, 1, 1]


def test_golden_classify_fastq(golden_ibf_paths):
    """classifyTests: 3/3 reads found in <= 4 chunks of 360, all in chunk 0 at 250."""
    f = rb.IBF.load(golden_ibf_paths["classify_test"])
    of = oracle.OracleIBF.load(golden_ibf_paths["classify_test"])
    reads = [s for _, s in read_fasta(data_path("classify_test.fastq"))]
    for cl, expect in [(250, [0, 0, 0]), (360, [0, 0, 3])]:
        chunks, owner = [], []
        for ri, s in enumerate(reads):
            for i in range(5):
                chunks.append(s[i * cl:min((i + 1) * cl, len(s))])
                owner.append((ri, i))
        off = np.cumsum([0] + [len(c) for c in chunks]).astype(np.uint64)
        bases = np.frombuffer(b"".join(chunks), np.uint8)
        lut = rb.threshold_lut(0.1, 15)
        got = f.count_batch(bases, off, lut, dense=True)
        exp = of.count_batch(bases, off, oracle.threshold_lut(0.1, 15))
        assert_same_results(got, exp)
        first = [min(i for (r, i), m in zip(owner, got["max_count"]) if r == ri and m > 0) for ri in range(3)]
        assert first == expect


# ---- synthetic differential tests ---------------------------------------------------------------------
def wtable_bytes(k, bin_width, span):
    """Size of the window k-mer table (ibf_wtable.cu: wtable_geometry), None when not applicable."""
    L = k + span - 1
    if bin_width > 2 or span < 2 or L > 16:
        return None
    lanes = 2 if span == 2 else 4
    entries = 2 ** (2 * L - 1) if L % 2 else 4 ** L
    return entries * lanes * 16 * bin_width


RAGGED = [250] * 40 + [0, 1, 12, 13, 14, 31, 32, 33, 64, 100, 249, 251, 360, 500, 1023, 1024, 1036, 1037, 1500, 2100, 5000]


@pytest.mark.parametrize("kernel", ["tile", "stream", "table", "table_atomic"])
@pytest.mark.parametrize("n_seqs,seq_len,frag,k", [
    (1, 500000, 100000, 13),     # 6 bins,  W=1   (config #1 shape)
    (100, 20000, 21000, 13),     # 100 bins, W=2  (config #2 shape)
    (130, 3000, 4000, 15),       # 130 bins, W=3
    (200, 3000, 4000, 13),       # 200 bins, W=4
    (1, 700 * 2000 + 7, 2000, 13),   # 701 bins, W=11 (odd stride -> 8-byte loads, tiles 4+4+3)
    (1100, 1500, 2000, 11),      # 1100 bins, W=18 (even stride -> 16-byte loads)
    (60, 3000, 4000, 10),        # 60 bins, W=1, k=10: window tables of span 2 (canonical), 3, 4 (canonical)
    (100, 3000, 4000, 11),       # 100 bins, W=2, k=11: span 2, 3 (canonical), 4
])
def test_count_matches_oracle(kernel, n_seqs, seq_len, frag, k):
    rb.set_count_kernel(KERNELS[kernel])
    plan, of, gf = make_filter_pair(n_seqs, seq_len, frag, k)
    assert np.array_equal(gf.download(), of.words()[:plan["n_bits"] // 64])
    bases, off = synth.ragged_reads(plan["bases"], RAGGED, seed=7, frac_from_ref=0.7, n_frac=0.003, lower_frac=0.1)
    lut = rb.threshold_lut(0.1, k)
    if kernel.startswith("table"):
        if gf.bin_width > 4:                             # wide rows: postings table (sorted bin lists per k-mer)
            exp = of.count_batch(bases, off, lut, n_threads=4)
            short = [250] * 70 + [0, 1, k - 1, k, k + 1, 31, 32, 33, 64, 100, 249, 251, 254 + k, 255 + k - 1]
            sb, so = synth.ragged_reads(plan["bases"], short, seed=12, frac_from_ref=0.7, n_frac=0.004, lower_frac=0.1)
            sb[int(so[5]):int(so[6])] = ord("N")
            sb[int(so[7]) + 100] = ord("U")
            sexp = of.count_batch(sb, so, lut, n_threads=4)
            if gf.bin_width <= 32:                       # rows of 5..32 words: the group-loaded k-mer table comes first
                os.environ["RB_CTABLE_WIDE"] = "1"       # (17..32 words: only for long lists unless forced)
                try:
                    gf.enable_kmer_table(0)
                finally:
                    del os.environ["RB_CTABLE_WIDE"]
                lanes = 8 if gf.bin_width <= 8 else 16 if gf.bin_width <= 16 else 32
                assert (gf.kmer_table_kind(), gf.kmer_table_bytes(), gf.kmer_table_span()) == (4, 4 ** k * lanes * 16, 1)
                assert_same_results(gf.count_batch(bases, off, lut, dense=True), exp)
                assert_same_results(gf.count_batch(sb, so, lut, dense=True), sexp)
                assert_same_results(gf.count_batch(sb, so, lut, dense=False), sexp, dense=False)
                gf.disable_kmer_table()
            # lists: pointer + variable-length lists (default); slots: one fixed slot per k-mer fetched by bulk copies
            for layout, kind in (("slots", 3), ("lists", 2)):
                os.environ["RB_POSTINGS_LAYOUT"] = layout
                os.environ["RB_CTABLE"] = "0"
                try:
                    gf.enable_kmer_table(0)
                finally:
                    del os.environ["RB_POSTINGS_LAYOUT"], os.environ["RB_CTABLE"]
                assert gf.kmer_table_kind() == kind and gf.kmer_table_bytes() > 4 ** k * 4
                for sub in (("", "0", "2", "4", "8") if layout == "lists" else ("",)):     # lanes per list of the lists kernel
                    if sub:
                        os.environ["RB_POSTINGS_SUB"] = sub
                    try:
                        assert_same_results(gf.count_batch(bases, off, lut, dense=True), exp)          # long reads: 16-bit counters
                        assert_same_results(gf.count_batch(sb, so, lut, dense=True), sexp)             # <= 255 positions: 8-bit counters
                        assert_same_results(gf.count_batch(sb, so, lut, dense=False), sexp, dense=False)
                    finally:
                        os.environ.pop("RB_POSTINGS_SUB", None)
                gf.disable_kmer_table()
                assert gf.kmer_table_kind() == 0
            return
        exp = of.count_batch(bases, off, lut, n_threads=4)
        span1 = 4 ** k * 16 * gf.bin_width
        gf.enable_kmer_table(span1)                      # budget admits only one k-mer per entry
        assert (gf.kmer_table_bytes(), gf.kmer_table_span()) == (span1, 1)
        # rows of 3-4 words: entries padded to 4 words and loaded by 4 lanes (ibf_ctable.cu) when that fits the budget
        assert gf.kmer_table_kind() == (4 if gf.bin_width == 4 else 1)
        assert_same_results(gf.count_batch(bases, off, lut, dense=True), exp)
        if gf.bin_width in (3, 4):                       # ... and the lane-per-entry kernel of ibf_table.cu on the unpadded layout
            os.environ["RB_CTABLE"] = "0"
            try:
                gf.enable_kmer_table(span1)
            finally:
                del os.environ["RB_CTABLE"]
            assert (gf.kmer_table_kind(), gf.kmer_table_bytes()) == (1, span1)
            assert_same_results(gf.count_batch(bases, off, lut, dense=True), exp)
            if gf.bin_width == 3 and 4 ** k * 64 < (70 << 30):
                gf.enable_kmer_table(4 ** k * 64)        # the padded layout for 3-word rows
                assert (gf.kmer_table_kind(), gf.kmer_table_bytes()) == (4, 4 ** k * 64)
                assert_same_results(gf.count_batch(bases, off, lut, dense=True), exp)
        if gf.bin_width <= 2 and kernel == "table":      # window tables: 2..4 consecutive k-mers per entry
            for span in (2, 3, 4):
                need = wtable_bytes(k, gf.bin_width, span)
                if need is None or need > (70 << 30):
                    continue
                os.environ["RB_KMER_TABLE_SPAN"] = str(span)
                try:
                    gf.enable_kmer_table(0)
                finally:
                    del os.environ["RB_KMER_TABLE_SPAN"]
                assert (gf.kmer_table_bytes(), gf.kmer_table_span()) == (need, span)
                assert_same_results(gf.count_batch(bases, off, lut, dense=True), exp)     # long reads: warp per read
                # short reads only (<= 127*span positions, <= 561 bases): one read per group of lanes
                limit = min(127 * span + k - 1, 561)
                short = [250] * 70 + [0, 1, k - 1, k, k + 1, k + span - 1, k + span, 31, 32, 33, 47, 48, 49, 64, 100,
                                      249, 251, 255, 256, 257, limit - 1, limit]
                sb, so = synth.ragged_reads(plan["bases"], short, seed=11 + span, frac_from_ref=0.7, n_frac=0.004,
                                            lower_frac=0.1)
                sb[int(so[5]):int(so[6])] = ord("N")                        # an all-N read
                sb[int(so[7]) + 100] = ord("U")                             # U counts as T
                sb[int(so[8]) + 249] = ord("R")                             # IUPAC code in the last window
                sexp = of.count_batch(sb, so, lut, n_threads=4)
                assert_same_results(gf.count_batch(sb, so, lut, dense=True), sexp)
                assert_same_results(gf.count_batch(sb, so, lut, dense=False), sexp, dense=False)
                rb.set_count_kernel(KERNELS["table_warp"])                  # same reads through the warp-per-read kernel
                assert_same_results(gf.count_batch(sb, so, lut, dense=True), sexp)
                rb.set_count_kernel(KERNELS[kernel])
        return
    assert np.array_equal(lut, oracle.threshold_lut(0.1, k))
    got = gf.count_batch(bases, off, lut, dense=True)
    exp = of.count_batch(bases, off, lut, n_threads=4)
    assert_same_results(got, exp)
    assert exp["hit"].sum() > 10 and (exp["short_read"] == 1).sum() >= 2


def test_kmer_table_handles_n_rich_reads_and_is_dropped_by_insert():
    """Windows containing N are not in the table and must take the hashed path; inserts invalidate the table."""
    plan, of, gf = make_filter_pair(100, 20000, 21000, 13)
    bases, off = synth.ragged_reads(plan["bases"], [250] * 64 + [13, 14, 40, 1200], seed=21, frac_from_ref=0.9,
                                    n_frac=0.05, lower_frac=0.3)
    bases[int(off[3]):int(off[4])] = ord("N")                  # an all-N read
    bases[int(off[5])] = ord("U")                               # U counts as T
    lut = rb.threshold_lut(0.1, 13)
    exp = of.count_batch(bases, off, lut, n_threads=4)
    for which, span in ((3, 3), (3, 2), (3, 1), (4, 1)):
        rb.set_count_kernel(which)
        os.environ["RB_KMER_TABLE_SPAN"] = str(span)
        try:
            gf.enable_kmer_table(0)
        finally:
            del os.environ["RB_KMER_TABLE_SPAN"]
        assert gf.kmer_table_span() == span
        assert_same_results(gf.count_batch(bases, off, lut, dense=True), exp)
    rb.set_count_kernel(0)
    # insert one more fragment into bin 0: table must be rebuilt, results must follow the new matrix
    extra = synth.random_bases(5000, 999)
    gf.insert_batch(extra, [0], [5000], [0])
    of.insert_batch(extra, [0], [5000], [0])
    assert gf.kmer_table_bytes() == 0
    b2, o2 = synth.ragged_reads(extra, [250] * 16, seed=1, frac_from_ref=1.0, error_rate=0.0)
    rb.set_count_kernel(3)
    got = gf.count_batch(b2, o2, lut, dense=True)
    assert_same_results(got, of.count_batch(b2, o2, lut))
    assert got["hit"].all() and gf.kmer_table_bytes() > 0


def test_host_buffer_pipeline_packed_pieces_and_rounds(monkeypatch):
    """rb_ibf_count_batch with many pieces and several staging rounds: reads packed into bit planes by the host threads
    (ragged lengths, offsets not multiples of 32, N / IUPAC / U / lower case, an over-long read that sends its piece down
    the ASCII path) must give the oracle's answers, pinned and pageable result buffers alike."""
    import torch
    plan, of, gf = make_filter_pair(100, 20000, 21000, 13)
    rng = np.random.default_rng(5)
    lengths = rng.integers(0, 394, size=60000).tolist()
    lengths[40000] = 3000                                   # not a group-kernel read: its piece is shipped as ASCII
    bases, off = synth.ragged_reads(plan["bases"], lengths, seed=3, frac_from_ref=0.6, n_frac=0.002, lower_frac=0.05)
    bases[::9973] = ord("U"); bases[5::7919] = ord("R")
    luts = np.stack([rb.threshold_lut(0.1, 13), rb.threshold_lut(0.08, 13)])
    gf.enable_kmer_table(0)
    assert gf.kmer_table_span() == 3
    monkeypatch.setenv("RB_PIECE_MB", "1")                  # ~12 pieces
    monkeypatch.setenv("RB_STAGE_MB", "4")                  # ~3 rounds
    got = gf.count_batch(bases, off, luts)                  # pageable numpy buffers: results copied at the end
    exp = [of.count_batch(bases, off, luts[t], n_threads=8) for t in range(2)]
    for t in range(2):
        for key in ("max_count", "hit", "argmax_bin"):
            assert np.array_equal(got[key][t], exp[t][key]), (t, key)
    assert np.array_equal(got["read_flag"], exp[0]["short_read"])
    # pinned inputs and outputs (mapped result stores), straight through the C ABI
    n = len(lengths)
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
    hb, ho = pin(bases), pin(off)
    r_max = torch.empty(2 * n, dtype=torch.int16, pin_memory=True).numpy().view(np.uint16)
    r_hit = torch.empty(2 * n, dtype=torch.uint8, pin_memory=True).numpy()
    r_am = torch.empty(2 * n, dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
    r_flag = torch.empty(n, dtype=torch.uint8, pin_memory=True).numpy()
    P = rb.capi._np_ptr
    xfer = {}
    for pack, share in (("1", None), ("1", "0.25"), ("1", "0.5"), ("0", None)):
        # pinned bases: a share of the pieces crosses PCIe as ASCII while the host threads pack the others
        monkeypatch.setenv("RB_HOST_PACK", pack)
        if share is None:
            monkeypatch.delenv("RB_ASCII_SHARE", raising=False)
        else:
            monkeypatch.setenv("RB_ASCII_SHARE", share)
        x0 = rb.transfer_bytes()[0]
        r_max[:] = 0xFFFF; r_hit[:] = 7; r_am[:] = 5; r_flag[:] = 9
        rb.capi._check(rb.lib().rb_ibf_count_batch(gf._h, P(hb), P(ho), n, P(luts), 2, None, None, P(r_max), P(r_hit), P(r_am),
                                                   P(r_flag), None))
        for t in range(2):
            assert np.array_equal(r_max[t * n:(t + 1) * n], exp[t]["max_count"]), pack
            assert np.array_equal(r_hit[t * n:(t + 1) * n], exp[t]["hit"]), pack
            assert np.array_equal(r_am[t * n:(t + 1) * n], exp[t]["argmax_bin"]), pack
        assert np.array_equal(r_flag, exp[0]["short_read"]), pack
        xfer[(pack, share)] = rb.transfer_bytes()[0] - x0
    assert xfer[("1", None)] < xfer[("1", "0.25")] < xfer[("1", "0.5")] < xfer[("0", None)]


def test_packed_pieces_ship_the_bad_plane_only_when_needed(monkeypatch):
    """The packed transfer writes and ships the "not ACGT" plane of a piece only if the piece has such a base: a batch whose
    N / IUPAC bases sit in a few reads (one dirty 128 K-base task inside an otherwise clean piece, one dirty piece between
    clean ones, a dirty last word) gives the oracle's answers, twice over the same staging memory (a clean piece after a dirty
    one must not see its stale plane), and crosses PCIe with ~2 instead of 3 bits per base."""
    plan, of, gf = make_filter_pair(100, 20000, 21000, 13)
    n = 40000
    bases, off, _ = synth.sample_reads(plan["bases"], n, 250, seed=77)
    luts = np.stack([rb.threshold_lut(0.1, 13), rb.threshold_lut(0.08, 13)])
    gf.enable_kmer_table(0)
    assert gf.kmer_table_span() == 3
    monkeypatch.setenv("RB_PIECE_MB", "1")                  # 4 194 reads per piece, 8 tasks per piece
    monkeypatch.setenv("RB_HOST_PACK", "1")

    def run(b):
        x0 = rb.transfer_bytes()[0]
        got = gf.count_batch(b, off, luts)
        moved = rb.transfer_bytes()[0] - x0
        for t in range(2):
            exp = of.count_batch(b, off, luts[t], dense=False, n_threads=8)
            for key in ("max_count", "hit", "argmax_bin"):
                assert np.array_equal(got[key][t], exp[key]), (t, key)
        return moved

    run(bases)                                              # (the first call also uploads the threshold tables)
    clean = run(bases)                                      # no such base anywhere: two planes per piece
    dirty = bases.copy()
    for r, at in ((9000, 17), (9001, 249), (21000, 100), (n - 1, 249)):     # pieces 2, 5 and the last one
        dirty[r * 250 + at] = ord("N")
    dirty[30000 * 250:30010 * 250] = ord("R")               # ten reads of IUPAC codes in piece 7
    some = run(dirty)
    again = run(bases)                                      # the same staging memory, clean again
    every = run(np.where(np.arange(bases.size) % 1000 == 0, ord("N"), bases).astype(np.uint8))
    nb = bases.size
    assert again == clean and clean < some < every
    assert abs(clean - (nb / 4 + 8 * (n + 1))) < 0.02 * nb and abs(every - (3 * nb / 8 + 8 * (n + 1))) < 0.02 * nb
    assert some - clean < 5 * (1 << 20) / 8 + 4096          # four dirty pieces of 1 MB of bases: one more bit per base each


@pytest.mark.parametrize("n_bins,k,layout", [(100, 11, ""), (60, 10, ""), (300, 10, ""), (1000, 10, ""), (1100, 11, "lists"),
                                             (1100, 11, "slots")])
def test_broken_max_read_len_promise_is_flagged(n_bins, k, layout, monkeypatch):
    """rb_ibf_count_batch_dev: max_read_len selects narrow counters.  A read longer than promised must come back as read_flag 3
    with key 0 (or, from kernels that do not rely on the promise, fully classified) -- never with wrapped counters -- and the
    other reads of the batch must be unaffected.  Window table, group-loaded table, postings lists and slots."""
    import torch
    if layout:
        monkeypatch.setenv("RB_POSTINGS_LAYOUT", layout)
        monkeypatch.setenv("RB_CTABLE", "0")
    plan, of, gf = make_filter_pair(n_bins, 1500, 2000, k)
    gf.enable_kmer_table(0)
    lengths = [250] * 20 + [1400, 250, 300, 249]                       # reads 20 and 22 break the promise of 250
    bases, off = synth.ragged_reads(plan["bases"], lengths, seed=41, frac_from_ref=1.0, n_frac=0.0)
    lut = rb.threshold_lut(0.1, k)
    exp = of.count_batch(bases, off, lut, dense=False, n_threads=4)
    n = len(lengths)
    d_b, d_o = torch.from_numpy(bases).cuda(), torch.from_numpy(off.astype(np.int64)).cuda()
    d_lut = torch.from_numpy(lut.view(np.int16)).cuda()
    d_keys = torch.full((n,), -1, dtype=torch.int64, device="cuda")
    d_flag = torch.full((n,), 77, dtype=torch.uint8, device="cuda")
    gf.count_batch_dev(d_b, d_o, n, d_lut, 1, d_keys, max_read_len=250, d_read_flag=d_flag)
    torch.cuda.synchronize()
    flag = d_flag.cpu().numpy()
    mx, hit, am = rb.keys_decode(d_keys.cpu().numpy().view(np.uint64))
    for i in range(n):
        if flag[i] == 3:
            assert lengths[i] > 250 and (mx[i], hit[i]) == (0, 0), i
        else:
            assert flag[i] == exp["short_read"][i], i
            assert (mx[i], hit[i], am[i]) == (exp["max_count"][i], exp["hit"][i], exp["argmax_bin"][i]), i
    assert exp["hit"][20] == 1 and exp["max_count"][20] > 255           # the case that would wrap an 8-bit counter


@pytest.mark.parametrize("n_bins,k,env", [(300, 10, {}), (1000, 10, {}), (1100, 11, {"RB_CTABLE": "0", "RB_POSTINGS_LAYOUT": "lists"}),
                                          (1100, 11, {"RB_CTABLE": "0", "RB_POSTINGS_LAYOUT": "slots", "RB_SLOT_BYTES": "128"})])
def test_lane_group_kernels_edge_batches(n_bins, k, env, monkeypatch):
    """Group-loaded table, 2-lane list kernel and slots read by lane groups on degenerate batches: four threshold tables in one
    pass, a single read, reads that are empty / shorter than k / all N / all lower case, and a batch smaller than one CTA."""
    for key, val in env.items():
        monkeypatch.setenv(key, val)
    plan, of, gf = make_filter_pair(n_bins, 1500, 2000, k)
    gf.enable_kmer_table(0)
    assert gf.kmer_table_kind() == (4 if not env else 2 if env["RB_POSTINGS_LAYOUT"] == "lists" else 3)
    luts = np.stack([rb.threshold_lut(e, k) for e in (0.1, 0.08, 0.05, 0.15)])
    bases, off = synth.ragged_reads(plan["bases"], [250, 0, k - 1, 250, 250, 300, 250], seed=5, frac_from_ref=1.0, n_frac=0.0)
    bases[int(off[3]):int(off[4])] = ord("N")                                   # all N: every window takes the hashed path
    lo, hi = int(off[4]), int(off[5])
    bases[lo:hi] = np.frombuffer(bytes(bases[lo:hi]).lower(), np.uint8)          # all lower case
    got = gf.count_batch(bases, off, luts)
    for t in range(4):
        exp = of.count_batch(bases, off, luts[t], dense=False, n_threads=4)
        for key in ("max_count", "hit", "argmax_bin"):
            assert np.array_equal(got[key][t], exp[key]), (t, key)
    assert np.array_equal(got["read_flag"], exp["short_read"]) and got["hit"][0][4] == 1
    # (the all-N read is one k-mer 241 times: whatever bins its three rows share, it "hits" them -- like the reference)
    one_b, one_o = bases[:250].copy(), np.array([0, 250], np.uint64)             # a single read
    assert_same_results(gf.count_batch(one_b, one_o, luts[0], dense=True), of.count_batch(one_b, one_o, luts[0], n_threads=1))


@pytest.mark.parametrize("inflight", ["", "atomic", "1", "8"])
@pytest.mark.parametrize("n_bins,k", [(129, 11), (192, 10), (256, 11), (257, 11), (320, 10), (512, 11), (513, 10), (1000, 11), (1024, 10),
                                        (1025, 10), (1500, 11), (2048, 10)])
def test_medium_rows_group_loaded_table(n_bins, k, inflight, monkeypatch):
    """Rows of 3..32 words (129..2048 bins): k-mer table entries padded to 4 / 8 / 16 / 32 words, one entry per group of as many lanes
    (ibf_ctable.cu).  Every width class at both ends, ragged and multi-chunk reads, N / IUPAC windows (hashed on the fly), two
    threshold tables in one pass, dense counts; bit-sliced register counters (default) and the shared-memory-atomic variant with
    the default and extreme numbers of entries in flight per lane.  Reads up to 250 bases take the single-chunk accumulator
    (max_read_len known), the ragged batch the 16-plane one."""
    if inflight:                                        # the shared-memory-atomic variant of the kernel (default: bit-sliced)
        monkeypatch.setenv("RB_CTABLE_ATOMIC", "1")
        if inflight != "atomic":
            if (n_bins, k) not in ((192, 10), (512, 11), (1000, 11), (1500, 11)):
                pytest.skip("in-flight variants on one shape per width class")
            monkeypatch.setenv("RB_CTABLE_U", inflight)
    plan, of, gf = make_filter_pair(n_bins, 1500, 2000, k)
    assert plan["n_bins"] == n_bins
    if n_bins > 1024:
        # 17..32 row words: by default the table is taken only when the sampled postings lists are long (> 4 units); these
        # filters have ~1.5-unit lists, so the automatic choice is the short-list postings kernel
        gf.enable_kmer_table(0)
        assert gf.kmer_table_kind() == 2
        monkeypatch.setenv("RB_CTABLE_WIDE", "1")
    gf.enable_kmer_table(0)
    lanes = 4 if n_bins <= 256 else 8 if n_bins <= 512 else 16 if n_bins <= 1024 else 32
    assert (gf.kmer_table_kind(), gf.kmer_table_span(), gf.kmer_table_bytes()) == (4, 1, 4 ** k * lanes * 16)
    bases, off = synth.ragged_reads(plan["bases"], RAGGED + [k - 1, k, k + 1, 1023 + k, 1024 + k, 3000], seed=31, frac_from_ref=0.7,
                                    n_frac=0.004, lower_frac=0.1)
    bases[int(off[3]):int(off[4])] = ord("N")
    bases[int(off[6]) + 100] = ord("R")
    luts = np.stack([rb.threshold_lut(0.1, k), rb.threshold_lut(0.08, k)])
    for t in range(2):
        exp = of.count_batch(bases, off, luts[t], n_threads=8)
        assert_same_results(gf.count_batch(bases, off, luts[t], dense=True), exp)
    both = gf.count_batch(bases, off, luts)
    for t in range(2):
        exp = of.count_batch(bases, off, luts[t], dense=False, n_threads=8)
        for key in ("max_count", "hit", "argmax_bin"):
            assert np.array_equal(both[key][t], exp[key]), (t, key)
    short = [250] * 70 + [0, 1, k - 1, k, k + 1, 31, 32, 33, 64, 100, 249, 251, 253 + k, 254 + k - 1]
    sb, so = synth.ragged_reads(plan["bases"], short, seed=12, frac_from_ref=0.7, n_frac=0.004, lower_frac=0.1)
    sb[int(so[5]):int(so[6])] = ord("N")
    sexp = of.count_batch(sb, so, luts[0], n_threads=8)
    assert_same_results(gf.count_batch(sb, so, luts[0], dense=True), sexp)
    assert_same_results(gf.count_batch(sb, so, luts[0], dense=False), sexp, dense=False)
    # the roofline's traffic figure: one entry per k-mer position, whole 128-byte lines
    import torch
    d_b, d_o = torch.from_numpy(bases).cuda(), torch.from_numpy(off.astype(np.int64)).cuda()
    tb, reqs, io = gf.count_traffic_dev(d_b, d_o, len(off) - 1, 1)
    npos = sum(max(0, int(n) - k + 1) for n in np.diff(off.astype(np.int64)) if n <= 65535)
    assert reqs == npos and tb == npos * max(128, lanes * 16)


@pytest.mark.parametrize("order", ["1", "0"])
@pytest.mark.parametrize("n_blocks", [2600, 3700, 6000, 9000])
@pytest.mark.parametrize("slot_bytes,ring", [(0, 0), (128, 0), (256, 1), (1024, 2)])
def test_postings_slots(n_blocks, order, slot_bytes, ring, monkeypatch):
    """The same over-full filters through the SLOT layout: the sampled slot size (0) and forced ones -- 128 / 256 bytes push most
    / many lists into the overflow area, 1 024 holds nearly all -- with ring depths 1 and 2 next to the default."""
    monkeypatch.setenv("RB_POSTINGS_ORDER", order)
    monkeypatch.setenv("RB_POSTINGS_LAYOUT", "slots")
    monkeypatch.setenv("RB_CTABLE", "0")                    # 18 row words: the group-loaded k-mer table would come first
    if slot_bytes:
        monkeypatch.setenv("RB_SLOT_BYTES", str(slot_bytes))
    if ring:
        monkeypatch.setenv("RB_SLOT_RING", str(ring))
    k, n_hash = 11, 3
    ref = [synth.random_bases(1500, 300 + i) for i in range(1100)]
    plan = synth.build_plan(ref, 2000, k, n_hash=n_hash)
    n_bits = n_blocks * 64 * 18
    of = oracle.OracleIBF.create(1100, n_hash, k, n_bits)
    of.insert_batch(plan["bases"], plan["frag_begin"], plan["frag_end"], plan["frag_bin"], n_threads=4)
    gf = rb.IBF.create(1100, n_hash, k, n_bits)
    gf.insert_batch(plan["bases"], plan["frag_begin"], plan["frag_end"], plan["frag_bin"])
    gf.enable_kmer_table(0)
    assert gf.kmer_table_kind() == 3
    if slot_bytes:
        assert gf.kmer_table_bytes() >= 4 ** k * slot_bytes
    lut = rb.threshold_lut(0.1, k)
    short = [250] * 40 + [0, 1, k - 1, k, k + 1, 31, 64, 100, 249, 251, 254 + k]
    sb, so = synth.ragged_reads(plan["bases"], short, seed=21, frac_from_ref=0.7, n_frac=0.004, lower_frac=0.1)
    sexp = of.count_batch(sb, so, lut, n_threads=8)
    longer = [250] * 8 + [255 + k, 400, 700, 1500]
    lb, lo = synth.ragged_reads(plan["bases"], longer, seed=22, frac_from_ref=0.7, n_frac=0.003, lower_frac=0.1)
    lexp = of.count_batch(lb, lo, lut, n_threads=8)
    # slots of 128 / 256 bytes: 8 / 16 lanes load a slot straight into registers (default), or the bulk-copy ring kernel
    for direct in (("1", "0") if gf.kmer_table_bytes() < 4 ** k * 384 + (64 << 20) else ("1",)):
        monkeypatch.setenv("RB_SLOTS_SUB", direct)
        assert_same_results(gf.count_batch(sb, so, lut, dense=True), sexp)                 # 8-bit counters
        assert_same_results(gf.count_batch(sb, so, lut, dense=False), sexp, dense=False)
        assert_same_results(gf.count_batch(lb, lo, lut, dense=True), lexp)                 # 16-bit counters
    # the measured-geometry traffic of the roofline: a slot per (position, strand) + what overflowed
    import torch
    d_b, d_o = torch.from_numpy(sb).cuda(), torch.from_numpy(so.astype(np.int64)).cuda()
    tb, reqs, io = gf.count_traffic_dev(d_b, d_o, len(so) - 1, 1)
    pairs = 2 * sum(max(0, x - k + 1) for x in short if x <= 65535)
    sbytes = (gf.kmer_table_bytes() // 4 ** k) // 128 * 128 if not slot_bytes else slot_bytes
    assert tb >= 0.95 * pairs * min(sbytes, 128) and reqs >= 0.95 * pairs and io > sb.size


@pytest.mark.parametrize("sub", ["", "0", "2", "4", "8"])
@pytest.mark.parametrize("order", ["1", "0"])
@pytest.mark.parametrize("n_blocks", [2600, 3700, 6000, 9000])
def test_postings_long_lists(n_blocks, order, sub, monkeypatch):
    """Postings lists of ~610 / ~380 / ~160 / ~70 bins per k-mer (an over-full 1 100-bin filter: 56 % .. 6 % false positives
    per bin), so that the lookup kernel walks full rounds, further rounds and every tail width (ibf_postings_layout.cuh),
    with the ids dealt over the groups (default) and ascending (RB_POSTINGS_ORDER=0)."""
    monkeypatch.setenv("RB_POSTINGS_ORDER", order)
    monkeypatch.setenv("RB_POSTINGS_LAYOUT", "lists")
    monkeypatch.setenv("RB_CTABLE", "0")                    # 18 row words: the group-loaded k-mer table would come first
    if sub:                                                 # lanes per list: 0 = the whole warp, 2 / 4 / 8 = the short-list kernel
        monkeypatch.setenv("RB_POSTINGS_SUB", sub)          # ("" = chosen by the table's mean list length)
    k, n_hash = 11, 3
    ref = [synth.random_bases(1500, 300 + i) for i in range(1100)]
    plan = synth.build_plan(ref, 2000, k, n_hash=n_hash)
    assert plan["n_bins"] == 1100
    n_bits = n_blocks * 64 * 18
    of = oracle.OracleIBF.create(1100, n_hash, k, n_bits)
    of.insert_batch(plan["bases"], plan["frag_begin"], plan["frag_end"], plan["frag_bin"], n_threads=4)
    gf = rb.IBF.create(1100, n_hash, k, n_bits)
    gf.insert_batch(plan["bases"], plan["frag_begin"], plan["frag_end"], plan["frag_bin"])
    assert np.array_equal(gf.download(), of.words()[:n_bits // 64])
    gf.enable_kmer_table(0)
    assert gf.kmer_table_kind() == 2
    mean_ids = (gf.kmer_table_bytes() - 4 ** k * 4) / 2 / 4 ** k
    fp = (1 - np.exp(-n_hash * (1500 - k + 1) / n_blocks)) ** n_hash
    assert 0.8 * 1100 * fp < mean_ids < 1.2 * 1100 * fp + 8
    lut = rb.threshold_lut(0.1, k)
    short = [250] * 40 + [0, 1, k - 1, k, k + 1, 31, 64, 100, 249, 251, 254 + k]
    sb, so = synth.ragged_reads(plan["bases"], short, seed=21, frac_from_ref=0.7, n_frac=0.004, lower_frac=0.1)
    sexp = of.count_batch(sb, so, lut, n_threads=8)
    assert_same_results(gf.count_batch(sb, so, lut, dense=True), sexp)                 # 8-bit counters
    assert_same_results(gf.count_batch(sb, so, lut, dense=False), sexp, dense=False)
    longer = [250] * 8 + [255 + k, 400, 700, 1500]
    lb, lo = synth.ragged_reads(plan["bases"], longer, seed=22, frac_from_ref=0.7, n_frac=0.003, lower_frac=0.1)
    assert_same_results(gf.count_batch(lb, lo, lut, dense=True), of.count_batch(lb, lo, lut, n_threads=8))   # 16-bit counters


def test_transfer_policy_times_both_ways(monkeypatch):
    """Large host-buffer batches with RB_HOST_PACK unset: two packed calls, two ASCII calls, then the faster way for good;
    every call returns the oracle's answers."""
    plan, of, gf = make_filter_pair(100, 20000, 21000, 13)
    bases, off, _ = synth.sample_reads(plan["bases"], 40000, 250, seed=31)
    lut = rb.threshold_lut(0.1, 13)
    exp = of.count_batch(bases, off, lut, n_threads=8)
    gf.enable_kmer_table(0)
    monkeypatch.setenv("RB_PIECE_MB", "1")                  # ~10 pieces: a "large" batch
    monkeypatch.setenv("RB_HOST_PACK", "1")                 # pinned choice: no measurement; uploads the thresholds once
    assert_same_results(gf.count_batch(bases, off, lut), exp, dense=False)
    monkeypatch.delenv("RB_HOST_PACK", raising=False)
    assert gf.transfer_policy()["choice"] == "undecided"
    x = [rb.transfer_bytes()[0]]
    for call in range(6):
        assert_same_results(gf.count_batch(bases, off, lut), exp, dense=False)
        x.append(rb.transfer_bytes()[0])
    moved = np.diff(x)
    assert moved[0] == moved[1] < moved[2] == moved[3]      # packed, packed, ASCII, ASCII
    pol = gf.transfer_policy()
    assert pol["choice"] in ("packed", "ascii") and pol["ns_per_base_packed"] > 0 and pol["ns_per_base_ascii"] > 0
    assert moved[4] == moved[5] == (moved[0] if pol["choice"] == "packed" else moved[2])
    monkeypatch.setenv("RB_PIECE_MB", "64")                 # one piece: small batches are always packed
    x0 = rb.transfer_bytes()[0]
    assert_same_results(gf.count_batch(bases, off, lut), exp, dense=False)
    assert abs(int(rb.transfer_bytes()[0] - x0) - int(moved[0])) < moved[0] // 100      # planes, fewer per-piece paddings


def test_create_shard_builds_a_column_slice_in_place():
    """rb_ibf_create_shard + inserts with global bin ids (bins of other shards are skipped) == the column slice of the
    whole filter, for every shard; the shards' keys combine (MAX) to the whole filter's keys."""
    plan, of, gf = make_filter_pair(1, 700 * 2000 + 7, 2000, 13)            # 701 bins, 11 row words
    whole = gf.download()
    bases, off = synth.ragged_reads(plan["bases"], [250] * 300 + [0, 5, 1000], seed=9, frac_from_ref=0.7, n_frac=0.002)
    lut = rb.threshold_lut(0.1, 13)
    exp = of.count_batch(bases, off, lut, n_threads=4)
    combined = None
    for s_ in range(3):
        sh = rb.IBF.create_shard(plan["n_bins"], 3, 13, plan["n_bits"], s_, 3)
        sh.insert_batch(plan["bases"], plan["frag_begin"], plan["frag_end"], plan["frag_bin"])
        ref = rb.IBF.from_words(whole, plan["n_bins"], 3, 13, plan["n_bits"], shard=s_, n_shards=3)
        assert (sh.col_begin, sh.col_words, sh.bin_begin, sh.n_bins_local) == (ref.col_begin, ref.col_words, ref.bin_begin, ref.n_bins_local)
        assert np.array_equal(sh.download(), ref.download())
        got = sh.count_batch(bases, off, lut)
        k_s = (got["hit"].astype(np.uint64) << np.uint64(48)) | (got["max_count"].astype(np.uint64) << np.uint64(32)) | \
              np.where(got["hit"] > 0, (~got["argmax_bin"]).astype(np.uint64) & np.uint64(0xFFFFFFFF), np.uint64(0))
        combined = k_s if combined is None else np.maximum(combined, k_s)
    mx, hit, am = rb.keys_decode(combined)
    assert np.array_equal(mx, exp["max_count"]) and np.array_equal(hit, exp["hit"]) and np.array_equal(am, exp["argmax_bin"])


@pytest.mark.parametrize("n_shards,tables", [(2, ""), (3, "lists"), (5, "lists"), (3, "slots"), (2, "ctable"), (5, "ctable")])
def test_sharded_call_folds_keys_into_one_array(n_shards, tables, monkeypatch):
    """rb_ibf_count_batch_sharded: every shard's count kernel folds its keys into shard 0's key array (atomicMax; over
    NVLink when the shards sit on different devices -- here round-robin over the visible ones), == whole filter == oracle.
    Without tables the streaming kernel runs, with tables the postings kernel of either layout; narrow shards take the
    hashed-probe kernel."""
    plan, of, gf = make_filter_pair(1, 700 * 2000 + 7, 2000, 13)            # 701 bins, 11 row words
    bases, off = synth.ragged_reads(plan["bases"], [250] * 300 + [0, 5, 12, 13, 400, 1000], seed=9, frac_from_ref=0.7, n_frac=0.002)
    luts = np.stack([rb.threshold_lut(0.1, 13), rb.threshold_lut(0.08, 13)])
    n_dev = rb.device_count()
    shards = []
    for s_ in range(n_shards):
        sh = rb.IBF.create_shard(plan["n_bins"], 3, 13, plan["n_bits"], s_, n_shards, device=s_ % n_dev)
        sh.insert_batch(plan["bases"], plan["frag_begin"], plan["frag_end"], plan["frag_bin"])
        shards.append(sh)
    if tables == "ctable":                              # shards of 3..16 row words: the group-loaded k-mer table folds its keys too
        rb.enable_kmer_tables(shards)
        assert [s.kmer_table_kind() for s in shards] == [4 if s.col_words >= 3 else 1 for s in shards]
    elif tables:
        monkeypatch.setenv("RB_POSTINGS_LAYOUT", tables)
        monkeypatch.setenv("RB_CTABLE", "0")
        rb.enable_kmer_tables(shards)
        assert [s.kmer_table_kind() for s in shards] == [(2 if tables == "lists" else 3) if s.col_words > 4 else 1 for s in shards]
    got = rb.count_batch_sharded(shards, bases, off, luts)
    whole = gf.count_batch(bases, off, luts)
    for t in range(2):
        exp = of.count_batch(bases, off, luts[t], dense=False, n_threads=4)
        for key in ("max_count", "hit", "argmax_bin"):
            assert np.array_equal(got[key][t], exp[key]), (t, key)
            assert np.array_equal(whole[key][t], exp[key])
    assert np.array_equal(got["read_flag"], exp["short_read"])


def test_two_threshold_tables_in_one_pass():
    plan, of, gf = make_filter_pair(100, 20000, 21000, 13)
    bases, off, _ = synth.sample_reads(plan["bases"], 3000, 250, seed=5, error_rate=0.12)
    luts = np.stack([rb.threshold_lut(0.1, 13), rb.threshold_lut(0.08, 13)])
    got = gf.count_batch(bases, off, luts)
    for t in range(2):
        exp = of.count_batch(bases, off, luts[t], dense=False, n_threads=4)
        assert np.array_equal(got["max_count"][t], exp["max_count"])
        assert np.array_equal(got["hit"][t], exp["hit"])
        assert np.array_equal(got["argmax_bin"][t], exp["argmax_bin"])
    assert (got["hit"][0] != got["hit"][1]).any()


def test_long_read_too_long_flag():
    plan, of, gf = make_filter_pair(1, 300000, 100000, 13)
    bases, off = synth.ragged_reads(plan["bases"], [70000, 65535, 300], seed=3, frac_from_ref=1.0)
    lut = rb.threshold_lut(0.1, 13)
    got = gf.count_batch(bases, off, lut, dense=True)
    exp = of.count_batch(bases, off, lut)
    assert got["read_flag"].tolist() == [2, 0, 0]
    assert_same_results(got, exp)


@pytest.mark.parametrize("kernel", ["tile", "stream"])
@pytest.mark.parametrize("n_shards", [2, 3])
def test_bin_sharded_counts_and_key_combine(kernel, n_shards):
    """Column-sliced handles: dense counts equal the oracle's bin range; MAX over shard keys equals the whole filter."""
    rb.set_count_kernel(KERNELS[kernel])
    plan, of, gf = make_filter_pair(1, 700 * 2000 + 7, 2000, 13)
    words = gf.download()
    bases, off = synth.ragged_reads(plan["bases"], [250] * 60 + [13, 5, 1200], seed=11, frac_from_ref=0.8)
    lut = rb.threshold_lut(0.1, 13)
    exp = of.count_batch(bases, off, lut)
    keys = np.zeros(len(off) - 1, np.uint64)
    for s in range(n_shards):
        sh = rb.IBF.from_words(words, plan["n_bins"], 3, 13, plan["n_bits"], shard=s, n_shards=n_shards)
        assert sh.col_begin == gf.bin_width * s // n_shards
        got = sh.count_batch(bases, off, lut, dense=True)
        lo, hi = sh.bin_begin, sh.bin_begin + sh.n_bins_local
        assert np.array_equal(got["counts_fwd"], exp["counts_fwd"][:, lo:hi])
        assert np.array_equal(got["counts_rev"], exp["counts_rev"][:, lo:hi])
        k_s = (got["hit"].astype(np.uint64) << np.uint64(48)) | (got["max_count"].astype(np.uint64) << np.uint64(32)) | \
              np.where(got["hit"] > 0, (~got["argmax_bin"]).astype(np.uint64) & np.uint64(0xFFFFFFFF), np.uint64(0))
        keys = np.maximum(keys, k_s)
    mx, hit, am = rb.keys_decode(keys)
    assert np.array_equal(mx, exp["max_count"]) and np.array_equal(hit, exp["hit"]) and np.array_equal(am, exp["argmax_bin"])


def test_sharded_load_from_file(tmp_path):
    plan, of, gf = make_filter_pair(1, 300 * 2000 + 7, 2000, 13)
    p = tmp_path / "wide.ibf"
    gf.store(p)
    assert oracle.OracleIBF.load(p).n_bins == plan["n_bins"]
    full = gf.download().reshape(-1)[:gf.n_blocks * gf.bin_width].reshape(gf.n_blocks, gf.bin_width)
    for s in range(2):
        sh = rb.IBF.load(p, shard=s, n_shards=2)
        loc = sh.download().reshape(sh.n_blocks, sh.col_words)
        assert np.array_equal(loc, full[:, sh.col_begin:sh.col_begin + sh.col_words])


def test_generic_hash_count_path():
    """n_hash != 3 takes the runtime-n_hash tile path (file format allows it; the reference fixes 3)."""
    plan, of, gf = make_filter_pair(20, 5000, 6000, 13, n_hash=2)
    bases, off = synth.ragged_reads(plan["bases"], [250] * 30 + [40, 9], seed=2, frac_from_ref=0.8)
    lut = rb.threshold_lut(0.1, 13)
    assert_same_results(gf.count_batch(bases, off, lut, dense=True), of.count_batch(bases, off, lut))


# ---- insert ---------------------------------------------------------------------------------------------
def test_insert_reports_out_of_range_bin():
    """Quirk Q3: a sequence with len mod F in (F-k+2, F-1] consumes one bin id too many."""
    ref = [synth.random_bases(199991, 1), synth.random_bases(50001, 2)]     # post-N-cut: 199990, 50000
    plan = synth.build_plan(ref, 100000, 13)
    assert plan["bin_ids_consumed"] == plan["n_bins"] + 1
    gf = rb.IBF.create(plan["n_bins"], 3, 13, plan["n_bits"])
    with pytest.raises(rb.RBError) as e:
        gf.insert_batch(plan["bases"], plan["frag_begin"], plan["frag_end"], plan["frag_bin"])
    assert e.value.status == 7
    of = oracle.OracleIBF.create(plan["n_bins"], 3, 13, plan["n_bits"])
    of.insert_batch(plan["bases"], plan["frag_begin"], plan["frag_end"], plan["frag_bin"])
    assert np.array_equal(gf.download(), of.words()[:plan["n_bits"] // 64])     # in-range fragments still inserted


def test_insert_is_idempotent_and_order_free():
    plan, of, gf = make_filter_pair(3, 250000, 100000, 13)
    before = gf.download()
    perm = np.random.default_rng(0).permutation(len(plan["frag_bin"]))
    gf.insert_batch(plan["bases"], plan["frag_begin"][perm], plan["frag_end"][perm], plan["frag_bin"][perm])
    assert np.array_equal(gf.download(), before)


# ---- column build (shared-memory bit columns + bit-tile transpose) vs the RED.OR kernel and the oracle ----------
def _messy_reference(n_seqs, seq_len, seed):
    """Sequences with N runs (cut by cutOutNNNs), lower-case bases and IUPAC codes (rank 4 k-mers are hashed too)."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n_seqs):
        s = synth.random_bases(seq_len + int(rng.integers(0, 50)), seed + i).copy()
        for _ in range(3):
            a = int(rng.integers(0, len(s) - 40))
            s[a:a + int(rng.integers(1, 30))] = ord("N")
        low = rng.integers(0, len(s), 200)
        s[low] |= 0x20
        s[rng.integers(0, len(s), 20)] = ord("R")
        out.append(s)
    return out


@pytest.mark.parametrize("n_seqs,seq_len,frag,k,scratch_mb", [
    (3, 250_000, 100_000, 13, 0),        # the reference's default sizing: 1 236 269 rows, 154 KB columns, one word per row
    (1, 700 * 2000 + 7, 2000, 13, 0),    # 701 bins (two 512-bin groups, partial last word), 24 785 rows
    (1, 1300 * 2000 + 7, 2000, 15, 2),   # three groups, scratch capped at 2 MB: one 512-bin group per pass
    (40, 3000, 100_000, 11, 0),          # one short fragment per bin, n_hash = 3, rows not a multiple of 1024
    (3, 450_000, 200_000, 13, 0),        # 2 472 526 rows: the column (309 KB) no longer fits shared memory -> scratch columns in L2
    (70, 30_000, 1_000_000, 13, 0),      # 12.4 M rows, 70 bins (two words per row), 1.5 MB columns through the L2 path
])
def test_column_build_equals_red_kernel_and_oracle(monkeypatch, n_seqs, seq_len, frag, k, scratch_mb):
    if scratch_mb:
        monkeypatch.setenv("RB_INSERT_SCRATCH_MB", str(scratch_mb))
    ref = _messy_reference(n_seqs, seq_len, 300)
    plan = synth.build_plan(ref, frag, k)
    of = oracle.OracleIBF.create(plan["n_bins"], 3, k, plan["n_bits"])
    of.insert_batch(plan["bases"], plan["frag_begin"], plan["frag_end"], plan["frag_bin"], n_threads=4)
    exp = of.words()[:plan["n_bits"] // 64]
    got = {}
    for variant in (1, 2):
        rb.set_insert_kernel(variant)
        gf = rb.IBF.create(plan["n_bins"], 3, k, plan["n_bits"])
        launches0 = rb.kernel_launches()
        gf.insert_batch(plan["bases"], plan["frag_begin"], plan["frag_end"], plan["frag_bin"])
        got[variant] = rb.kernel_launches() - launches0
        assert np.array_equal(gf.download(), exp), variant
        gf.close()
    assert got[1] == 1 and got[2] >= 5           # the column path really ran (3 list kernels + build + merge per pass)
    if scratch_mb:
        assert got[2] == 3 + 2 * 3


def test_column_build_merges_into_existing_bits_shards_and_bad_bins():
    """Second insert ORs into what is there; several fragments per bin; fragments shorter than k; a bin past the end
    raises InsertSequenceException but the rest is written; bin shards take only their own columns."""
    rb.set_insert_kernel(2)
    ref = _messy_reference(2, 150 * 2000, 301)
    plan = synth.build_plan(ref, 2000, 13)
    n = len(plan["frag_bin"])
    rng = np.random.default_rng(5)
    fbin = plan["frag_bin"].copy()
    fbin[rng.integers(0, n, 40)] = rng.integers(0, plan["n_bins"], 40)       # crowd some bins
    fb = np.concatenate([plan["frag_begin"], [10, 500]]).astype(np.uint64)
    fe = np.concatenate([plan["frag_end"], [10 + 12, 500]]).astype(np.uint64)       # 12 bases (< k) and empty
    fbin = np.concatenate([fbin, [3, 4]]).astype(np.uint64)
    of = oracle.OracleIBF.create(plan["n_bins"], 3, 13, plan["n_bits"])
    gf = rb.IBF.create(plan["n_bins"], 3, 13, plan["n_bits"])
    half = n // 2
    for sl in (slice(0, half), slice(half, None)):
        of.insert_batch(plan["bases"], fb[sl], fe[sl], fbin[sl])
        gf.insert_batch(plan["bases"], fb[sl], fe[sl], fbin[sl])
    exp = of.words()[:plan["n_bits"] // 64]
    assert np.array_equal(gf.download(), exp)
    # out-of-range bin: reported, everything else still inserted
    g2 = rb.IBF.create(plan["n_bins"], 3, 13, plan["n_bits"])
    bad = fbin.copy()
    bad[7] = plan["n_bins"]
    with pytest.raises(rb.RBError) as e:
        g2.insert_batch(plan["bases"], fb, fe, bad)
    assert e.value.status == 7
    o2 = oracle.OracleIBF.create(plan["n_bins"], 3, 13, plan["n_bits"])
    keep = np.arange(len(bad)) != 7
    o2.insert_batch(plan["bases"], fb[keep], fe[keep], bad[keep])
    assert np.array_equal(g2.download(), o2.words()[:plan["n_bits"] // 64])
    # bin shards
    full = exp.reshape(gf.n_blocks, gf.bin_width)
    zeros = np.zeros(plan["n_bits"] // 64, np.uint64)
    for s in range(3):
        sh = rb.IBF.from_words(zeros, plan["n_bins"], 3, 13, plan["n_bits"], shard=s, n_shards=3)
        sh.insert_batch(plan["bases"], fb, fe, fbin)
        loc = sh.download().reshape(sh.n_blocks, sh.col_words)
        assert np.array_equal(loc, full[:, sh.col_begin:sh.col_begin + sh.col_words]), s


def test_golden_rebuild_with_column_build(known, golden_sparse):
    """The reference's own .ibf fixtures, rebuilt through the column path (two fragments per bin pair)."""
    rb.set_insert_kernel(2)
    for name, fasta, k in (("lib_test", "lib_test.fasta", 13), ("lib_test1", "lib_test1.fasta", 13),
                           ("classify_test", "classify_test.fasta", 15)):
        seqs = [s for _, s in read_fasta(data_path(fasta))]
        plan = synth.build_plan(seqs + seqs, 100000, k)
        meta, words = golden_words(known, golden_sparse, name)
        f = rb.IBF.create(plan["n_bins"], 3, k, plan["n_bits"])
        f.insert_batch(plan["bases"], plan["frag_begin"], plan["frag_end"], plan["frag_bin"])
        assert np.array_equal(f.download(), words[:plan["n_bits"] // 64]), name


# ---- BASELINE config #1: usage=classify on testData/testQueries.fasta vs an IBF of a synthetic 5 Mb reference ---
def test_config1_testqueries_vs_5mb_reference():
    ref = [synth.random_bases(5_000_001, 1)]                     # 5 000 000 after the N-cut quirk -> 51 bins
    plan = synth.build_plan(ref, 100000, 13)
    assert plan["n_bins"] == 51 and plan["n_bits"] == 79121216 and plan["bin_ids_consumed"] == 51
    gf = rb.IBF.create(plan["n_bins"], 3, 13, plan["n_bits"])
    gf.insert_batch(plan["bases"], plan["frag_begin"], plan["frag_end"], plan["frag_bin"])
    of = oracle.OracleIBF.create(plan["n_bins"], 3, 13, plan["n_bits"])
    of.insert_batch(plan["bases"], plan["frag_begin"], plan["frag_end"], plan["frag_bin"], n_threads=8)
    assert np.array_equal(gf.download(), of.words()[:plan["n_bits"] // 64])
    name, query = read_fasta(data_path("testQueries.fasta"))[0]
    assert name == "1" and len(query) == 1890
    # the chunk schedule of classify_reads (classify.hpp:262-299): max_chunks=5 chunks of 250, plus
    # reference-derived chunks so that the batch has hits, plus 10 000 synthetic chunks
    chunks = [query[i * 250:(i + 1) * 250] for i in range(5)]
    syn_b, syn_o, _ = synth.sample_reads(plan["bases"], 10000, 250, seed=1234)
    bases = np.concatenate([np.frombuffer(b"".join(chunks), np.uint8), syn_b])
    off = np.concatenate([np.arange(6, dtype=np.uint64) * np.uint64(250), syn_o[1:] + np.uint64(1250)])
    luts = np.stack([rb.threshold_lut(0.1, 13), rb.threshold_lut(0.08, 13)])
    for which in (1, 2, 3):
        rb.set_count_kernel(which)
        got = gf.count_batch(bases, off, luts, dense=True)
        for t in range(2):
            exp = of.count_batch(bases, off, luts[t], n_threads=8)
            assert np.array_equal(got["counts_fwd"], exp["counts_fwd"]) and np.array_equal(got["counts_rev"], exp["counts_rev"])
            for key in ("max_count", "hit", "argmax_bin"):
                assert np.array_equal(got[key][t], exp[key]), (which, t, key)
    assert got["hit"][0][:5].sum() == 0                          # the human-like query does not match a random reference
    assert 0.45 < got["hit"][0][5:].mean() < 0.55                # half of the synthetic chunks come from the reference


# ---- BASELINE config #2 at full size: size-independent properties + sampled oracle check ----------------
def test_config2_full_size_properties():
    ref = [synth.random_bases(4_000_000, 2 + i) for i in range(100)]
    plan = synth.build_plan(ref, 4_200_000, 13)
    assert plan["n_bins"] == 100
    gf = rb.IBF.create(plan["n_bins"], 3, 13, plan["n_bits"])
    gf.insert_batch(plan["bases"], plan["frag_begin"], plan["frag_end"], plan["frag_bin"])
    n = 1_000_000
    bases, off, from_ref = synth.sample_reads(plan["bases"], n, 250, seed=1234)
    lut = rb.threshold_lut(0.1, 13)
    assert lut[250] == 18
    res = gf.count_batch(bases, off, lut)
    # (1) reverse-complementing every read swaps the strands: identical summaries
    rc = synth._COMP[bases.reshape(n, 250)[:, ::-1]].reshape(-1)
    res_rc = gf.count_batch(rc, off, lut)
    for key in ("max_count", "hit", "argmax_bin"):
        assert np.array_equal(res[key], res_rc[key])
    # (2) all three kernels agree on the whole batch (auto = direct k-mer table at this size)
    assert (gf.kmer_table_bytes(), gf.kmer_table_span()) == (wtable_bytes(13, 2, 3), 3)   # 64 GiB, canonical 15-mers
    for which in (1, 2, 4):
        rb.set_count_kernel(which)
        res_s = gf.count_batch(bases, off, lut)
        for key in ("max_count", "hit", "argmax_bin", "read_flag"):
            assert np.array_equal(res[key], res_s[key]), (which, key)
    rb.set_count_kernel(0)
    # (3) sanity of the classifier on the synthetic mix
    # (k=13 has only 4^13 = 67 M k-mers, so 4 Mb bins also contain ~6 % of any random read's k-mers and
    #  iid reads pass thr 18 too -- a property of the configuration, reproduced by the oracle below)
    assert res["hit"][from_ref].mean() > 0.97
    assert res["max_count"][from_ref].mean() > res["max_count"][~from_ref].mean() + 20
    # (4) oracle on a random sample of the full-size batch, against the downloaded filter
    of = oracle.OracleIBF.create(plan["n_bins"], 3, 13, plan["n_bits"])
    of.words()[:plan["n_bits"] // 64] = gf.download()
    pick = np.random.default_rng(9).choice(n, 3000, replace=False)
    sb = bases.reshape(n, 250)[pick].reshape(-1)
    so = np.arange(len(pick) + 1, dtype=np.uint64) * np.uint64(250)
    exp = of.count_batch(sb, so, lut, dense=False, n_threads=8)
    assert np.array_equal(res["max_count"][pick], exp["max_count"])
    assert np.array_equal(res["hit"][pick], exp["hit"])
    assert np.array_equal(res["argmax_bin"][pick], exp["argmax_bin"])


# ---- update_filter / resizeBins (SeqAn's resizeBins itself is unpinned by reference fixtures) ----
@pytest.mark.parametrize("old_seqs,new_seqs", [(3, 2), (60, 10), (64, 1), (100, 100)])
def test_resize_bins_and_append_equals_oracle(old_seqs, new_seqs):
    """Appending bins keeps every existing bit at its (row, bin) and widens rows when 64 is crossed."""
    plan, of, gf = make_filter_pair(old_seqs, 3000, 4000, 13)
    rows = gf.n_blocks
    total = old_seqs + new_seqs
    gf.resize_bins(total)
    of.resize_bins(total)
    assert (gf.n_bins, gf.n_blocks, gf.bin_width) == (total, rows, (total + 63) // 64) == (of.n_bins, of.n_blocks, of.bin_width)
    assert gf.n_bits == of.n_bits == rows * 64 * ((total + 63) // 64)
    assert np.array_equal(gf.download(), of.words()[:of.n_bits // 64])
    extra = [synth.random_bases(3000, 4000 + i) for i in range(new_seqs)]
    p2 = synth.build_plan(extra, 4000, 13)
    bins = p2["frag_bin"] + np.uint64(old_seqs)
    gf.insert_batch(p2["bases"], p2["frag_begin"], p2["frag_end"], bins)
    of.insert_batch(p2["bases"], p2["frag_begin"], p2["frag_end"], bins)
    assert np.array_equal(gf.download(), of.words()[:of.n_bits // 64])
    b_old, o_old = synth.ragged_reads(plan["bases"], [250] * 40, seed=3, frac_from_ref=1.0)
    b_new, o_new = synth.ragged_reads(p2["bases"], [250] * 40, seed=4, frac_from_ref=1.0)
    bases, off = np.concatenate([b_old, b_new]), np.concatenate([o_old, o_new[1:] + o_old[-1]])
    lut = rb.threshold_lut(0.1, 13)
    exp = of.count_batch(bases, off, lut)
    assert_same_results(gf.count_batch(bases, off, lut, dense=True), exp)
    assert (exp["argmax_bin"][exp["hit"] > 0] >= old_seqs).any() and (exp["argmax_bin"][exp["hit"] > 0] < old_seqs).any()
    with pytest.raises(rb.RBError):
        gf.resize_bins(total - 1)


@pytest.mark.parametrize("workload,extra", [
    ("mini_100x60kb_100bins", []),
    ("mini5_40Mb_512bins_per_gpu", []),            # config #5 shape: own genome group, own column slice (rb_ibf_create_shard)
])
def test_bench_line_on_mini_workloads(workload, extra):
    """bench.py end to end on small workloads: one JSON line with the contract's keys; its own asserts (GPU == oracle on
    the batch, host-API == device-API, argmax bins inside the sampling windows) must hold."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--workload", workload, "--steps", "3", "--warmup", "3"] + extra,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["value"] > 0 and d["gpu_launches"] > 0 and d["roofline"]["frac"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    if "per_gpu" not in workload:
        assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0
