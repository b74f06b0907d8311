"""Pins the CPU oracle to the reference's own golden fixtures and known answers
(SURVEY.md section 8c / Appendix B).  CPU only."""
import hashlib
import os

import numpy as np
import pytest

import oracle
from conftest import data_path, golden_file_bytes, golden_words, read_fasta

REF = "/root/reference"


def test_sparse_fixtures_match_reference_files_when_present(known, golden_sparse):
    """In the build container the sparse fixtures must reproduce the raw reference files."""
    if not os.path.isdir(REF):
        pytest.skip("reference tree not present (GPU box)")
    for name, meta in known["ibf"].items():
        raw = open(os.path.join(REF, meta["source"]), "rb").read()
        assert golden_file_bytes(known, golden_sparse, name) == raw


def test_load_header_and_metadata(known, golden_ibf_paths):
    exp = {"lib_test": (2, 3, 13, 79121216), "lib_test1": (4, 3, 13, 79121216),
           "classify_test": (2, 3, 15, 79119680)}
    for name, path in golden_ibf_paths.items():
        f = oracle.OracleIBF.load(path)
        assert (f.n_bins, f.n_hash, f.k, f.n_bits) == exp[name]
        assert known["ibf"][name]["tail"] == [f.n_bins, f.n_hash, f.k, f.k]
        assert f.bin_width == 1 and f.n_blocks == f.n_bits // 64


def test_store_roundtrip_is_byte_identical(known, golden_ibf_paths, tmp_path):
    for name, path in golden_ibf_paths.items():
        f = oracle.OracleIBF.load(path)
        out = tmp_path / (name + ".out.ibf")
        f.store(out)
        assert hashlib.md5(out.read_bytes()).hexdigest() == known["ibf"][name]["md5"]


def test_load_rejects_non_ibf(tmp_path):
    """configReader.cpp:210-224 sniffs FASTA-vs-IBF by a failing retrieve()."""
    with pytest.raises(oracle.OracleError) as e:
        oracle.OracleIBF.load(data_path("lib_test.fasta"))
    assert e.value.status == 4
    with pytest.raises(oracle.OracleError) as e:
        oracle.OracleIBF.load(tmp_path / "missing.ibf")
    assert e.value.status == 5


@pytest.mark.parametrize("name,fasta,k", [("lib_test", "lib_test.fasta", 13),
                                          ("lib_test1", "lib_test1.fasta", 13),
                                          ("classify_test", "classify_test.fasta", 15)])
def test_rebuild_from_fasta_is_byte_identical(known, golden_sparse, name, fasta, k):
    """create_filter (test build: queue replayed twice) reproduces each golden .ibf bit for bit."""
    seqs = [s for _, s in read_fasta(data_path(fasta))]
    f, stats = oracle.build_from_sequences(seqs, fragment_length=100000, k=k, passes=2)
    meta, words = golden_words(known, golden_sparse, name)
    assert f.n_bits + 256 == meta["bit_length"]
    assert np.array_equal(f.words(), words)
    assert stats["dropped_fragments"] == 0


def test_create_filter_stats(known):
    ka = known["known"]
    seqs = [s for _, s in read_fasta(data_path("lib_test.fasta"))]
    f, stats = oracle.build_from_sequences(seqs, 100000, k=13, passes=2)
    assert stats["totalBinsBinId"] == 2 and stats["sumSeqLen"] == 144          # createfilter.hpp:136-137
    assert stats["filter_size_bits"] == ka["filter_size_bits"]["value"]        # createfilter.hpp:148
    seqs = [s for _, s in read_fasta(data_path("lib_test1.fasta"))]
    f, stats = oracle.build_from_sequences(seqs, 100000, k=13, passes=2)
    assert stats["sumSeqLen"] == ka["filter_stats_test1"]["sumSeqLen"]         # createfilter.hpp:218
    assert stats["totalBinsBinId"] == ka["filter_stats_test1"]["totalBinsBinId"]


def test_cut_out_nnns(known):
    ka = known["known"]["cut_out_nnns"]
    assert oracle.cut_out_nnns(ka["in"]) == ka["out"].encode()
    assert oracle.cut_out_nnns("ACGTN") == b"ACGT"         # trailing N: nothing dropped
    assert oracle.cut_out_nnns("NNACGT") == b"ACG"         # no trailing N: last base dropped (Q1)
    assert oracle.cut_out_nnns("NNNN") == b""
    assert oracle.cut_out_nnns("ACnGT") == b"ACnG"         # lowercase n is not cut


def test_filter_size_bits(known):
    ka = known["known"]["filter_size_bits"]
    assert oracle.filter_size_bits(100000, 13, 3, 0.01, 2) == ka["value"]
    assert oracle.filter_size_bits(100000, 13, 3, 0.01, 2) == ka["bin_size_bits"] * 64
    assert oracle.filter_size_bits(100000, 13, 3, 0.01, 64) == ka["bin_size_bits"] * 128   # A.7: 64 bins -> 128
    assert oracle.filter_size_bits(100000, 15, 3, 0.01, 2) == 79119680


def test_fragment_schedule(known):
    b, e = oracle.fragment_schedule(72, 100000, 13)
    assert list(b) == [0] and list(e) == [72]                                  # createfilter.hpp:168-171
    b, e = oracle.fragment_schedule(250000, 100000, 13)
    assert list(b) == [0, 99988, 199988] and list(e) == [100000, 200000, 250000]
    # quirk Q3: len mod F in (F-k+2, F-1] consumes an extra (< k bases) fragment
    # (199990, 199991, 299999 overrun their len/F+1 bins; 199989, 250000, 300000 do not)
    for n, frags, overrun in [(199989, 2, False), (199990, 3, True), (199991, 3, True), (299999, 4, True),
                              (300000, 4, False), (250000, 3, False), (5000000, 51, False), (4999999, 51, True)]:
        b, e = oracle.fragment_schedule(n, 100000, 13)
        assert len(b) == frags, n
        assert (len(b) > oracle.bins_for_sequence(n, 100000)) == overrun, n
    assert len(oracle.fragment_schedule(1, 100000, 13)[0]) == 0
    assert len(oracle.fragment_schedule(0, 100000, 13)[0]) == 0


def test_calculate_ci_and_threshold(known):
    ka = known["known"]["ci_0.1_13_35_0.95"]
    assert oracle.calculate_ci(0.1, 13, 35, 0.95) == (ka["low"], ka["high"])       # read.hpp:156-157
    assert oracle.threshold(0.1, 13, 35) == (ka["threshold_int16"] & 0xFFFF)       # read.hpp:164 -> 65529
    # SURVEY A.8 reference values
    assert oracle.calculate_ci(0.1, 13, 250, 0.95) == (135, 220) and oracle.threshold(0.1, 13, 250) == 18
    assert oracle.calculate_ci(0.08, 13, 250, 0.95) == (111, 204) and oracle.threshold(0.08, 13, 250) == 34
    assert oracle.calculate_ci(0.1, 13, 354, 0.95) == (205, 306) and oracle.threshold(0.1, 13, 354) == 36
    assert oracle.threshold(0.1, 15, 250) == 8
    assert oracle.calculate_ci(0.1, 15, 360, 0.95) == (225, 324) and oracle.threshold(0.1, 15, 360) == 22
    lut = oracle.threshold_lut(0.1, 13)
    assert lut[250] == 18 and lut[35] == 65529 and lut[354] == 36


def test_known_counts_35mer(known, golden_ibf_paths):
    ka = known["known"]
    f = oracle.OracleIBF.load(golden_ibf_paths["lib_test"])
    fwd = f.count(ka["read35"]["seq"])
    rev = f.count(ka["read35"]["seq"], revcomp=True)
    assert list(fwd) == [23, 23] and list(rev) == [0, 0]
    # CountMatchesTest feeds the reverse complement: fwd 0, rev 23 (read.hpp:286-327)
    rc = ka["read35_revcomp"]["seq"]
    assert list(f.count(rc)) == [0, 0] and list(f.count(rc, revcomp=True)) == [23, 23]
    # with the production uint16 threshold (65529) nothing can match (quirk Q6)
    assert f.count_matches(ka["read35"]["seq"]) == 0


def test_known_count_matches_354(known, golden_ibf_paths):
    ka = known["known"]
    read = ka["read354"]["seq"]
    assert len(read) == ka["count_matches_354"]["readlen"]
    f0 = oracle.OracleIBF.load(golden_ibf_paths["lib_test"])
    f1 = oracle.OracleIBF.load(golden_ibf_paths["lib_test1"])
    assert f0.count_matches(read) == ka["count_matches_354"]["lib_test"]          # read.hpp:221-229
    assert f1.count_matches(read) == ka["count_matches_354"]["lib_test1"]
    assert list(f0.count(read, revcomp=True)) == [0, 0]
    assert oracle.classify_any([f0, f1], read) is True                             # read.hpp:202
    assert oracle.classify_best([f0, f1], read) == ka["count_matches_354"]["best_index"]   # read.hpp:231
    assert list(oracle.classify_pair([f0], [f1], read)) == ka["count_matches_354"]["pair"]  # read.hpp:250


def test_classify_exceptions(golden_ibf_paths):
    f0 = oracle.OracleIBF.load(golden_ibf_paths["lib_test"])
    for fn in (oracle.classify_any, oracle.classify_best):
        with pytest.raises(oracle.OracleError) as e:
            fn([], "ACGTACGTACGTACGT")
        assert e.value.status == 1                                                 # NullFilterException
        with pytest.raises(oracle.OracleError) as e:
            fn([f0], "ACGT")
        assert e.value.status == 2                                                 # ShortReadException
    with pytest.raises(oracle.OracleError) as e:
        oracle.classify_pair([], [f0], "ACGTACGTACGTACGT")
    assert e.value.status == 1
    assert oracle.classify_pair([f0], [f0], "ACGT") == (0, 0)                      # k > len skipped silently


def test_classify_reads_fixture(known, golden_ibf_paths):
    """classifyTests: 3/3 reads found (classifygtests.hpp:70-79); SURVEY Appendix B chunk details."""
    f = oracle.OracleIBF.load(golden_ibf_paths["classify_test"])
    reads = read_fasta(data_path("classify_test.fastq"))
    assert [len(s) for _, s in reads] == [1628, 8177, 17298]

    def run(chunk_length, max_chunks):
        found, first_chunk = 0, []
        for _, seq in reads:
            hit = -1
            for i in range(max_chunks):
                frag = seq[i * chunk_length: min((i + 1) * chunk_length, len(seq))]
                if oracle.classify_best([f], frag) != -1:      # target-only branch, classify.hpp:284
                    hit = i
                    break
            first_chunk.append(hit)
            found += hit >= 0
        return found, first_chunk

    assert run(250, 5) == (3, [0, 0, 0])
    assert run(360, 5) == (3, [0, 0, 3])
    assert run(360, 3)[0] == 2
    c250 = [int(f.count(seq[:250]).max()) for _, seq in reads]
    c360 = [int(f.count(seq[:360]).max()) for _, seq in reads]
    assert c250 == [236, 56, 16] and c360 == [346, 78, 16]


def test_check_unblock_table(known, golden_ibf_paths):
    read = known["known"]["read354"]["seq"]
    f0 = oracle.OracleIBF.load(golden_ibf_paths["lib_test"])
    f1 = oracle.OracleIBF.load(golden_ibf_paths["lib_test1"])
    rnd = "ACGTTGCATGCCGATAGCTAGCTAGGATCGATCGATTAGCGGCTATATCGCGATATCGGCTAGCTAGCTAGGCTCTAGAGAGCTCGCGATATAGC" * 3
    assert oracle.check_unblock([f0], [], read) == 1          # deplete-only hit -> unblock
    assert oracle.check_unblock([f0], [], rnd) == 0
    assert oracle.check_unblock([], [f0], read) == 2          # target-only hit -> stop_further_data
    assert oracle.check_unblock([], [f0], rnd) == 1
    assert oracle.check_unblock([f0], [f1], read) == 0        # both match, also at error_rate-0.02 -> keep
    assert oracle.check_unblock([f0], [f1], rnd) == 0


def test_dna5_table():
    for c, d in zip("ACGTacgtUuNnRYKM-*", [0, 1, 2, 3, 0, 1, 2, 3, 3, 3] + [4] * 8):
        assert oracle.dna5(c) == d
    assert oracle.kmer_hash("ACGTN", 5) == ((((0 * 5 + 1) * 5 + 2) * 5 + 3) * 5 + 4)


def test_count_batch_matches_single_calls(golden_ibf_paths, known):
    f = oracle.OracleIBF.load(golden_ibf_paths["lib_test1"])
    reads = [known["known"]["read354"]["seq"].encode(), b"ACGT", known["known"]["read35"]["seq"].encode(), b""]
    off = np.cumsum([0] + [len(r) for r in reads]).astype(np.uint64)
    bases = np.frombuffer(b"".join(reads), np.uint8)
    lut = oracle.threshold_lut(0.1, 13)
    for nt in (1, 3):
        res = f.count_batch(bases, off, lut, n_threads=nt)
        assert list(res["short_read"]) == [0, 1, 0, 1]
        assert list(res["max_count"]) == [182, 0, 0, 0]
        assert list(res["hit"]) == [1, 0, 0, 0]
        assert res["argmax_bin"][0] == 0 and res["argmax_bin"][1] == 0xFFFFFFFF
        assert np.array_equal(res["counts_fwd"][0], f.count(reads[0]))
        assert np.array_equal(res["counts_rev"][2], f.count(reads[2], revcomp=True))
