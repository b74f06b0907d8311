#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ from the reference's own test data.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):   python tests/golden/make_golden.py

Sources (all under /root/reference):
  src/test/libIBFTests/data/{test.fasta,test.ibf,test1.fasta,test1.ibf}
  src/test/classifyTests/data/{test.fasta,test.fastq,test.ibf}
  testData/testQueries.fasta

The three .ibf files are 9.9 MB of almost-all-zero words, so they are committed
in sparse form (indices + values of the non-zero 64-bit words, plus size, md5,
header and metadata tail) -- enough to reconstruct each file byte for byte.
The small FASTA/FASTQ inputs are copied verbatim (they are test DATA; no
reference source code is copied).  Known answers quoted from the reference's
gtest sources are written to known_answers.json with their file:line.
"""
import hashlib
import json
import os
import shutil

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

IBFS = {
    "lib_test": "src/test/libIBFTests/data/test.ibf",
    "lib_test1": "src/test/libIBFTests/data/test1.ibf",
    "classify_test": "src/test/classifyTests/data/test.ibf",
}
DATA = {
    "lib_test.fasta": "src/test/libIBFTests/data/test.fasta",
    "lib_test1.fasta": "src/test/libIBFTests/data/test1.fasta",
    "classify_test.fasta": "src/test/classifyTests/data/test.fasta",
    "classify_test.fastq": "src/test/classifyTests/data/test.fastq",
    "testQueries.fasta": "testData/testQueries.fasta",
}

KNOWN = {
    "read35": {"seq": "AAAAAAACCCCCCCCCGAGAGAGGAGAGAGGAGAG", "cite": "src/test/libIBFTests/read.hpp:113"},
    "read35_revcomp": {"seq": "CTCTCCTCTCTCCTCTCTCGGGGGGGGGTTTTTTT", "cite": "src/test/libIBFTests/read.hpp:273",
                       "max_kmer_count_int_compare": 23, "cite_count": "src/test/libIBFTests/read.hpp:327"},
    "ci_0.1_13_35_0.95": {"low": 5, "high": 30, "threshold_int16": -7,
                          "cite": "src/test/libIBFTests/read.hpp:156-164"},
    "count_matches_354": {"lib_test": 282, "lib_test1": 182, "best_index": 0, "pair": [282, 182],
                          "readlen": 354, "cite": "src/test/libIBFTests/read.hpp:199-251"},
    "filter_size_bits": {"fragment_length": 100000, "k": 13, "h": 3, "max_fp": 0.01, "bins": 2,
                         "value": 79121216, "bin_size_bits": 1236269,
                         "cite": "src/test/libIBFTests/createfilter.hpp:141-148"},
    "cut_out_nnns": {"in": "AAAAAAAACCCCCCCCCGAGAGAGGAGAGAGGAGAGAGAGAGCCCCAAAAGAGAGGAGATTTTANNNNNNNNTATATTATA",
                     "out": "AAAAAAAACCCCCCCCCGAGAGAGGAGAGAGGAGAGAGAGAGCCCCAAAAGAGAGGAGATTTTATATATTAT",
                     "cite": "src/test/libIBFTests/createfilter.hpp:107,135"},
    "fragment_arith": {"fragstart_after_first": 99988, "fragend": 72, "fragIdx": 1,
                       "cite": "src/test/libIBFTests/createfilter.hpp:168-171"},
    "filter_stats_test1": {"sumSeqLen": 2432, "totalBinsBinId": 4, "totalSeqsFile": 4,
                           "cite": "src/test/libIBFTests/createfilter.hpp:218-224 (test build parses twice)"},
    "ibf_config_defaults": {"MBinBits": 8388608, "overlap_length": 1500, "kmer_size": 13, "hash_functions": 3,
                            "threads": 2, "n_refs": 400, "n_batches": 500000, "max_fp": 0.01,
                            "cite": "src/test/libIBFTests/ibfconfigtest.hpp:32-59"},
    "classify_reads": {"found": 3, "failed": 0, "too_short": 0, "readCounter": 3,
                       "cite": "src/test/classifyTests/classifygtests.hpp:70-79"},
    "fragment_start_end": {"chunk_length": 360, "cite": "src/test/classifyTests/classifygtests.hpp:46-63"},
}


def main():
    os.makedirs(os.path.join(HERE, "data"), exist_ok=True)
    meta = {}
    arrays = {}
    for name, rel in IBFS.items():
        raw = open(os.path.join(REF, rel), "rb").read()
        bit_len = int(np.frombuffer(raw[:8], "<u8")[0])
        words = np.frombuffer(raw[8:], "<u8")
        nz = np.nonzero(words)[0].astype(np.uint64)
        arrays[name + "_idx"] = nz
        arrays[name + "_val"] = words[nz]
        nb = bit_len - 256
        meta[name] = {
            "source": rel, "file_bytes": len(raw), "md5": hashlib.md5(raw).hexdigest(),
            "bit_length": bit_len, "n_words": int(words.size),
            "tail": [int(x) for x in words[nb // 64: nb // 64 + 4]],
        }
    np.savez_compressed(os.path.join(HERE, "ibf_sparse.npz"), **arrays)
    for dst, rel in DATA.items():
        shutil.copyfile(os.path.join(REF, rel), os.path.join(HERE, "data", dst))
        os.chmod(os.path.join(HERE, "data", dst), 0o644)
    # the 354-base read of read.hpp:22 is a C++ string literal in the test source
    import re
    src = open(os.path.join(REF, "src/test/libIBFTests/read.hpp")).read()
    lits = re.findall(r'"([ACGTN]{100,})"', src)
    assert len(lits) >= 1 and len(lits[0]) == 354, [len(x) for x in lits]
    KNOWN["read354"] = {"seq": lits[0], "cite": "src/test/libIBFTests/read.hpp:22"}
    json.dump({"ibf": meta, "known": KNOWN}, open(os.path.join(HERE, "known_answers.json"), "w"), indent=1)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
