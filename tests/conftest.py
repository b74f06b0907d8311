import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def read_fasta(path):
    """Minimal FASTA/FASTQ reader for the fixtures (handles CRLF and wrapped FASTA)."""
    recs = []
    with open(path, "rb") as fh:
        lines = [ln.rstrip(b"\r\n") for ln in fh]
    i = 0
    while i < len(lines):
        ln = lines[i]
        if ln.startswith(b">"):
            name = ln[1:].decode()
            i += 1
            seq = []
            while i < len(lines) and not lines[i].startswith(b">"):
                seq.append(lines[i])
                i += 1
            recs.append((name, b"".join(seq)))
        elif ln.startswith(b"@"):
            recs.append((ln[1:].decode(), lines[i + 1]))
            i += 4
        else:
            i += 1
    return recs


@pytest.fixture(scope="session")
def known():
    return json.load(open(os.path.join(GOLDEN, "known_answers.json")))


@pytest.fixture(scope="session")
def golden_sparse():
    return np.load(os.path.join(GOLDEN, "ibf_sparse.npz"))


def golden_words(known, golden_sparse, name):
    """Reconstruct the full word array (incl. metadata tail) of a golden .ibf."""
    meta = known["ibf"][name]
    words = np.zeros(meta["n_words"], dtype=np.uint64)
    words[golden_sparse[name + "_idx"].astype(np.int64)] = golden_sparse[name + "_val"]
    return meta, words


def golden_file_bytes(known, golden_sparse, name):
    meta, words = golden_words(known, golden_sparse, name)
    return np.array([meta["bit_length"]], dtype="<u8").tobytes() + words.astype("<u8").tobytes()


@pytest.fixture(scope="session")
def golden_ibf_paths(known, golden_sparse, tmp_path_factory):
    """Materialise the three golden .ibf files (byte-identical, md5-checked) in a tmp dir."""
    import hashlib
    d = tmp_path_factory.mktemp("golden_ibf")
    out = {}
    for name in known["ibf"]:
        raw = golden_file_bytes(known, golden_sparse, name)
        assert hashlib.md5(raw).hexdigest() == known["ibf"][name]["md5"]
        assert len(raw) == known["ibf"][name]["file_bytes"]
        p = d / (name + ".ibf")
        p.write_bytes(raw)
        out[name] = str(p)
    return out


def data_path(name):
    return os.path.join(GOLDEN, "data", name)
