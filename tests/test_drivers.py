"""usage=build / usage=classify drivers (include/rb_drivers.hpp, tools/rb_readbouncer.cpp): TOML contract on
CPU; on the GPU the reference's classifyTests result (3/3 found) and a multi-filter scenario whose expected
outcome is computed read by read with the oracle, following src/main/classify.hpp."""
import hashlib
import os
import re
import subprocess

import numpy as np
import pytest

import oracle
import readbouncer_b200 as rb
from readbouncer_b200 import synth
from conftest import ROOT, data_path, read_fasta

EXE = os.path.join(ROOT, "readbouncer_b200", "bin", "rb_readbouncer")


def write_toml(path, usage, outdir, targets=(), depletes=(), reads=(), k=13, frag=100000, chunk=250, max_chunks=5, err=0.1):
    q = lambda xs: "[" + ", ".join("'%s'" % x for x in xs) + "]"
    txt = ["usage = \"%s\"   # comment" % usage, "output_directory = '%s'" % outdir, "log_directory = '%s/logs'" % outdir, "",
           "[IBF]", "kmer_size = %d" % k, "fragment_size = %d  # default 100000" % frag, "threads = 3",
           "exp_seq_error_rate = %s" % err, "chunk_length = %d" % chunk, "max_chunks = %d" % max_chunks]
    if targets:
        txt.append("target_files = " + q(targets))
    if depletes:
        txt.append("deplete_files = " + q(depletes))
    if reads:
        txt.append("read_files = " + q(reads))
    txt += ["", "[MinKNOW]", "host = \"localhost\"", "channels = [1,512]", "", "[Basecaller]", "caller = \"DeepNano\""]
    open(path, "w").write("\n".join(txt) + "\n")
    return str(path)


def run(cfg, *extra):
    rb.build_library()
    return subprocess.run([EXE, "--config", cfg, *extra], capture_output=True, text=True)


def test_toml_contract_and_defaults(tmp_path):
    cfg = write_toml(tmp_path / "c.toml", "classify", tmp_path / "out", targets=[data_path("lib_test.fasta")],
                     reads=[data_path("classify_test.fastq")], k=15, chunk=360, max_chunks=4)
    out = run(cfg, "--print-config")
    assert out.returncode == 0, out.stderr
    assert "kmer_size=15" in out.stdout and "chunk_length=360" in out.stdout and "max_chunks=4" in out.stdout
    assert "fragment_size=100000" in out.stdout and "targets=1 depletes=0 reads=1" in out.stdout
    # defaults of ConfigReader::readIBF (configReader.cpp:238-243)
    p = tmp_path / "d.toml"
    p.write_text("usage = 'build'\n[IBF]\ntarget_files = ['%s']\n" % data_path("lib_test.fasta"))
    out = run(str(p), "--print-config")
    assert "kmer_size=13 fragment_size=100000 threads=1 exp_seq_error_rate=0.1 chunk_length=250 max_chunks=5" in out.stdout
    p.write_text("usage = 'classify'\n[IBF]\nkmer_size = 13\n")
    out = run(str(p), "--print-config")
    assert out.returncode == 1 and "At least one target or deplete file" in out.stderr
    p.write_text("usage = 'classify'\n[IBF]\ntarget_files = ['/nonexistent.ibf']\n")
    out = run(str(p), "--print-config")
    assert out.returncode == 1 and "does not exist" in out.stderr


@pytest.mark.gpu
def test_build_usage_writes_reference_identical_ibf(tmp_path, known, golden_sparse):
    """usage=build on the production path (each sequence once): file equals the oracle's build byte for byte."""
    outdir = tmp_path / "out"
    cfg = write_toml(tmp_path / "b.toml", "build", outdir, targets=[data_path("classify_test.fasta")], k=15)
    out = run(cfg)
    assert out.returncode == 0, out.stderr + out.stdout
    built = (outdir / "classify_test.ibf").read_bytes()
    seqs = [s for _, s in read_fasta(data_path("classify_test.fasta"))]
    of, stats = oracle.build_from_sequences(seqs, 100000, k=15, passes=1)
    ref = tmp_path / "oracle.ibf"
    of.store(ref)
    assert hashlib.md5(built).hexdigest() == hashlib.md5(ref.read_bytes()).hexdigest()
    assert "1 sequences in 1 bins were written to the IBF" in out.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("chunk,max_chunks,found", [(360, 5, 3), (360, 4, 3), (360, 1, 2), (250, 5, 3)])
def test_classify_usage_on_reference_fixture(tmp_path, golden_ibf_paths, chunk, max_chunks, found):
    """classifyTests/ClassifyReadsTest: found 3, failed 0, too_short 0, reads 3 (classifygtests.hpp:70-79)."""
    outdir = tmp_path / "out"
    cfg = write_toml(tmp_path / "c.toml", "classify", outdir, targets=[golden_ibf_paths["classify_test"]],
                     reads=[data_path("classify_test.fastq")], k=15, chunk=chunk, max_chunks=max_chunks)
    out = run(cfg)
    assert out.returncode == 0, out.stderr + out.stdout
    m = re.search(r"RESULT found=(\d+) failed=(\d+) too_short=(\d+) reads=(\d+)", out.stdout)
    assert tuple(map(int, m.groups())) == (found, 0, 0, 3)
    assert "Number of classified reads                         :   %d" % found in out.stdout
    recs = read_fasta(str(outdir / "classify_test.fasta"))
    assert len(recs) == found
    un = read_fasta(str(outdir / "unclassified.fasta"))
    assert len(un) == 3 - found


_expected_classify_reads = oracle.classify_reads_serial      # src/main/classify.hpp:229-303 read by read with the oracle


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["both", "deplete", "target"])
def test_classify_usage_multi_filter_equals_serial_reference_logic(tmp_path, mode):
    rng = np.random.default_rng(5)
    genomes = [synth.random_bases(120000, 700 + i) for i in range(3)]          # 2 targets + 1 deplete, FASTA inputs
    paths = []
    for i, g in enumerate(genomes):
        p = tmp_path / ("g%d.fasta" % i)
        p.write_bytes(b">g%d some description\n" % i + b"\n".join(g.tobytes()[j:j + 80] for j in range(0, len(g), 80)) + b"\n")
        paths.append(str(p))
    # reads: from each genome, chimeric (target prefix + deplete suffix), random, short, with N; ragged lengths
    reads = []
    for i in range(240):
        kind = i % 6
        L = int(rng.integers(200, 1400))
        if kind < 3:
            g = genomes[kind]
            s = int(rng.integers(0, len(g) - L))
            r = g[s:s + L].copy()
        elif kind == 3:
            a, b = genomes[0], genomes[2]
            sa, sb = int(rng.integers(0, len(a) - 700)), int(rng.integers(0, len(b) - 700))
            r = np.concatenate([a[sa:sa + 130], b[sb:sb + 130], a[sa + 130:sa + 130 + max(0, L - 260)]])
        elif kind == 4:
            r = synth.random_bases(L, 10_000 + i)
        else:
            r = synth.random_bases(int(rng.integers(20, 260)), 20_000 + i)
        mut = rng.random(len(r)) < 0.06
        r[mut] = synth.ACGT[rng.integers(0, 4, size=int(mut.sum()))]
        if i % 17 == 0:
            r[rng.integers(0, len(r), size=3)] = ord("N")
        reads.append(r.tobytes())
    rf = tmp_path / "reads.fasta"
    rf.write_bytes(b"".join(b">read%d\n%s\n" % (i, r) for i, r in enumerate(reads)))
    tg, dp = ([paths[0], paths[1]] if mode != "deplete" else []), ([paths[2]] if mode != "target" else [])
    outdir = tmp_path / "out"
    cfg = write_toml(tmp_path / "c.toml", "classify", outdir, targets=tg, depletes=dp, reads=[str(rf)], k=13, frag=100000,
                     chunk=250, max_chunks=5)
    out = run(cfg)
    assert out.returncode == 0, out.stderr + out.stdout
    # expected: the same filters built by the oracle, the serial loop of classify.hpp
    ofs = [oracle.build_from_sequences([g], 100000, k=13)[0] for g in genomes]
    exp = _expected_classify_reads(reads, [ofs[2]] if dp else [], [ofs[0], ofs[1]] if tg else [], 250, 5, 0.1)
    m = re.search(r"RESULT found=(\d+) failed=(\d+) too_short=(\d+) reads=(\d+)", out.stdout)
    found, failed, too_short, n = map(int, m.groups())
    assert n == len(reads) and too_short == sum(a == -4 for a in exp) and failed == sum(a == -3 for a in exp)
    assert found == sum(a >= 0 or a == -2 for a in exp)
    assert found > 30 and too_short > 10
    for ti, name in enumerate(["g0", "g1"] if tg else []):
        ids = [nm for nm, _ in read_fasta(str(outdir / (name + ".fasta")))]
        assert ids == ["read%d" % i for i, a in enumerate(exp) if a == ti]
    un = [nm for nm, _ in read_fasta(str(outdir / "unclassified.fasta"))]
    assert un == ["read%d" % i for i, a in enumerate(exp) if a == -1]
    # the FASTA inputs were built into <output_dir>/<stem>.ibf on the way (ibfbuild.hpp:111-115), identical to the oracle's
    for gi in ([0, 1] if tg else []) + ([2] if dp else []):
        ref = tmp_path / ("o%d.ibf" % gi)
        ofs[gi].store(ref)
        assert (outdir / ("g%d.ibf" % gi)).read_bytes() == ref.read_bytes()
