"""Live micro-batching adaptor (include/rb_live.hpp) against a serial simulation of classify_live_reads
(src/main/adaptive_sampling.hpp:214-356) driven by the oracle's check_unblock."""
import os
import subprocess

import numpy as np
import pytest

import oracle
import readbouncer_b200 as rb
from readbouncer_b200 import synth
from conftest import ROOT

INCLUDE = os.path.join(ROOT, "include")
LIBDIR = os.path.join(ROOT, "readbouncer_b200", "lib")


def compile_live(tmp_path):
    rb.build_library()
    exe = str(tmp_path / "test_live")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I" + INCLUDE, os.path.join(ROOT, "tests", "cpp", "test_live.cpp"),
                           "-o", exe, "-L" + LIBDIR, "-lrb_ibf", "-Wl,-rpath," + LIBDIR, "-lpthread"])
    return exe


def test_live_adaptor_compiles(tmp_path):
    compile_live(tmp_path)


def serial_live(stream, dep, tgt, err):
    """classify_live_reads, one chunk at a time in arrival order."""
    once_seen, out = {}, []
    for rid, seq in stream:
        try:
            d = oracle.check_unblock(dep, tgt, seq, err)
        except oracle.OracleError:
            continue                                  # logged and dropped (adaptive_sampling.hpp:340-349)
        if d == 1:
            seen = once_seen.pop(rid, b"") + seq
            out.append((rid, 1, 0, len(seen)))
        elif d == 2:
            once_seen.pop(rid, None)
            out.append((rid, 2, 0, len(seq)))
        elif rid in once_seen:
            cat = once_seen[rid] + seq
            try:
                d2 = oracle.check_unblock(dep, tgt, cat, err)
            except oracle.OracleError:
                continue
            if d2 in (1, 2):
                del once_seen[rid]
                out.append((rid, d2, 0, len(cat)))
            elif len(cat) > 1500:
                del once_seen[rid]
                out.append((rid, 2, 1, len(cat)))
            else:
                once_seen[rid] = cat
        else:
            once_seen[rid] = seq
    return out, len(once_seen)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["both", "deplete", "target"])
def test_live_microbatches_equal_serial_loop(tmp_path, mode):
    exe = compile_live(tmp_path)
    rng = np.random.default_rng(11)
    genomes = [synth.random_bases(150000, 900 + i) for i in range(2)]       # 0 = deplete, 1 = target
    ofs, paths = [], []
    for i, g in enumerate(genomes):
        of, _ = oracle.build_from_sequences([g], 100000, k=13)
        p = tmp_path / ("f%d.ibf" % i)
        of.store(p)
        ofs.append(of)
        paths.append(str(p))
    # 300 reads of 8 chunks (250 bases every "0.4 s"): deplete-like, target-like, chimeric, random; ~8 % errors
    reads = []
    for r in range(300):
        kind = r % 4
        L = 2000
        if kind < 2:
            g = genomes[kind]
            s = int(rng.integers(0, len(g) - L))
            seq = g[s:s + L].copy()
        elif kind == 2:
            seq = synth.random_bases(L, 5000 + r)
        else:
            a, b = genomes[0], genomes[1]
            sa, sb = int(rng.integers(0, len(a) - L)), int(rng.integers(0, len(b) - L))
            seq = np.concatenate([a[sa:sa + 125], b[sb:sb + 125]] * 8)
        mut = rng.random(len(seq)) < 0.08
        seq[mut] = synth.ACGT[rng.integers(0, 4, size=int(mut.sum()))]
        reads.append(seq.tobytes())
    # arrival order: chunk c of every read in a shuffled order, in micro-batches of 64; a decided read sends no more chunks
    stream, decided = [], set()
    dep = [ofs[0]] if mode != "target" else []
    tgt = [ofs[1]] if mode != "deplete" else []
    lines, batch_no = [], 0
    sim_once, expected = {}, []
    for c in range(8):
        order = rng.permutation(len(reads))
        live = [r for r in order if r not in decided]
        for b0 in range(0, len(live), 64):
            chunk_batch = [(("read%d" % r), reads[r][c * 250:(c + 1) * 250]) for r in live[b0:b0 + 64]]
            # a few very short chunks exercise the exception path of the single-list modes
            if c == 1 and b0 == 0:
                chunk_batch[0] = (chunk_batch[0][0], chunk_batch[0][1][:9])
            for rid, seq in chunk_batch:
                lines.append("%d\t%s\t%s" % (batch_no, rid, seq.decode()))
                stream.append((rid, seq))
            batch_no += 1
        exp_so_far, _ = serial_live(stream, dep, tgt, 0.1)
        decided = {int(rid[4:]) for rid, *_ in exp_so_far}
    expected, pending = serial_live(stream, dep, tgt, 0.1)
    sf = tmp_path / "stream.tsv"
    sf.write_text("\n".join(lines) + "\n")
    cmd = [exe, str(sf), "0.1", str(len(dep))] + ([paths[0]] if dep else []) + [str(len(tgt))] + ([paths[1]] if tgt else [])
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    rows = [ln.split("\t") for ln in out.stdout.strip().splitlines()]
    got = [(r[0], int(r[1]), int(r[2]), int(r[3])) for r in rows if r[0] != "PENDING"]
    got_pending = int([r for r in rows if r[0] == "PENDING"][0][1])
    # decisions of one micro-batch are emitted in arrival order, so the whole sequence matches the serial loop
    assert got == expected
    assert got_pending == pending
    kinds = {a for _, a, _, _ in got}
    if mode != "target":                               # target-only never keeps going: no hit means unblock
        assert any(g for _, _, g, _ in got)           # some reads were given up on after > 1500 bases
    assert (1 in kinds or mode == "target") and (2 in kinds or mode == "deplete")
