#!/usr/bin/env python3
"""The reference's one published workload (README.md:233-262), reshaped with synthetic data: 3 target genomes + 1
depletion genome (sizes of B. subtilis / E. faecalis / E. coli / S. cerevisiae: 4.2 / 2.9 / 4.6 / 12.1 Mb in 1 + 1 + 1 + 17
records), fragment_size 100 000, k = 13, and N reads of nanopore-like length through usage=classify -- FASTA in, one FASTA
per target + unclassified.fasta out -- with the deplete + target retry (classify.hpp:58-111).  The reference reports
"Average Processing Time Read Classification : 0.00197617" s per read for 100 000 reads (hardware unstated).

Runs the C++ driver (readbouncer_b200/bin/rb_readbouncer, include/rb_drivers.hpp) on cuda:0, checks the assignment of a
sample of the reads against the oracle running the reference's serial loop (which is also the CPU time beside it), and
prints one JSON line.  `run()` is what bench.py's secondary entry and tests/test_drivers.py call.
"""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

README_S_PER_READ = 0.00197617
GENOMES = [("Bacillus_subtilis_like", [4_215_606]), ("Enterococcus_faecalis_like", [2_939_973]), ("Escherichia_coli_like", [4_641_652]),
           ("Saccharomyces_cerevisiae_like", [230_218, 813_184, 316_620, 1_531_933, 576_874, 270_161, 1_090_940, 562_643, 439_888,
                                              745_751, 666_816, 1_078_177, 924_431, 784_333, 1_091_291, 948_066, 85_779])]
LENGTH_CLASSES = [(400, 0.10), (900, 0.20), (1800, 0.25), (3500, 0.25), (7000, 0.15), (14000, 0.05)]


def write_fasta(path, records, width=70):
    with open(path, "wb") as fh:
        for name, seq in records:
            fh.write(b">" + name.encode() + b"\n")
            b = seq.tobytes() if isinstance(seq, np.ndarray) else bytes(seq)
            for j in range(0, len(b), width):
                fh.write(b[j:j + width] + b"\n")


def make_workload(outdir, n_reads, seed=77):
    """Genome FASTA files, the read FASTA, and what is needed to re-derive expectations."""
    from readbouncer_b200 import synth
    genomes, paths = [], []
    for gi, (name, lens) in enumerate(GENOMES):
        recs = [("%s_%d" % (name, i), synth.hash_bases(0, n, 9000 + 100 * gi + i)) for i, n in enumerate(lens)]
        p = os.path.join(outdir, name + ".fasta")
        write_fasta(p, recs)
        genomes.append(recs)
        paths.append(p)
    rng = np.random.default_rng(seed)
    # read origin: the three targets 16 / 14 / 20 %, the depletion genome 20 %, "the rest of the community" (iid) 30 %
    origin = rng.choice(5, size=n_reads, p=[0.16, 0.14, 0.20, 0.20, 0.30])
    cls = rng.choice(len(LENGTH_CLASSES), size=n_reads, p=[p for _, p in LENGTH_CLASSES])
    reads = [None] * n_reads
    for o in range(5):
        for c, (L, _) in enumerate(LENGTH_CLASSES):
            idx = np.nonzero((origin == o) & (cls == c))[0]
            if not idx.size:
                continue
            if o < 4:
                cat = np.concatenate([s for _, s in genomes[o]])
                b, _, _ = synth.sample_reads(cat, idx.size, L, seed=int(rng.integers(1 << 30)), frac_from_ref=1.0, error_rate=0.08)
            else:
                b = synth.random_bases(idx.size * L, int(rng.integers(1 << 30)))
            b = b.reshape(idx.size, L)
            for j, i in enumerate(idx):
                reads[i] = b[j].tobytes()
    rp = os.path.join(outdir, "reads.fasta")
    with open(rp, "wb") as fh:
        for i, r in enumerate(reads):
            fh.write(b">read%06d\n" % i + r + b"\n")
    return paths, genomes, rp, reads, origin


def read_ids(path):
    return [ln[1:].strip().decode() for ln in open(path, "rb") if ln.startswith(b">")]


def run(n_reads=100_000, sample=1500, workdir=None, keep=False, cpu_threads=None):
    import oracle
    import readbouncer_b200 as rb
    rb.build_library()
    exe = os.path.join(ROOT, "readbouncer_b200", "bin", "rb_readbouncer")
    tmp = workdir or tempfile.mkdtemp(prefix="rb_readme_")
    os.makedirs(tmp, exist_ok=True)
    t0 = time.time()
    paths, genomes, rp, reads, origin = make_workload(tmp, n_reads)
    gen_s = time.time() - t0
    out = os.path.join(tmp, "out")
    cfg = os.path.join(tmp, "config.toml")
    q = lambda xs: "[" + ", ".join("'%s'" % x for x in xs) + "]"
    open(cfg, "w").write("usage = \"classify\"\noutput_directory = '%s'\nlog_directory = '%s/logs'\n\n[IBF]\nkmer_size = 13\n"
                         "fragment_size = 100000\nthreads = 3\ntarget_files = %s\ndeplete_files = %s\nread_files = %s\n"
                         % (out, out, q(paths[:3]), q(paths[3:]), q([rp])))
    t0 = time.time()
    p = subprocess.run([exe, "--config", cfg], capture_output=True, text=True)
    wall_s = time.time() - t0
    if p.returncode != 0:
        raise RuntimeError("driver failed: " + p.stderr[-2000:])
    m = re.search(r"RESULT found=(\d+) failed=(\d+) too_short=(\d+) reads=(\d+) avg_classify_s=(\S+) read_file_s=(\S+) table_setup_s=(\S+)", p.stdout)
    found, failed, too_short, n = map(int, m.groups()[:4])
    avg_s, read_file_s, table_s = map(float, m.groups()[4:])
    assert n == n_reads
    names = [name for name, _ in GENOMES[:3]]
    assign = np.full(n_reads, -9, np.int64)
    for ti, name in enumerate(names):
        for rid in read_ids(os.path.join(out, name + ".fasta")):
            assign[int(rid[4:])] = ti
    for rid in read_ids(os.path.join(out, "unclassified.fasta")):
        assign[int(rid[4:])] = -1
    assert (assign != -9).sum() + too_short + failed == n_reads
    per_target = [int((assign == ti).sum()) for ti in range(3)]
    # the oracle runs the reference's serial loop on a sample of the reads: expected assignments + CPU seconds per read
    ofs = []
    for recs in genomes:
        raw = [s.tobytes() for _, s in recs]
        ofs.append(oracle.build_from_sequences(raw, 100000, k=13, n_threads=cpu_threads or (os.cpu_count() or 1))[0])
    # the FASTA inputs were built into <output_dir>/<stem>.ibf by the driver (ibfbuild.hpp:111-115): byte-identical to the oracle's
    for gi, (name, _) in enumerate(GENOMES):
        ref = os.path.join(tmp, "oracle_%d.ibf" % gi)
        ofs[gi].store(ref)
        assert open(os.path.join(out, name + ".ibf"), "rb").read() == open(ref, "rb").read(), "GPU-built %s.ibf differs from the oracle's" % name
        os.unlink(ref)
    pick = np.sort(np.random.default_rng(3).choice(n_reads, min(sample, n_reads), replace=False))
    t0 = time.time()
    exp = oracle.classify_reads_serial([reads[i] for i in pick], [ofs[3]], ofs[:3], 250, 5, 0.1)
    cpu_s = time.time() - t0
    exp = np.asarray(exp)
    got = assign[pick]
    got_cmp = np.where(got == -9, np.where(exp == -4, -4, -3), got)          # too short / failed reads are in no output file
    assert np.array_equal(got_cmp, exp), "driver assignments differ from the serial oracle on %d sampled reads" % int((got_cmp != exp).sum())
    res = {"workload": "readme_3targets_1deplete", "reads": n_reads, "genomes_mb": [sum(l) / 1e6 for _, l in GENOMES],
           "bins": [int(f.n_bins) for f in ofs], "kmer_size": 13, "fragment_size": 100000, "chunk_length": 250, "max_chunks": 5,
           "found": found, "failed": failed, "too_short": too_short, "per_target": dict(zip(names, per_target)),
           "s_per_read": avg_s, "reads_per_s": 1.0 / avg_s if avg_s else None,
           "s_per_read_what": "chunk loop + decisions + output FASTA records (what the reference's per-read timer covers), batched per chunk index",
           "readme_s_per_read": README_S_PER_READ, "speedup_vs_readme": README_S_PER_READ / avg_s if avg_s else None,
           "cpu_serial_loop_s_per_read": cpu_s / len(pick), "cpu_serial_loop": "oracle port of classify.hpp:229-303, one thread, %d sampled reads" % len(pick),
           "speedup_vs_cpu_serial_loop": (cpu_s / len(pick)) / avg_s if avg_s else None,
           "read_file_s": read_file_s, "table_setup_s": table_s, "driver_wall_s": wall_s, "generate_inputs_s": gen_s,
           "parity": "assignment (target index / unclassified / too short) of %d sampled reads == serial oracle loop; the 4 GPU-built .ibf files "
                     "are byte-identical to the oracle's" % len(pick)}
    tabs = re.findall(r"k-mer table: (.*)", p.stderr)
    if tabs:
        res["kmer_tables"] = tabs
    if not keep and workdir is None:
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)
    return res


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    print(json.dumps(run(n)))
