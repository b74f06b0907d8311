"""CPU-only tests of the product's host side: the C-ABI library loads and exports every symbol
include/rb_ibf.h declares, the FP64/host helpers agree with the oracle, compute entry points fail
loudly without a GPU, and the C++ interleave:: shim compiles and passes its host checks."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import oracle
import readbouncer_b200 as rb
from readbouncer_b200 import capi, synth
from conftest import ROOT, data_path, read_fasta

INCLUDE = os.path.join(ROOT, "include")
LIBDIR = os.path.join(ROOT, "readbouncer_b200", "lib")


def declared_symbols():
    hdr = open(os.path.join(INCLUDE, "rb_ibf.h")).read()
    return sorted(set(re.findall(r"RB_API[^;(]*?\b(rb_\w+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    rb.build_library()
    names = declared_symbols()
    assert len(names) >= 25
    L = ctypes.CDLL(rb.lib_path())
    for n in names:
        assert hasattr(L, n), n
    assert sorted(names) == sorted(capi.EXPORTS)      # the binding covers the whole header
    nm = subprocess.run(["nm", "-D", "--defined-only", rb.lib_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (rb_\w+)", nm))
    assert set(names) <= exported


def test_status_strings():
    L = rb.lib()
    assert L.rb_status_string(0) == b"ok"
    assert L.rb_status_string(1) == b"NullFilterException" and L.rb_status_string(2) == b"ShortReadException"
    assert L.rb_status_string(7) == b"InsertSequenceException" and L.rb_status_string(11) == b"no CUDA device"


def test_threshold_lut_equals_oracle():
    for e, k in [(0.1, 13), (0.08, 13), (0.1, 15), (0.08, 15), (0.05, 11), (0.2, 21), (0.1, 27)]:
        assert np.array_equal(rb.threshold_lut(e, k), oracle.threshold_lut(e, k)), (e, k)
    lut = rb.threshold_lut(0.1, 13)
    assert lut[35] == 65529 and lut[250] == 18 and lut[354] == 36          # read.hpp:164; SURVEY A.8
    assert rb.threshold_lut(0.08, 13)[250] == 34 and rb.threshold_lut(0.1, 15)[360] == 22
    assert rb.calculate_ci(0.1, 13, 35) == (5, 30) == oracle.calculate_ci(0.1, 13, 35, 0.95)
    with pytest.raises(rb.RBError) as e:
        rb.threshold_lut(0.0, 13)
    assert e.value.status == 8


def test_size_bits_cut_and_schedule_equal_oracle():
    rng = np.random.default_rng(0)
    for F, k, bins in [(100000, 13, 2), (100000, 13, 64), (100000, 15, 2), (4200000, 13, 100), (100000, 13, 31024),
                       (2000, 11, 1100)]:
        assert rb.ibf_size_bits(F, k, 3, 0.01, bins) == oracle.filter_size_bits(F, k, 3, 0.01, bins)
    assert rb.ibf_size_bits(100000, 13, 3, 0.01, 2) == 79121216            # createfilter.hpp:148
    for _ in range(300):
        n = int(rng.integers(0, 80))
        s = bytes(rng.choice(np.frombuffer(b"ACGTNNNn", np.uint8), size=n).tolist())
        assert rb.cut_out_nnns(s) == oracle.cut_out_nnns(s), s
    for _ in range(300):
        F = int(rng.integers(20, 500))
        k = int(rng.integers(5, 20))
        n = int(rng.integers(0, 3000))
        b0, e0 = rb.fragment_schedule(n, F, k)
        b1, e1 = oracle.fragment_schedule(n, F, k)
        assert np.array_equal(b0, b1) and np.array_equal(e0, e1), (n, F, k)


def test_build_plan_matches_oracle_builder():
    seqs = [s for _, s in read_fasta(data_path("lib_test1.fasta"))] + [b"ACGT", b"ACGTNNNNACGTACGTACGTAAA"]
    plan = synth.build_plan(seqs, 300, 13)
    of, stats = oracle.build_from_sequences(seqs, 300, k=13)
    assert plan["n_bins"] == stats["totalBinsBinId"] and plan["n_bits"] == stats["filter_size_bits"]
    assert plan["invalid_seqs"] == stats["invalidSeqs"] == 1 and plan["sum_seq_len"] == stats["sumSeqLen"]
    assert plan["bin_ids_consumed"] == stats["bin_ids_consumed"]


def test_no_gpu_means_loud_failure_not_fallback(golden_ibf_paths):
    if rb.device_count() > 0:
        pytest.skip("a CUDA device is present")
    for fn in (lambda: rb.IBF.create(2, 3, 13, 79121216), lambda: rb.IBF.load(golden_ibf_paths["lib_test"]),
               lambda: rb.IBF.from_words(np.zeros(79121216 // 64, np.uint64), 2, 3, 13, 79121216)):
        with pytest.raises(rb.RBError) as e:
            fn()
        assert e.value.status == 11
    # file problems are still diagnosed before the device is needed
    with pytest.raises(rb.RBError) as e:
        rb.IBF.load(data_path("lib_test.fasta"))
    assert e.value.status == 4
    L = rb.lib()
    assert L.rb_ibf_count_batch(None, None, None, 1, None, 1, None, None, None, None, None, None, None) == 1
    assert L.rb_ibf_insert_batch(None, None, 0, None, None, None, 1, None) == 1


def test_product_never_touches_the_oracle():
    """The product path must not import, link or call anything under oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "readbouncer_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "ibf_oracle" not in txt and "oracle/" not in txt, f
    for f in os.listdir(INCLUDE):
        assert "oracle" not in open(os.path.join(INCLUDE, f)).read(), f
    ldd = subprocess.run(["ldd", rb.lib_path()], capture_output=True, text=True).stdout
    assert "oracle" not in ldd


def _compile(src, out, link=True):
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I" + INCLUDE, "-I/usr/local/cuda/include", src, "-o", out]
    if link:
        cmd += ["-L" + LIBDIR, "-lrb_ibf", "-Wl,-rpath," + LIBDIR]
    subprocess.check_call(cmd)
    return out


def test_plain_c_client_of_the_header(tmp_path):
    """include/rb_ibf.h is valid, warning-free C99 and a pure C program (what cgo / JNI / N-API would bind) gets the
    reference's known answers from the host-side entry points; on this GPU-less host the compute calls fail loudly."""
    exe = str(tmp_path / "c_client")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + INCLUDE,
                           os.path.join(ROOT, "tests", "c", "test_c_client.c"), "-L" + LIBDIR, "-lrb_ibf",
                           "-Wl,-rpath," + LIBDIR, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "c client ok" in out.stdout, out.stdout + out.stderr


def test_fast_mod_exhaustive_edges(tmp_path):
    exe = _compile(os.path.join(ROOT, "tests", "cpp", "test_fastmod.cpp"), str(tmp_path / "test_fastmod"), link=False)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "fast_mod OK" in out.stdout, out.stdout + out.stderr


def test_bit_transpose_matches_definition(tmp_path):
    """rb::transpose32 (the column build's 32 bins x 32 rows tile) against the bit-by-bit definition."""
    exe = _compile(os.path.join(ROOT, "tests", "cpp", "test_transpose.cpp"), str(tmp_path / "test_transpose"), link=False)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stdout + out.stderr


def test_postings_list_order_is_a_bijection_and_spreads_banks(tmp_path):
    """rb::list_position (order of the bin ids inside a postings list) for every list length up to 3 000."""
    exe = _compile(os.path.join(ROOT, "tests", "cpp", "test_postings_layout.cpp"), str(tmp_path / "test_postings_layout"), link=False)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout + out.stderr


def test_slot_order_is_a_bijection_and_spreads_banks(tmp_path):
    """rb::slot_position / slot_pad_id (order of the bin ids inside a postings slot) for every length a slot can hold."""
    exe = _compile(os.path.join(ROOT, "tests", "cpp", "test_slot_layout.cpp"), str(tmp_path / "test_slot_layout"), link=False)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout + out.stderr


@pytest.mark.parametrize("isa", ["5", "2", "0"])
def test_host_packer_matches_restatement(tmp_path, isa):
    """Bit planes of the host packer (AVX-512 / AVX2 / scalar paths, thread pool) against a plain loop."""
    csrc = os.path.join(ROOT, "readbouncer_b200", "csrc")
    exe = str(tmp_path / "test_host_pack")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-I" + csrc, os.path.join(ROOT, "tests", "cpp", "test_host_pack.cpp"),
                           os.path.join(csrc, "host_pack.cpp"), "-lpthread", "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, env=dict(os.environ, RB_HOST_PACK_ISA=isa, RB_HOST_THREADS="4"))
    assert out.returncode == 0 and "host_pack OK" in out.stdout, out.stdout + out.stderr


def test_cpp_shim_host_checks(tmp_path, golden_ibf_paths):
    rb.build_library()
    exe = _compile(os.path.join(ROOT, "tests", "cpp", "test_shim.cpp"), str(tmp_path / "test_shim"))
    out = subprocess.run([exe, "host"], capture_output=True, text=True)
    assert out.returncode == 0 and "host OK" in out.stdout, out.stdout + out.stderr
    if rb.device_count() == 0:
        out = subprocess.run([exe, "nogpu", golden_ibf_paths["lib_test"]], capture_output=True, text=True)
        assert out.returncode == 0 and "nogpu OK" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_cpp_shim_on_gpu(tmp_path, golden_ibf_paths, known):
    exe = _compile(os.path.join(ROOT, "tests", "cpp", "test_shim.cpp"), str(tmp_path / "test_shim"))
    out = subprocess.run([exe, "gpu", golden_ibf_paths["lib_test"], golden_ibf_paths["lib_test1"],
                          data_path("lib_test.fasta"), str(tmp_path), known["known"]["read354"]["seq"]],
                         capture_output=True, text=True)
    assert out.returncode == 0 and "gpu OK" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_cpp_host_bin_sharded_combine(tmp_path):
    """A C++ host with nothing but the C ABI: whole filter == rb_ibf_count_batch_sharded (keys folded over NVLink peer
    memory; 2 shards per visible device) == per-rank count + rb_keys_combine_nccl (when there are >= 2 GPUs)."""
    exe = str(tmp_path / "test_shard_combine")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I" + INCLUDE, "-I/usr/local/cuda/include",
                           os.path.join(ROOT, "tests", "cpp", "test_shard_combine.cpp"), "-o", exe, "-L" + LIBDIR, "-lrb_ibf",
                           "-Wl,-rpath," + LIBDIR, "-L/usr/local/cuda/lib64", "-lcudart", "-lnccl"])
    out = subprocess.run([exe, "3000"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "shard combine OK" in out.stdout, out.stdout + out.stderr


def test_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the CPU oracle port on all host cores) on a small workload: one JSON line with the
    contract's keys, no GPU work, and -- under a multi-rank launch -- nothing from ranks other than 0."""
    import json
    import sys
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "mini_100x60kb_100bins",
           "--steps", "1", "--warmup", "0"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "classified_250b_read_chunks_per_sec" and d["unit"] == "chunks/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    other = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert other.returncode == 0 and other.stdout.strip() == ""


def test_bench_refuses_to_run_without_a_gpu():
    """No CPU fallback anywhere: on a box without a CUDA device the product arm of bench.py stops with an error."""
    import sys
    if rb.device_count() > 0:
        pytest.skip("a CUDA device is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "mini_100x60kb_100bins", "--steps", "1"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode != 0 and "no CPU fallback" in out.stderr
    assert not any(ln.startswith("{") for ln in out.stdout.splitlines())
