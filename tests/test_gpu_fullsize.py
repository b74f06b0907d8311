"""GPU parity at BASELINE.json's full sizes (configs #3, #4, #5) and under concurrent callers.

The multi-Gb references are generated in HBM (rb_synth_bases_dev: a pure function of seed and position); the host
regenerates only what it needs (windows the reads are sampled from, the fragments the oracle builds).  Everything the
GPU returns is compared with the CPU oracle on >= 3 000 sampled chunks: max_count, hit AND argmax_bin, both thresholds.
"""
import threading

import numpy as np
import pytest

import oracle
import readbouncer_b200 as rb
from readbouncer_b200 import synth

pytestmark = pytest.mark.gpu

SAMPLE = 3000


@pytest.fixture(autouse=True)
def _reset_kernel_choice():
    yield
    rb.set_count_kernel(0)
    rb.set_insert_kernel(0)


def _torch():
    import torch
    torch.cuda.set_device(0)
    return torch


class _DevWords:
    """The bit matrix of a handle as a torch tensor (no copy), through __cuda_array_interface__."""

    def __init__(self, gf):
        self.__cuda_array_interface__ = {"shape": (gf.n_local_words,), "typestr": "<i8", "data": (gf.device_words_ptr(), False),
                                         "version": 2}


def _build_on_gpu(gf, ref, plan, torch, variant=0):
    dev = torch.device("cuda", gf.device)
    rb.set_insert_kernel(variant)
    d_ref = ref.to_device(dev)
    d = [torch.from_numpy(plan[key].astype(np.int64)).to(dev) for key in ("frag_begin", "frag_end", "frag_bin")]
    gf.insert_batch_dev(d_ref, d[0], d[1], d[2], len(plan["frag_bin"]), int((plan["frag_end"] - plan["frag_begin"]).max()))
    torch.cuda.synchronize()
    rb.set_insert_kernel(0)
    return d_ref


def _classify_dev(gf, bases, off, luts, chunk, torch):
    dev = torch.device("cuda", gf.device)
    n = len(off) - 1
    d_bases = torch.from_numpy(bases).to(dev)
    d_off = torch.from_numpy(off.astype(np.int64)).to(dev)
    d_lut = torch.from_numpy(luts.view(np.int16)).to(dev)
    d_keys = torch.zeros(len(luts) * n, dtype=torch.int64, device=dev)
    gf.count_batch_dev(d_bases, d_off, n, d_lut, len(luts), d_keys, max_read_len=chunk)
    torch.cuda.synchronize()
    return d_keys.cpu().numpy().view(np.uint64).reshape(len(luts), n)


def test_synth_generator_device_equals_host():
    torch = _torch()
    for start, n, seed in ((0, 1000, 3), (37, 4099, 11), (1 << 33, 513, 5)):
        d = torch.zeros(n + 64, dtype=torch.uint8, device="cuda")
        rb.capi.synth_bases_dev(d.data_ptr() + 7, n, seed, start)             # unaligned destination
        torch.cuda.synchronize()
        got = d.cpu().numpy()
        assert np.array_equal(got[7:7 + n], synth.hash_bases(start, n, seed))
        assert not got[:7].any() and not got[7 + n:].any()
    ref = synth.HashReference([1000, 777, 5000], 11)
    assert np.array_equal(ref.to_device(torch.device("cuda", 0)).cpu().numpy()[:len(ref)], ref.host())


# ---- BASELINE config #3 (human depletion) and #4 (its build) at full size -------------------------------------------
def test_config3_and_config4_full_size_vs_oracle():
    torch = _torch()
    k, frag, chunk, n = 13, 100_000, 250, 65_536
    ref = synth.HashReference([129_166_666] * 24, 300)                     # 3.1 Gb
    plan = ref.plan(frag, k)
    assert plan["n_bins"] == 31_008 and plan["bin_ids_consumed"] == 31_008
    gf = rb.IBF.create(plan["n_bins"], 3, k, plan["n_bits"])
    assert gf.bin_width == 485 and gf.n_blocks == 1_236_269
    d_ref = _build_on_gpu(gf, ref, plan, torch, variant=2)                 # column build (config #4)
    # (1) the two build kernels agree on the whole 4.8 GB matrix
    gf_red = rb.IBF.create(plan["n_bins"], 3, k, plan["n_bits"])
    rb.set_insert_kernel(1)                                                # 64-bit RED.OR per (k-mer, hash)
    d = [torch.from_numpy(plan[key].astype(np.int64)).cuda() for key in ("frag_begin", "frag_end", "frag_bin")]
    gf_red.insert_batch_dev(d_ref, d[0], d[1], d[2], len(plan["frag_bin"]), frag + k)
    torch.cuda.synchronize()
    rb.set_insert_kernel(0)
    assert torch.equal(torch.as_tensor(_DevWords(gf), device="cuda"), torch.as_tensor(_DevWords(gf_red), device="cuda"))
    gf_red.close()
    words = gf.download()
    del d_ref, d
    torch.cuda.empty_cache()
    # (2) the oracle builds the first 512 fragments on the CPU: those bin columns must be equal in ALL rows
    m = 512
    end = int(plan["frag_end"][m - 1])
    ob = oracle.OracleIBF.create(plan["n_bins"], 3, k, plan["n_bits"])
    ob.insert_batch(ref.host(end), plan["frag_begin"][:m], plan["frag_end"][:m], plan["frag_bin"][:m], n_threads=16)
    assert np.array_equal(words.reshape(-1, 485)[:, :m // 64], ob.words()[:plan["n_bits"] // 64].reshape(-1, 485)[:, :m // 64])
    ob.close()
    # (3) classify 65 536 chunks (postings kernel), both thresholds; oracle on 3 000 sampled chunks against the same matrix
    bases, off, from_ref = synth.sample_reads(ref, n, chunk, seed=1234)
    luts = np.stack([rb.threshold_lut(0.1, k), rb.threshold_lut(0.08, k)])
    keys = _classify_dev(gf, bases, off, luts, chunk, torch)
    assert gf.kmer_table_kind() == 2                                   # postings
    mx, hit, am = rb.keys_decode(keys)
    of = oracle.OracleIBF.create(plan["n_bins"], 3, k, plan["n_bits"])
    of.words()[:plan["n_bits"] // 64] = words
    del words
    pick = np.sort(np.random.default_rng(9).choice(n, SAMPLE, replace=False))
    sb = bases.reshape(n, chunk)[pick].reshape(-1)
    so = np.arange(SAMPLE + 1, dtype=np.uint64) * np.uint64(chunk)
    for t in range(2):
        exp = of.count_batch(sb, so, luts[t], dense=False, n_threads=16)
        assert np.array_equal(mx[t][pick], exp["max_count"])
        assert np.array_equal(hit[t][pick], exp["hit"])
        assert np.array_equal(am[t][pick], exp["argmax_bin"])
    assert 0.4 < hit[0][from_ref].mean() and hit[0][~from_ref].mean() < 0.05     # decisions are not all-hit here
    # (4) the host-buffer call returns the same arrays for the whole batch
    res = gf.count_batch(bases, off, luts)
    assert np.array_equal(res["max_count"], mx) and np.array_equal(res["hit"], hit) and np.array_equal(res["argmax_bin"], am)
    # (5) the streaming kernel (no table) agrees on a slice of the batch
    rb.set_count_kernel(2)
    keys_s = _classify_dev(gf, bases[:4096 * chunk], off[:4097], luts, chunk, torch)
    assert np.array_equal(keys_s, keys[:, :4096])


# ---- BASELINE config #5: one GPU's share (its own 3.74 Gb, its own 37 440-bin column slice of the 299 520-bin filter) ----
def test_config5_per_rank_shape_vs_oracle():
    torch = _torch()
    k, frag, chunk, n = 13, 100_000, 250, 65_536
    shard, n_shards, per_bins = 3, 8, 37_440
    ref = synth.HashReference([3_743_950_000], 5000 + shard)
    plan = ref.plan(frag, k, bin0=shard * per_bins, n_bins=n_shards * per_bins)
    assert plan["bin_ids_consumed"] == per_bins
    gf = rb.IBF.create_shard(plan["n_bins"], 3, k, plan["n_bits"], shard, n_shards)
    assert (gf.col_words, gf.bin_begin, gf.n_bins_local) == (585, shard * per_bins, per_bins)
    d_ref = _build_on_gpu(gf, ref, plan, torch)
    del d_ref
    torch.cuda.empty_cache()
    src = synth.HashReference([64_000_000], 5000 + shard)                  # reads come from the head of this shard's genome
    bases, off, from_ref = synth.sample_reads(src, n, chunk, seed=1234)
    luts = np.stack([rb.threshold_lut(0.1, k), rb.threshold_lut(0.08, k)])
    keys = _classify_dev(gf, bases, off, luts, chunk, torch)
    assert gf.kmer_table_kind() == 2                                   # postings
    mx, hit, am = rb.keys_decode(keys)
    # the oracle holds the same column slice as a filter of its own: same rows, local bin ids
    of = oracle.OracleIBF.create(per_bins, 3, k, gf.n_blocks * 64 * gf.col_words)
    assert of.n_blocks == gf.n_blocks and of.bin_width == gf.col_words
    of.words()[:gf.n_blocks * gf.col_words] = gf.download()[:gf.n_blocks * gf.col_words]
    pick = np.sort(np.random.default_rng(9).choice(n, SAMPLE, replace=False))
    sb = bases.reshape(n, chunk)[pick].reshape(-1)
    so = np.arange(SAMPLE + 1, dtype=np.uint64) * np.uint64(chunk)
    for t in range(2):
        exp = of.count_batch(sb, so, luts[t], dense=False, n_threads=16)
        assert np.array_equal(mx[t][pick], exp["max_count"])
        assert np.array_equal(hit[t][pick], exp["hit"])
        glob = np.where(exp["hit"] > 0, exp["argmax_bin"] + np.uint32(shard * per_bins), np.uint32(0xFFFFFFFF))
        assert np.array_equal(am[t][pick], glob)
    assert hit[0][from_ref].mean() > 0.4


# ---- concurrent callers on shared filters (adaptive_sampling.hpp:745-751: IBF.threads classify workers) -----------------
@pytest.mark.parametrize("shape", ["narrow", "wide"])
def test_many_host_threads_one_handle(shape):
    if shape == "narrow":
        ref = [synth.random_bases(60_000, 40 + i) for i in range(100)]
        plan = synth.build_plan(ref, 61_000, 13)                          # 100 bins: window-table kernels
    else:
        ref = [synth.random_bases(700_500, 70 + i) for i in range(3)]
        plan = synth.build_plan(ref, 2_000, 13)                           # ~1 050 bins: postings kernel
    gf = rb.IBF.create(plan["n_bins"], 3, 13, plan["n_bits"])
    gf.insert_batch(plan["bases"], plan["frag_begin"], plan["frag_end"], plan["frag_bin"])
    of = oracle.OracleIBF.create(plan["n_bins"], 3, 13, plan["n_bits"])
    of.words()[:plan["n_bits"] // 64] = gf.download()
    gf.enable_kmer_table(0)
    luts = np.stack([rb.threshold_lut(0.1, 13), rb.threshold_lut(0.08, 13)])
    n_threads, rounds = 8, 3
    batches, expect = [], []
    for t in range(n_threads):
        lens = np.random.default_rng(100 + t).integers(0, 600, size=1500 + 37 * t)
        b, o = synth.ragged_reads(plan["bases"], lens, seed=200 + t, frac_from_ref=0.6, n_frac=0.002)
        batches.append((b, o))
        expect.append([of.count_batch(b, o, luts[i], dense=False, n_threads=4) for i in range(2)])
    errors = []

    def worker(t):
        try:
            b, o = batches[t]
            for _ in range(rounds):
                got = gf.count_batch(b, o, luts)
                for i in range(2):
                    for key in ("max_count", "hit", "argmax_bin"):
                        if not np.array_equal(got[key][i], expect[t][i][key]):
                            raise AssertionError("thread %d threshold %d %s differs" % (t, i, key))
                if not np.array_equal(got["read_flag"], expect[t][0]["short_read"]):
                    raise AssertionError("thread %d read_flag differs" % t)
        except Exception as e:                                            # noqa: BLE001 -- reported by the main thread
            errors.append(e)

    th = [threading.Thread(target=worker, args=(t,)) for t in range(n_threads)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errors, errors[:2]


def test_one_process_two_devices_two_handles():
    """one host process drives several devices, one handle each, from its own threads."""
    if rb.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    ref = [synth.random_bases(60_000, 40 + i) for i in range(100)]
    plan = synth.build_plan(ref, 61_000, 13)
    of = oracle.OracleIBF.create(plan["n_bins"], 3, 13, plan["n_bits"])
    of.insert_batch(plan["bases"], plan["frag_begin"], plan["frag_end"], plan["frag_bin"], n_threads=8)
    lut = rb.threshold_lut(0.1, 13)
    handles = []
    for dev in range(2):
        gf = rb.IBF.create(plan["n_bins"], 3, 13, plan["n_bits"], device=dev)
        gf.insert_batch(plan["bases"], plan["frag_begin"], plan["frag_end"], plan["frag_bin"])
        assert gf.device == dev and np.array_equal(gf.download(), of.words()[:plan["n_bits"] // 64])
        handles.append(gf)
    errors = []

    def worker(dev):
        try:
            b, o, _ = synth.sample_reads(plan["bases"], 20_000, 250, seed=50 + dev)
            exp = of.count_batch(b, o, lut, dense=False, n_threads=4)
            for _ in range(3):
                got = handles[dev].count_batch(b, o, lut)
                for key in ("max_count", "hit", "argmax_bin"):
                    if not np.array_equal(got[key], exp[key]):
                        raise AssertionError("device %d %s differs" % (dev, key))
        except Exception as e:                                            # noqa: BLE001
            errors.append(e)

    th = [threading.Thread(target=worker, args=(d,)) for d in range(2)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errors, errors[:2]


def test_bench_default_line_shape_with_secondary():
    """bench.py with a `secondary` list (the default run carries configs #3 / k = 15 / k = 17; here their mini twins): every
    entry has value, roofline.frac from measured-geometry DRAM bytes, the parity record, and the one-off costs."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--workload", "mini_100x60kb_100bins", "--steps", "3",
                          "--warmup", "3", "--secondary", "mini3_40Mb_408bins,mini_k15"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    d = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][0])
    assert [s["config"]["workload"] for s in d["secondary"]] == ["mini3_40Mb_408bins", "mini_k15"]
    for e in [d] + d["secondary"]:
        r = e["roofline"]
        assert r["traffic"] > 0 and abs(r["frac"] - r["traffic"] / (r["kernel_ms"] * 1e-3) / 1e9 / r["peak"]) < 1e-9
        assert "oracle" in e["parity"] and "host_call" in e["parity"]
        assert e["config"]["cold_first_call_ms"] > 0 and e["e2e"]["value"] > 0
    # 408 bins = 7 row words: the group-loaded k-mer table (ibf_ctable.cu); config #3's own postings path is covered at full size above
    assert d["secondary"][0]["roofline"]["kernel"] == "count_ctable_kernel" and "build" in d["secondary"][0]["parity"]
    assert d["e2e"]["h2d_ceiling_gbs"] > 1
