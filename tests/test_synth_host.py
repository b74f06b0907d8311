"""CPU tests of the host-side workload helpers the GPU tests and bench.py stand on: the position-hash reference generator
(numpy twin of rb_synth_bases_dev), HashReference plans against the generic build plan, the unchecked threshold table, and
the shapes of the bench workloads."""
import os
import sys

import numpy as np

import readbouncer_b200 as rb
from readbouncer_b200 import dist, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_hash_bases_is_a_pure_function_of_seed_and_position():
    a = synth.hash_bases(0, 5000, 7)
    assert set(np.unique(a)) <= set(b"ACGT") and len(a) == 5000
    for start, n in ((0, 1), (31, 2), (32, 64), (37, 1000), (4999, 1)):
        assert np.array_equal(synth.hash_bases(start, n, 7), a[start:start + n])
    assert not np.array_equal(synth.hash_bases(0, 5000, 8), a)
    big = synth.hash_bases((1 << 33) + 5, 100, 3)                     # positions beyond 2^32
    assert np.array_equal(big[10:20], synth.hash_bases((1 << 33) + 15, 10, 3))
    counts = np.bincount(synth.hash_bases(0, 400_000, 11), minlength=256)[[65, 67, 71, 84]]
    assert counts.min() > 98_000 and counts.max() < 102_000


def test_hash_reference_windows_plan_and_sampling():
    ref = synth.HashReference([30_000, 12_345, 50_001], 40)
    host = ref.host()
    assert len(ref) == len(host) == 92_346 and np.array_equal(ref.host(20_000), host[:20_000])
    assert np.array_equal(ref.host(35_000), host[:35_000])
    pos = np.array([0, 29_900, 30_000, 42_300, 92_000])
    w = ref.windows(pos, 300)
    # a window that would cross the end of its sequence is moved back to end there
    for p, row in zip([0, 29_700, 30_000, 42_045, 92_000], w):
        assert np.array_equal(row, host[p:p + 300])
    # the plan on lengths alone == the generic plan on the raw records (raw = kept bases + the base cutOutNNNs drops)
    raw = [np.concatenate([host[o:o + n], np.frombuffer(b"G", np.uint8)]) for o, n in zip(ref.offsets[:-1], ref.lengths)]
    p1, p2 = ref.plan(10_000, 13), synth.build_plan(raw, 10_000, 13)
    for key in ("frag_begin", "frag_end", "frag_bin"):
        assert np.array_equal(p1[key], p2[key]), key
    assert (p1["n_bins"], p1["n_bits"], p1["bin_ids_consumed"]) == (p2["n_bins"], p2["n_bits"], p2["bin_ids_consumed"])
    assert np.array_equal(p2["bases"], host)
    shard = ref.plan(10_000, 13, bin0=64, n_bins=128)
    assert shard["frag_bin"][0] == 64 and shard["n_bins"] == 128 and shard["n_bits"] == rb.ibf_size_bits(10_000, 13, 3, 0.01, 128)
    b1, o1, f1 = synth.sample_reads(ref, 500, 250, seed=3)
    b2, o2, f2 = synth.sample_reads(host, 500, 250, seed=3)
    assert np.array_equal(o1, o2) and np.array_equal(f1, f2) and b1.shape == b2.shape
    assert (b1.reshape(500, 250)[~f1] == b2.reshape(500, 250)[~f2]).all()     # the iid reads do not depend on the reference object


def test_unchecked_threshold_table():
    for k in (13, 15):
        assert np.array_equal(rb.threshold_lut(0.1, k), rb.threshold_lut(0.1, k, raw=True))
        assert np.array_equal(rb.threshold_lut(0.08, k), rb.threshold_lut(0.08, k, raw=True))
    zero = rb.threshold_lut(0.0, 13, raw=True)          # the reference's retry at error_rate - 0.02 with error_rate = 0.02
    assert zero[250] == 238 and zero[13] == 1           # NaN interval bound casts to 0: every k-mer must match
    rb.threshold_lut(-0.01, 13, raw=True)               # no exception either
    try:
        rb.threshold_lut(0.0, 13)
        raise AssertionError("the checked variant must refuse a rate of 0")
    except rb.RBError as e:
        assert e.status == 8


def test_bench_workloads_consume_exactly_their_bins():
    sys.path.insert(0, ROOT)
    import bench
    for name, w in bench.WORKLOADS.items():
        if w.get("per_rank"):
            for world in (1, 2, 8):
                ref = bench.make_reference(w, world - 1)
                per = w["seq_len"] // w["fragment"] + 1
                p = ref.plan(w["fragment"], w["k"], bin0=(world - 1) * per, n_bins=world * per)
                assert per == w["bins_per_rank"] and per % 64 == 0 and p["bin_ids_consumed"] == per, name
                assert dist.per_rank_bin_ranges(per, world)[world - 1] == ((world - 1) * per, world * per)
        else:
            p = bench.make_reference(w).plan(w["fragment"], w["k"])
            assert p["bin_ids_consumed"] == p["n_bins"], name      # no quirk-Q3 overrun in the bench references
    assert set(bench.SECONDARY_1GPU) - {"readme_3targets_1deplete", "live_3targets_1deplete"} <= set(bench.WORKLOADS)
    assert set(bench.SECONDARY_NGPU) <= set(bench.WORKLOADS)
