"""world_size=2 gloo tests (CPU) of the multi-GPU host logic: read sharding needs no collective and
reassembles exactly; bin sharding combines per-read keys with one all-reduce(MAX) and reproduces the
whole-filter summaries.  Per-shard results come from the oracle (the GPU kernels are covered by -m gpu)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from readbouncer_b200 import capi, dist as rbdist, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _setup():
    ref = [synth.random_bases(3000, 50 + i) for i in range(300)]       # 300 bins -> 5 row words
    of, stats = oracle.build_from_sequences(ref, 4000, k=13)
    plan_bases = np.concatenate([np.frombuffer(oracle.cut_out_nnns(r.tobytes()), np.uint8) for r in ref])
    bases, off = synth.ragged_reads(plan_bases, [250] * 50 + [5, 300, 13, 0, 777], seed=4, frac_from_ref=0.8)
    lut = oracle.threshold_lut(0.1, 13)
    return of, bases, off, lut


def _keys_from_dense(cf, cr, lut, lens, bin_lo, bin_hi):
    """What a bin shard's kernel reports: packed key over the shard's bins only."""
    keys = np.zeros(len(lens), np.uint64)
    for i, L in enumerate(lens):
        if L < 13 or L > 65535:
            continue
        thr = lut[L]
        f, r = cf[i, bin_lo:bin_hi].astype(np.int64), cr[i, bin_lo:bin_hi].astype(np.int64)
        ok = (f >= thr) | (r >= thr)
        if ok.any():
            m = np.maximum(f, r)
            best = m[ok].max()
            b = bin_lo + int(np.nonzero(ok & (m == best))[0][0])
            keys[i] = (1 << 48) | (int(best) << 32) | ((~b) & 0xFFFFFFFF)
    return keys


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    of, bases, off, lut = _setup()
    n = len(off) - 1
    lens = np.diff(off.astype(np.int64))
    full = of.count_batch(bases, off, lut, dense=True)

    # ---- read-sharded: classify my slice only, then reassemble -------------------------------------
    lo, hi = rbdist.shard_range(n, rank, world)
    my_off = off[lo:hi + 1] - off[lo]
    my_bases = bases[int(off[lo]):int(off[hi])]
    mine = of.count_batch(my_bases, my_off, lut, dense=False)
    got = rbdist.gather_results(torch.from_numpy(mine["max_count"].astype(np.int32)), n)
    ok_read = np.array_equal(got.numpy(), full["max_count"].astype(np.int32))
    got_hit = rbdist.gather_results(torch.from_numpy(mine["hit"].copy()), n)
    ok_read &= np.array_equal(got_hit.numpy(), full["hit"])

    # ---- bin-sharded: all reads vs my bin columns, one all-reduce(MAX) of the keys -----------------
    cb, cw = rbdist.bin_shard_columns(of.bin_width, rank, world)
    bin_lo, bin_hi = 64 * cb, min(of.n_bins, 64 * (cb + cw))
    keys = _keys_from_dense(full["counts_fwd"], full["counts_rev"], lut, lens, bin_lo, bin_hi)
    t = torch.from_numpy(keys.view(np.int64).copy())
    rbdist.combine_keys(t)
    mx, hit, am = capi.keys_decode(t.numpy().view(np.uint64))
    ok_bin = (np.array_equal(mx, full["max_count"]) and np.array_equal(hit, full["hit"])
              and np.array_equal(am, full["argmax_bin"]))
    ret[rank] = (bool(ok_read), bool(ok_bin), int(full["hit"].sum()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_read_and_bin_sharding():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert len(ret) == world
        for r in range(world):
            ok_read, ok_bin, hits = ret[r]
            assert ok_read and ok_bin and hits > 10


def test_shard_arithmetic():
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            spans = [rbdist.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    for W in (5, 8, 485, 4692):
        for w in (1, 2, 4, 8):
            cols = [rbdist.bin_shard_columns(W, r, w) for r in range(w)]
            if W >= w:
                assert sum(c for _, c in cols) == W and cols[0][0] == 0
                assert all(cols[i][0] + cols[i][1] == cols[i + 1][0] for i in range(w - 1))


def test_per_rank_built_slices_start_at_their_own_bins():
    """BASELINE config #5 built slice by slice: with a multiple of 64 bins per rank, rank r's slice holds exactly its bins."""
    from readbouncer_b200 import dist as rbdist
    for per in (64, 512, 37440):
        for world in (1, 2, 3, 4, 8):
            rng = rbdist.per_rank_bin_ranges(per, world)
            assert [lo for lo, _ in rng] == [r * per for r in range(world)]
            assert [hi for _, hi in rng] == [(r + 1) * per for r in range(world)]
    # 8 x 37 440 bins = the 30 Gb filter of config #5: 4 681 row words, 585 (the last 586) per GPU
    assert rbdist.bin_shard_columns(8 * 37440 // 64 + 1, 7, 8) == (4095, 586)
