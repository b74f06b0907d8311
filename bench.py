#!/usr/bin/env python3
"""bench.py -- throughput of the IBF classify hot path on B200 (BASELINE.json metric).

One "step" = one pass of the hot path (2 x seqan::count + threshold + select/max_matches per
chunk, src/IBF/IBFClassify.cpp:138-171) over one batch of synthetic 250-base read chunks per GPU.
Default workload = BASELINE config #2: 1 M chunks vs a 100-bin IBF built (on the GPU, by the
insert kernel) from 100 synthetic 4 Mb genomes.  Multi-GPU = read-sharded with a replicated IBF
(no data-path collective, weak scaling); `--mode bin_sharded` shards the bit matrix by bin
columns and combines per-read summary keys with one NCCL all-reduce(MAX).

  python bench.py --gpus 1 --steps 10 --warmup 3
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...      # the CPU oracle port on all host cores (reference arm)

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "classified_250b_read_chunks_per_sec"
UNIT = "chunks/s"

WORKLOADS = {
    # name: (n_seqs, seq_len, fragment_size, k, chunk_len, default chunks per GPU per step)
    "cfg1_5Mb_51bins": dict(n_seqs=1, seq_len=5_000_001, fragment=100_000, k=13, chunk=250, reads=1_000_000),
    "cfg2_100x4Mb_100bins": dict(n_seqs=100, seq_len=4_000_000, fragment=4_200_000, k=13, chunk=250, reads=1_000_000),
    "cfg3_3.1Gb_31kbins": dict(n_seqs=24, seq_len=129_166_667, fragment=100_000, k=13, chunk=250, reads=65_536),
    "w1_50x4Mb_50bins": dict(n_seqs=50, seq_len=4_000_000, fragment=4_200_000, k=13, chunk=250, reads=1_000_000),
    "mini_100x60kb_100bins": dict(n_seqs=100, seq_len=60_000, fragment=61_000, k=13, chunk=250, reads=65_536),
    # BASELINE config #5 (30 Gb, ~300 k bins, 8 bin shards), one eighth per GPU: every rank generates ITS OWN 3.74 Gb genome
    # group, builds ITS OWN column slice (37 440 bins = 585 row words, 5.8 GB) and never sees the rest of the filter;
    # N ranks = an N x 37 440-bin filter over N x 3.74 Gb.  Needs --mode bin_sharded semantics (implied).
    "cfg5_3.7Gb_37kbins_per_gpu": dict(per_rank=True, seq_len=3_743_950_001, window=64_000_000, fragment=100_000, k=13, chunk=250,
                                       reads=65_536, bins_per_rank=37_440),
    "mini5_40Mb_512bins_per_gpu": dict(per_rank=True, seq_len=51_150_001, window=8_000_000, fragment=100_000, k=13, chunk=250,
                                       reads=16_384, bins_per_rank=512),
}
DEFAULT_WORKLOAD = "cfg2_100x4Mb_100bins"
ERROR_RATE = 0.1
SIGNIFICANCE = 0.95


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--reads", type=int, default=0, help="chunks per GPU per step (0 = workload default)")
    ap.add_argument("--mode", default="read_sharded", choices=["read_sharded", "bin_sharded"])
    ap.add_argument("--kernel", type=int, default=0, help="0 auto, 1 tile, 2 stream, 3 k-mer table")
    ap.add_argument("--l2-gran", type=int, default=0, help="cudaLimitMaxL2FetchGranularity (0 = library default)")
    ap.add_argument("--from-ref", type=float, default=0.5,
                    help="fraction of the chunks sampled from the reference with 10 %% errors (rest iid); 0 = all negative")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def make_reference(w):
    from readbouncer_b200 import synth
    return [synth.random_bases(w["seq_len"], 2 + i) for i in range(w["n_seqs"])]


def algorithmic_bytes_per_chunk(w, bin_width, n_hash=3):
    lookups = 2 * (w["chunk"] - w["k"] + 1)
    return lookups, lookups * n_hash * bin_width * 8


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 100 ms while the GPU is under the bench load."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def n_samples(self):
        try:
            return sum(1 for ln in open(self.f.name) if ln.count(",") >= 8)
        except OSError:
            return 0

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [ln.split(",") for ln in open(self.f.name).read().strip().splitlines() if ln.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return out
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            for name, v in zip(names, r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=float(rows[0][2]), reasons=sorted(reasons), samples=len(rows),
                   power_w_max=max(float(r[3]) for r in rows))
        return out


# ------------------------------------------------------------------------------------------------------
def run_reference(args, w, n_reads):
    """Reference arm: the CPU restatement of the reference's classify path (oracle port; the
    reference binary cannot be compiled here), all host threads, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    from readbouncer_b200 import synth     # numpy read sampler only; no library call on this arm
    cores = os.cpu_count() or 1
    ref = make_reference(w)
    t0 = time.time()
    of, stats = oracle.build_from_sequences(ref, w["fragment"], k=w["k"], n_threads=cores)
    build_s = time.time() - t0
    plan = {"bases": np.concatenate([np.frombuffer(oracle.cut_out_nnns(s.tobytes()), np.uint8) for s in ref]),
            "n_bins": of.n_bins, "n_bits": of.n_bits}
    del ref
    lut = oracle.threshold_lut(ERROR_RATE, w["k"], SIGNIFICANCE)
    bases, off, _ = synth.sample_reads(plan["bases"], min(n_reads, 400_000), w["chunk"], seed=1234, frac_from_ref=args.from_ref)
    n_avail = len(off) - 1
    # calibrate, then size each step to ~6 s of CPU work so the whole run ends within a few minutes
    cal = min(2000, n_avail)
    t0 = time.time()
    of.count_batch(bases[:cal * w["chunk"]], off[:cal + 1], lut, dense=False, n_threads=cores)
    rate = cal / max(time.time() - t0, 1e-6)
    sample = int(max(cal, min(n_avail, rate * 6.0)))
    sb, so = bases[:sample * w["chunk"]], off[:sample + 1]
    for _ in range(args.warmup):
        of.count_batch(sb, so, lut, dense=False, n_threads=cores)
    t0 = time.time()
    for _ in range(args.steps):
        of.count_batch(sb, so, lut, dense=False, n_threads=cores)
    dt = time.time() - t0
    value = sample * args.steps / dt
    sample_desc = "first %d of the %d-chunk batch per step, %d threads; IBF built on CPU in %.1f s" % (
        sample, n_reads, cores, build_s)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": args.workload, "chunks_per_step_sample": sample, "chunk_length": w["chunk"], "kmer_size": w["k"],
                   "bins": plan["n_bins"], "filter_bytes": plan["n_bits"] // 8, "error_rate": ERROR_RATE,
                   "read_mix": "%g%% reference-derived @10%% errors, %g%% iid" % (100 * args.from_ref, 100 - 100 * args.from_ref)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample_desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------------
def run_ours(args, w, n_reads):
    import torch
    import torch.distributed as dist
    import readbouncer_b200 as rb
    from readbouncer_b200 import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or rb.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rb.set_count_kernel(args.kernel)
    if args.l2_gran:
        rb.set_l2_fetch_granularity(args.l2_gran, device=local)
    stream = torch.cuda.current_stream()
    per_rank = bool(w.get("per_rank"))
    bin_sharded = (args.mode == "bin_sharded" and world > 1) or per_rank

    # ---- build the IBF on the GPU (insert kernel) ------------------------------------------------------
    if per_rank:
        # the first `window` bases of every rank's genome group are known to all ranks (the reads are sampled there, so
        # all ranks classify the same batch); the rest is generated by its owner only
        windows = [synth.random_bases(w["window"], 5000 + g) for g in range(world)]
        seq = np.concatenate([windows[rank], synth.random_bases(w["seq_len"] - w["window"], 6000 + rank)])
        bases_own = seq[:-1]                 # cutOutNNNs drops the last base of a sequence without trailing N (quirk Q1)
        del seq
        fb, fe = rb.capi.fragment_schedule(len(bases_own), w["fragment"], w["k"])
        per_bins = len(bases_own) // w["fragment"] + 1              # IBFBuild.cpp:90
        assert per_bins == w["bins_per_rank"] == len(fb) and per_bins % 64 == 0
        n_bins = world * per_bins
        plan = {"bases": bases_own, "frag_begin": fb, "frag_end": fe,
                "frag_bin": np.arange(rank * per_bins, (rank + 1) * per_bins, dtype=np.uint64), "n_bins": n_bins,
                "n_bits": rb.ibf_size_bits(w["fragment"], w["k"], 3, 0.01, n_bins)}
        gf_full = rb.IBF.create_shard(n_bins, 3, w["k"], plan["n_bits"], rank, world, device=local)
        assert gf_full.bin_begin == rank * per_bins and gf_full.n_bins_local >= per_bins
    else:
        ref = make_reference(w)
        plan = synth.build_plan(ref, w["fragment"], w["k"])
        del ref
        gf_full = rb.IBF.create(plan["n_bins"], 3, w["k"], plan["n_bits"], device=local)
    d_ref = torch.from_numpy(plan["bases"]).to(dev)
    d_fb = torch.from_numpy(plan["frag_begin"].astype(np.int64)).to(dev)
    d_fe = torch.from_numpy(plan["frag_end"].astype(np.int64)).to(dev)
    d_fbin = torch.from_numpy(plan["frag_bin"].astype(np.int64)).to(dev)
    max_frag = int((plan["frag_end"] - plan["frag_begin"]).max())
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # the same fragments twice (inserts are idempotent ORs, so the bits are those of one build): the first call pays the
    # one-off costs of the process (memory-pool growth for the scratch columns, kernel attributes), the second is timed
    build_cold_ms = None
    for it in range(2):
        torch.cuda.synchronize()
        ev0.record(stream)
        gf_full.insert_batch_dev(d_ref, d_fb, d_fe, d_fbin, len(plan["frag_bin"]), max_frag, stream=stream)
        ev1.record(stream)
        torch.cuda.synchronize()
        build_ms = ev0.elapsed_time(ev1)
        if it == 0:
            build_cold_ms = build_ms
    n_kmers_ref = int(np.maximum(plan["frag_end"] - plan["frag_begin"], w["k"] - 1).sum() - (w["k"] - 1) * len(plan["frag_bin"]))
    del d_ref, d_fb, d_fe, d_fbin
    gf = gf_full
    full_keys_check = None
    if per_rank:
        plan["bases"] = np.concatenate(windows)       # what the reads are sampled from (same on all ranks)
        del windows, bases_own
    elif bin_sharded:
        full_keys_check = True
        words = gf_full.download()
        gf = rb.IBF.from_words(words, plan["n_bins"], 3, w["k"], plan["n_bits"], device=local, shard=rank, n_shards=world)
        del words

    # ---- reads: host (pinned) and device copies --------------------------------------------------------
    # read-sharded: every rank classifies its own batch; bin-sharded: all ranks see the same batch
    seed = 1234 if bin_sharded else 1234 + rank
    bases_np, off_np, from_ref = synth.sample_reads(plan["bases"], n_reads, w["chunk"], seed=seed, frac_from_ref=args.from_ref)
    h_bases = torch.empty(bases_np.size, dtype=torch.uint8, pin_memory=True)
    h_bases.numpy()[:] = bases_np
    h_off = torch.empty(off_np.size, dtype=torch.int64, pin_memory=True)
    h_off.numpy()[:] = off_np.astype(np.int64)
    luts_np = np.stack([rb.threshold_lut(ERROR_RATE, w["k"], SIGNIFICANCE),
                        rb.threshold_lut(ERROR_RATE - 0.02, w["k"], SIGNIFICANCE)])
    n_lut = 2                                     # both thresholds of check_unblock in one pass
    d_bases = h_bases.to(dev, non_blocking=True)
    d_off = h_off.to(dev, non_blocking=True)
    d_lut = torch.from_numpy(luts_np.view(np.int16)).to(dev)
    d_keys = torch.zeros(n_lut * n_reads, dtype=torch.int64, device=dev)
    d_max = torch.zeros(n_lut * n_reads, dtype=torch.int16, device=dev)
    d_hit = torch.zeros(n_lut * n_reads, dtype=torch.uint8, device=dev)
    d_amax = torch.zeros(n_lut * n_reads, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    lib = rb.lib()

    def step():
        gf.count_batch_dev(d_bases, d_off, n_reads, d_lut, n_lut, d_keys, max_read_len=w["chunk"], stream=stream)
        if bin_sharded and world > 1:
            dist.all_reduce(d_keys, op=dist.ReduceOp.MAX)       # keys < 2^49: int64 MAX == uint64 MAX
        rb.capi._check(lib.rb_keys_decode_dev(rb.capi._dev_ptr(d_keys), n_lut * n_reads, rb.capi._dev_ptr(d_max),
                                              rb.capi._dev_ptr(d_hit), rb.capi._dev_ptr(d_amax), local,
                                              rb.capi._stream_ptr(stream)))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value) ------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    # clocks / throttle reasons: nvidia-smi every 100 ms from here to the end of the timed regions (device-resident and
    # end-to-end).  A timed region of a few tens of ms would see at most one sample, so the GPU is kept under the same
    # load (untimed steps) until the first samples have arrived.
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        t_wait = time.time()
        while sampler.n_samples() < 3 and time.time() - t_wait < 3.0:
            gf.count_batch_dev(d_bases, d_off, n_reads, d_lut, n_lut, d_keys, max_read_len=w["chunk"], stream=stream)
            torch.cuda.synchronize()              # no collective here: the other ranks wait at the barrier below
    barrier()
    if per_rank:
        # nobody holds the whole filter: the combined answer of a reference-derived chunk must name a bin of the sampling
        # windows (the first window/fragment + 1 bins of a rank's range; 1.7 % of the bins by chance)
        amax = d_amax[:n_reads].cpu().numpy().astype(np.int64)
        hit = d_hit[:n_reads].cpu().numpy() > 0
        sel = hit & from_ref
        in_window = (amax[sel] % w["bins_per_rank"]) <= w["window"] // w["fragment"] + 1
        assert sel.sum() > 0.9 * from_ref.sum() and in_window.mean() > 0.99, (sel.sum(), from_ref.sum(), in_window.mean())
    elif bin_sharded:
        # the all-reduced shard keys must equal the keys of the whole (replicated) filter
        d_full = torch.zeros_like(d_keys)
        gf_full.count_batch_dev(d_bases, d_off, n_reads, d_lut, n_lut, d_full, max_read_len=w["chunk"], stream=stream)
        torch.cuda.synchronize()
        assert torch.equal(d_full, d_keys), "bin-sharded combine differs from the whole-filter result"
        del d_full
        gf_full.close()
        torch.cuda.empty_cache()
    launches0 = rb.kernel_launches()
    k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_ev0, t_ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_ev0.record(stream)
    for i in range(args.steps):
        k_ev[i][0].record(stream)
        gf.count_batch_dev(d_bases, d_off, n_reads, d_lut, n_lut, d_keys, max_read_len=w["chunk"], stream=stream)
        k_ev[i][1].record(stream)
        if bin_sharded and world > 1:
            dist.all_reduce(d_keys, op=dist.ReduceOp.MAX)
        rb.capi._check(lib.rb_keys_decode_dev(rb.capi._dev_ptr(d_keys), n_lut * n_reads, rb.capi._dev_ptr(d_max),
                                              rb.capi._dev_ptr(d_hit), rb.capi._dev_ptr(d_amax), local,
                                              rb.capi._stream_ptr(stream)))
    t_ev1.record(stream)
    barrier()
    launches = rb.kernel_launches() - launches0
    total_ms = t_ev0.elapsed_time(t_ev1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in k_ev]))
    t = torch.tensor([total_ms, kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kernel_ms = float(t[0]), float(t[1])
    hits_dev = int((d_hit[:n_reads] > 0).sum())

    # ---- end-to-end timing through the host-buffer C ABI (H2D + kernels + D2H every step) ------------------
    e2e = None
    if not args.no_e2e and not bin_sharded:       # (bin shards: the host call returns decoded arrays, the combine needs keys)
        hb, ho = h_bases.numpy(), h_off.numpy().view(np.uint64)
        res_max = torch.empty(n_lut * n_reads, dtype=torch.int16, pin_memory=True).numpy().view(np.uint16)
        res_hit = torch.empty(n_lut * n_reads, dtype=torch.uint8, pin_memory=True).numpy()
        res_am = torch.empty(n_lut * n_reads, dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
        res_flag = torch.empty(n_reads, dtype=torch.uint8, pin_memory=True).numpy()

        def e2e_step():
            rb.capi._check(lib.rb_ibf_count_batch(gf._h, rb.capi._np_ptr(hb), rb.capi._np_ptr(ho), n_reads,
                                                  rb.capi._np_ptr(luts_np), n_lut, None, None,
                                                  rb.capi._np_ptr(res_max), rb.capi._np_ptr(res_hit),
                                                  rb.capi._np_ptr(res_am), rb.capi._np_ptr(res_flag),
                                                  rb.capi._stream_ptr(stream)))
        for _ in range(max(args.warmup, 5)):       # the library times both ways in (2 calls each) and keeps the faster
            e2e_step()
        barrier()
        xfer0 = rb.transfer_bytes()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        xfer1 = rb.transfer_bytes()
        te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te[0])
        assert int((res_hit[:n_reads] > 0).sum()) == hits_dev, "host-API results differ from device-API results"
        # bytes that crossed PCIe, counted by the library at its copy calls: the host threads turn the ASCII bases into
        # 3 bit planes before the transfer, so this is less than the size of the host input
        e2e = {"value": world * n_reads * args.steps / e2e_s, "unit": UNIT,
               "h2d_bytes_per_step": (xfer1[0] - xfer0[0]) // args.steps,
               "d2h_bytes_per_step": (xfer1[1] - xfer0[1]) // args.steps,
               "host_input_bytes_per_step": int(hb.nbytes + ho.nbytes + luts_np.nbytes),
               "host_result_bytes_per_step": int(res_max.nbytes + res_hit.nbytes + res_am.nbytes + res_flag.nbytes),
               "ms_per_step": 1000 * e2e_s / args.steps, "api": "rb_ibf_count_batch (host buffers, pinned)",
               "host_pack": dict(rb.host_pack_info(), enabled=os.environ.get("RB_HOST_PACK", "1") != "0"),
               "transfer_policy": gf.transfer_policy()}

    clocks = sampler.stop() if sampler else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (count kernel) ----------------------------------------------------
    lookups, bytes_per_chunk = algorithmic_bytes_per_chunk(w, gf.col_words)
    peak, peak_src = measured_peaks()
    achieved = n_reads * bytes_per_chunk / (kernel_ms * 1e-3) / 1e9
    span = gf.kmer_table_span()
    group_ok = span >= 2 and w["chunk"] - w["k"] + 1 <= 127 * span and w["chunk"] <= 545   # ibf_wtable.cu: wgroup_applicable
    kernel_name = ("count_postings_kernel" if gf.kmer_table_kind() == 2 else "count_wgroup_kernel" if group_ok else "count_wtable_kernel" if span >= 2 else "count_table_kernel" if span == 1 else
                   "count_stream_kernel" if (gf.col_words > 4 or args.kernel == 2) else "count_tile_kernel")
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):          # measured once under ncu --set full for this workload and kernel
        ent = json.load(open(tp)).get(args.workload, {})
        if ent.get("kernel") == kernel_name and ent.get("chunks_per_launch") == n_reads and world == 1:
            traffic = ent.get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": kernel_name,
                "kernel_ms": kernel_ms, "algorithmic_bytes_per_chunk": bytes_per_chunk, "peak_source": peak_src,
                "kmer_lookups_per_s": (n_reads if bin_sharded else world * n_reads) * lookups / (total_ms * 1e-3 / args.steps)}
    if traffic:     # what the kernel really moves (table entries / postings lists, not the reference's row probes) against HBM peak
        roofline["dram"] = {"bytes_per_launch": traffic, "bytes_per_chunk": traffic / n_reads,
                            "achieved": traffic / (kernel_ms * 1e-3) / 1e9, "unit": "GB/s",
                            "frac": traffic / (kernel_ms * 1e-3) / 1e9 / peak,
                            "how": "ncu dram bytes of one launch (profiles/traffic.json) / this run's kernel time"}
    # Narrow rows are bound by memory REQUESTS, not bytes: measure the box's random-gather ceiling
    # over the very buffer the kernel reads, with the kernel's request shape, and report requests/s against it.
    row_bytes = int(gf.col_words * 8)
    sink = torch.zeros(1, dtype=torch.int64, device=dev)
    blocks = 148 * 8

    def time_gather(fn):
        for it in range(3):
            if it == 1:
                torch.cuda.synchronize(); ev0.record(stream)
            fn()
        ev1.record(stream)
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) / 2

    npos = w["chunk"] - w["k"] + 1
    if gf.kmer_table_kind() == 2:
        pass                # postings lists are streamed: the HBM-bandwidth roofline above applies
    elif span >= 2:     # window table: one request per entry of `lanes` slots, adjacent lanes
        lanes = 2 if span == 2 else 4
        entry_bytes = lanes * 16 * int(gf.col_words)
        n_rows = gf.kmer_table_bytes() // entry_bytes
        ppg = 64 * lanes
        g_ms = time_gather(lambda: rb.microbench_gather_coop(gf.device_kmer_table_ptr(), n_rows, entry_bytes,
                                                              16 * int(gf.col_words), ppg, blocks, sink, stream=stream))
        peak_req = blocks * 256 // lanes * ppg / (g_ms * 1e-3)
        req_per_chunk = -(-npos // span)
        roofline["requests"] = {"what": "one %d-byte table entry (%d k-mer positions, both strands) per request" % (entry_bytes, span),
                                "per_chunk": req_per_chunk, "peak_per_s": peak_req,
                                "achieved_per_s": n_reads * req_per_chunk / (kernel_ms * 1e-3),
                                "how": "rb_microbench_gather_coop over the k-mer table itself (%d entries)" % n_rows}
    elif span == 1:
        entry_bytes = 16 * int(gf.col_words)
        n_rows = gf.kmer_table_bytes() // entry_bytes
        g_ms = time_gather(lambda: rb.microbench_gather(gf.device_kmer_table_ptr(), n_rows, entry_bytes, 256, blocks, sink, stream=stream)) \
            if entry_bytes in (16, 32, 64) else None
        if g_ms:
            roofline["requests"] = {"what": "one %d-byte table entry (1 k-mer position, both strands) per request" % entry_bytes,
                                    "per_chunk": npos, "peak_per_s": blocks * 256 * 256 / (g_ms * 1e-3),
                                    "achieved_per_s": n_reads * npos / (kernel_ms * 1e-3),
                                    "how": "rb_microbench_gather over the k-mer table itself (%d entries)" % n_rows}
    elif row_bytes in (8, 16, 32):
        g_ms = time_gather(lambda: rb.microbench_gather(gf.device_words_ptr(), gf.n_blocks, row_bytes, 256, blocks, sink, stream=stream))
        roofline["requests"] = {"what": "one %d-byte row probe per request" % row_bytes, "per_chunk": lookups * 3,
                                "peak_per_s": blocks * 256 * 256 / (g_ms * 1e-3),
                                "achieved_per_s": n_reads * lookups * 3 / (kernel_ms * 1e-3),
                                "how": "rb_microbench_gather over the same %d-row matrix" % gf.n_blocks}
    if "requests" in roofline:
        roofline["requests"]["frac"] = roofline["requests"]["achieved_per_s"] / roofline["requests"]["peak_per_s"]

    # ---- CPU baseline: the oracle port on this box's cores, bounded sample (N=1 only) ----------------------
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        import oracle
        cores = os.cpu_count() or 1
        of = oracle.OracleIBF.create(plan["n_bins"], 3, w["k"], plan["n_bits"])
        of.words()[:plan["n_bits"] // 64] = gf.download()
        lut0 = luts_np[0]
        cal = min(2000, n_reads)
        t0 = time.time()
        exp = of.count_batch(bases_np[:cal * w["chunk"]], off_np[:cal + 1], lut0, dense=False, n_threads=cores)
        rate = cal / max(time.time() - t0, 1e-6)
        got_hit = d_hit[:cal].cpu().numpy()
        assert np.array_equal(got_hit, exp["hit"]), "GPU decisions differ from the oracle on the bench batch"
        sample = int(max(cal, min(n_reads, rate * 15.0)))
        sb, so = bases_np[:sample * w["chunk"]], off_np[:sample + 1]
        # 10-20 s of CPU work: the sample is classified again until 12 s have passed (a batch of 1 M chunks takes ~4 s)
        passes, dt = 0, 0.0
        while passes == 0 or (dt < 12.0 and passes < 64):
            t0 = time.time()
            exp = of.count_batch(sb, so, lut0, dense=False, n_threads=cores)
            dt += time.time() - t0
            passes += 1
        assert np.array_equal(d_max[:sample].cpu().numpy().view(np.uint16), exp["max_count"])
        # one host thread, and the reference's call shape (one read per call: IBFClassify.cpp:138-171 is entered per read)
        n1 = int(max(64, min(sample, rate / cores * 3.0)))
        t0 = time.time()
        of.count_batch(sb[:n1 * w["chunk"]], so[:n1 + 1], lut0, dense=False, n_threads=1)
        dt1 = time.time() - t0
        n2 = min(n1, 2000)
        t0 = time.time()
        for i in range(n2):
            of.count_batch(sb[i * w["chunk"]:(i + 1) * w["chunk"]], so[:2], lut0, dense=False, n_threads=1)
        dt2 = time.time() - t0
        cpu_baseline = {"value": sample * passes / dt, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "first %d chunks of the step's batch x %d passes, %d threads, %.1f s; results equal the GPU's" % (
                            sample, passes, cores, dt),
                        "single_thread": {"value": n1 / dt1, "chunks": n1},
                        "one_read_per_call_single_thread": {"value": n2 / dt2, "chunks": n2}}

    # read-sharded: every rank classifies its own batch; bin-sharded: all ranks share ONE batch
    units = n_reads if bin_sharded else world * n_reads
    value = units * args.steps / (total_ms * 1e-3)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak" if (not bin_sharded or per_rank) else "strong",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": args.workload,
                   "mode": ("bin_sharded, one genome group and one column slice per GPU, whole filter never in one place: %d bins / %.1f Gb"
                            % (plan["n_bins"], world * w["seq_len"] / 1e9)) if per_rank else args.mode if world > 1 else "single_gpu",
                   "chunks_per_gpu_per_step": n_reads, "chunk_length": w["chunk"], "kmer_size": w["k"],
                   "bins": plan["n_bins"], "row_bytes": int(gf.bin_width * 8), "filter_bytes": plan["n_bits"] // 8,
                   "thresholds_per_pass": n_lut, "l2_fetch_granularity": rb.get_l2_fetch_granularity(local), "error_rate": ERROR_RATE, "read_mix": "%g%% reference-derived @10%% errors, %g%% iid" % (100 * args.from_ref, 100 - 100 * args.from_ref),
                   "l2": "no flush: inputs (%.0f MB reads + %.0f MB filter) exceed the 126 MB L2" % (
                       bases_np.nbytes / 1e6, plan["n_bits"] / 8e6),
                   "hit_fraction": hits_dev / n_reads, "kmer_table_bytes": gf.kmer_table_bytes(), "kmer_table_span": gf.kmer_table_span(), "kmer_table_kind": gf.kmer_table_kind(), "ibf_build_ms_gpu": build_ms, "ibf_build_ms_gpu_first_call": build_cold_ms,
                   "ibf_build_kmers_per_s": n_kmers_ref / (build_ms * 1e-3)},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    w = WORKLOADS[args.workload]
    n_reads = args.reads or w["reads"]
    if args.impl == "reference":
        run_reference(args, w, n_reads)
    else:
        run_ours(args, w, n_reads)


if __name__ == "__main__":
    main()
