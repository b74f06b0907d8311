#!/usr/bin/env python3
"""bench.py -- throughput of the IBF classify hot path on B200 (BASELINE.json metric).

One "step" = one pass of the hot path (2 x seqan::count + threshold + select/max_matches per
chunk, src/IBF/IBFClassify.cpp:138-171) over one batch of synthetic 250-base read chunks per GPU.
Default workload = BASELINE config #2: 1 M chunks vs a 100-bin IBF built (on the GPU, by the
insert kernels) from 100 synthetic 4 Mb genomes.  Multi-GPU = read-sharded with a replicated IBF
(no data-path collective, weak scaling).

The default line also carries `secondary`: at 1 GPU BASELINE config #3 (human depletion: 3.1 Gb, 31 008
bins, postings kernel) with its GPU build (= config #4) beside a CPU build sample, config #2 at k = 15 and
k = 17 (where the window table stops applying), config #1's 51-bin filter, two target-panel filters of 303 and 2 020 bins
(rows of 5 and 32 words: the group-loaded k-mer table of ibf_ctable.cu), and the reference's published
3-target + 1-depletion workload through the C++ driver; at N > 1 GPUs BASELINE config #5's per-GPU workload
(bin-sharded 30 Gb filter, one eighth per GPU) with the NCCL all-reduce(MAX) of the per-read keys inside the
timed region.  `--no-secondary` or any explicit `--workload` prints the primary line alone.

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

# Sequence lengths are what is left of a raw record after cutOutNNNs dropped its last base (quirk Q1, IBFBuild.cpp:121-125).
WORKLOADS = {
    "cfg1_5Mb_51bins": dict(lengths=[5_000_000], seed0=1, fragment=100_000, k=13, chunk=250, reads=1_000_000),
    "cfg2_100x4Mb_100bins": dict(lengths=[3_999_999] * 100, seed0=2, fragment=4_200_000, k=13, chunk=250, reads=1_000_000),
    "cfg2_k15": dict(lengths=[3_999_999] * 100, seed0=2, fragment=4_200_000, k=15, chunk=250, reads=1_000_000),
    "cfg2_k17": dict(lengths=[3_999_999] * 100, seed0=2, fragment=4_200_000, k=17, chunk=250, reads=262_144),
    "cfg2_k14": dict(lengths=[3_999_999] * 100, seed0=2, fragment=4_200_000, k=14, chunk=250, reads=1_000_000),
    "cfg2_k16": dict(lengths=[3_999_999] * 100, seed0=2, fragment=4_200_000, k=16, chunk=250, reads=262_144),
    "w4_200x2Mb_200bins": dict(lengths=[1_999_999] * 200, seed0=2, fragment=2_100_000, k=13, chunk=250, reads=1_000_000),
    # mid-size filters (a few target chromosomes): rows of 5, 16, 32 words (group-loaded k-mer table) and 64 words (postings)
    "w5_30Mb_303bins": dict(lengths=[10_050_000] * 3, seed0=400, fragment=100_000, k=13, chunk=250, reads=262_144, cpu_build_frags=64),
    "w16_100Mb_1010bins": dict(lengths=[10_050_000] * 10, seed0=400, fragment=100_000, k=13, chunk=250, reads=262_144,
                               cpu_build_frags=128),
    "w32_200Mb_2020bins": dict(lengths=[10_050_000] * 20, seed0=400, fragment=100_000, k=13, chunk=250, reads=262_144,
                               cpu_build_frags=128),
    "w16_k15": dict(lengths=[10_050_000] * 10, seed0=400, fragment=100_000, k=15, chunk=250, reads=262_144, cpu_build_frags=128),
    "w64_400Mb_4040bins": dict(lengths=[10_050_000] * 40, seed0=400, fragment=100_000, k=13, chunk=250, reads=131_072,
                               cpu_build_frags=256),
    "w128_800Mb_8080bins": dict(lengths=[10_050_000] * 80, seed0=400, fragment=100_000, k=13, chunk=250, reads=131_072,
                                cpu_build_frags=256),
    "w256_1.6Gb_16160bins": dict(lengths=[10_050_000] * 160, seed0=400, fragment=100_000, k=13, chunk=250, reads=65_536,
                                 cpu_build_frags=256),
    "cfg3_3.1Gb_31kbins": dict(lengths=[129_166_666] * 24, seed0=300, fragment=100_000, k=13, chunk=250, reads=65_536,
                               cpu_build_frags=1024),
    "w1_50x4Mb_50bins": dict(lengths=[3_999_999] * 50, seed0=2, fragment=4_200_000, k=13, chunk=250, reads=1_000_000),
    "mini_100x60kb_100bins": dict(lengths=[59_999] * 100, seed0=2, fragment=61_000, k=13, chunk=250, reads=65_536),
    "mini_k15": dict(lengths=[59_999] * 100, seed0=2, fragment=61_000, k=15, chunk=250, reads=16_384),
    "mini3_40Mb_408bins": dict(lengths=[3_350_000] * 12, seed0=300, fragment=100_000, k=13, chunk=250, reads=16_384,
                               cpu_build_frags=128),
    # BASELINE config #5 (30 Gb, ~300 k bins, 8 bin shards), one eighth per GPU: every rank generates ITS OWN 3.74 Gb genome
    # group, builds ITS OWN column slice (37 440 bins = 585 row words, 5.8 GB) and never sees the rest of the filter;
    # N ranks = an N x 37 440-bin filter over N x 3.74 Gb.
    "cfg5_3.7Gb_37kbins_per_gpu": dict(per_rank=True, seq_len=3_743_950_000, window=64_000_000, fragment=100_000, k=13, chunk=250,
                                       reads=65_536, bins_per_rank=37_440),
    "mini5_40Mb_512bins_per_gpu": dict(per_rank=True, seq_len=51_150_000, window=8_000_000, fragment=100_000, k=13, chunk=250,
                                       reads=16_384, bins_per_rank=512),
}
DEFAULT_WORKLOAD = "cfg2_100x4Mb_100bins"
SECONDARY_1GPU = ["cfg3_3.1Gb_31kbins", "cfg2_k15", "cfg2_k17", "cfg1_5Mb_51bins", "w5_30Mb_303bins", "w32_200Mb_2020bins",
                  "readme_3targets_1deplete", "live_3targets_1deplete"]
README_READS = 100_000      # the reference's one published workload (README.md:233-262), tools/readme_bench.py
SECONDARY_NGPU = ["cfg5_3.7Gb_37kbins_per_gpu"]
ERROR_RATE = 0.1
SIGNIFICANCE = 0.95
ORACLE_SAMPLE = 3000


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--secondary", default=None, help="comma-separated workloads to append as `secondary` (default: see module doc)")
    ap.add_argument("--no-secondary", action="store_true")
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


def make_reference(w, rank=0):
    from readbouncer_b200 import synth
    if w.get("per_rank"):
        return synth.HashReference([w["seq_len"]], 5000 + rank)
    return synth.HashReference(w["lengths"], w["seed0"])


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


def read_mix(args):
    return "%g%% reference-derived @10%% errors, %g%% iid" % (100 * args.from_ref, 100 - 100 * args.from_ref)


# ------------------------------------------------------------------------------------------------------
def run_reference(args, name, w, n_reads):
    """Reference arm: the CPU restatement of the reference's classify path (oracle port; the
    reference binary cannot be compiled here), all host threads, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    from readbouncer_b200 import synth     # numpy generators and the scalar host helpers only; no kernel on this arm
    cores = os.cpu_count() or 1
    if w.get("per_rank"):
        raise SystemExit("the reference arm runs the replicated-filter workloads (the CPU holds one filter)")
    ref = make_reference(w)
    plan = ref.plan(w["fragment"], w["k"])
    bases_ref = ref.host()
    t0 = time.time()
    of = oracle.OracleIBF.create(plan["n_bins"], 3, w["k"], plan["n_bits"])
    of.insert_batch(bases_ref, plan["frag_begin"], plan["frag_end"], plan["frag_bin"], n_threads=cores)
    build_s = time.time() - t0
    lut = oracle.threshold_lut(ERROR_RATE, w["k"], SIGNIFICANCE)
    bases, off, _ = synth.sample_reads(bases_ref, min(n_reads, 400_000), w["chunk"], seed=1234, frac_from_ref=args.from_ref)
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
        "config": {"workload": name, "chunks_per_step_sample": sample, "chunk_length": w["chunk"], "kmer_size": w["k"],
                   "bins": plan["n_bins"], "filter_bytes": plan["n_bits"] // 8, "error_rate": ERROR_RATE,
                   "read_mix": read_mix(args), "ibf_build_s_cpu": build_s,
                   "ibf_build_kmers_per_s_cpu": float(np.maximum(plan["frag_end"] - plan["frag_begin"], w["k"] - 1).sum()
                                                      - (w["k"] - 1) * len(plan["frag_bin"])) / build_s},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample_desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------------
class Env:
    """Process-wide state of the repo arm: rank, device, stream, the bound library."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import readbouncer_b200 as rb
        self.torch, self.dist, self.rb = torch, dist, rb
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available() or rb.device_count() < 1:
            raise RuntimeError("bench.py needs a CUDA device: the hot path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        rb.set_count_kernel(args.kernel)
        if args.l2_gran:
            rb.set_l2_fetch_granularity(args.l2_gran, device=self.local)
        self.stream = torch.cuda.current_stream()
        self.lib = rb.lib()
        self.cores = os.cpu_count() or 1

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def events(self, n=2):
        return [self.torch.cuda.Event(enable_timing=True) for _ in range(n)]


def h2d_ceiling(env, nbytes=256 << 20, reps=8):
    """Platform ceiling of the host feed: bare pinned cudaMemcpyAsync H2D, all ranks at the same time (GB/s, whole job)."""
    torch = env.torch
    h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h.numpy()[:] = 1
    d = torch.empty(nbytes, dtype=torch.uint8, device=env.dev)
    d.copy_(h, non_blocking=True)
    env.barrier()
    e0, e1 = env.events()
    e0.record(env.stream)
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e1.record(env.stream)
    torch.cuda.synchronize()
    ms, = env.max_over_ranks(e0.elapsed_time(e1))
    del h, d
    return env.world * nbytes * reps / (ms * 1e-3) / 1e9


def run_workload(env, args, name, n_reads, steps, warmup, role):
    """One workload on this process's GPU: build, classify (device-resident and end to end), parity, roofline.
    Returns the JSON-line dict on rank 0 (None elsewhere)."""
    torch, dist, rb, lib = env.torch, env.dist, env.rb, env.lib
    from readbouncer_b200 import synth
    w = WORKLOADS[name]
    rank, world, local, dev, stream = env.rank, env.world, env.local, env.dev, env.stream
    per_rank = bool(w.get("per_rank"))
    bin_sharded = (args.mode == "bin_sharded" and world > 1 and role == "primary") or per_rank
    primary = role == "primary"
    k, chunk = w["k"], w["chunk"]

    # ---- the reference is generated in HBM (rb_synth_bases_dev), the IBF is built on the GPU ---------------------------
    ref = make_reference(w, rank)
    if per_rank:
        per_bins = w["seq_len"] // w["fragment"] + 1                  # IBFBuild.cpp:90
        assert per_bins == w["bins_per_rank"] and per_bins % 64 == 0
        plan = ref.plan(w["fragment"], k, bin0=rank * per_bins, n_bins=world * per_bins)
        assert plan["bin_ids_consumed"] == per_bins
        gf_full = rb.IBF.create_shard(plan["n_bins"], 3, k, plan["n_bits"], rank, world, device=local)
        assert gf_full.bin_begin == rank * per_bins and gf_full.n_bins_local >= per_bins
        # the first `window` bases of every rank's genome group are what the reads are sampled from (any rank regenerates
        # them: the bases are a pure function of (seed, position)), so all ranks classify the same batch
        read_src = synth.HashReference([w["window"]] * world, 5000)
    else:
        plan = ref.plan(w["fragment"], k)
        gf_full = rb.IBF.create(plan["n_bins"], 3, k, plan["n_bits"], device=local)
        read_src = ref
    d_ref = ref.to_device(dev, stream)
    d_fb = torch.from_numpy(plan["frag_begin"].astype(np.int64)).to(dev)
    d_fe = torch.from_numpy(plan["frag_end"].astype(np.int64)).to(dev)
    d_fbin = torch.from_numpy(plan["frag_bin"].astype(np.int64)).to(dev)
    max_frag = int((plan["frag_end"] - plan["frag_begin"]).max())
    ev0, ev1 = env.events()
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
    n_kmers_ref = int(np.maximum(plan["frag_end"] - plan["frag_begin"], k - 1).sum() - (k - 1) * len(plan["frag_bin"]))
    del d_ref, d_fb, d_fe, d_fbin
    torch.cuda.empty_cache()
    gf = gf_full
    if bin_sharded and not per_rank:
        words = gf_full.download()
        gf = rb.IBF.from_words(words, plan["n_bins"], 3, k, plan["n_bits"], device=local, shard=rank, n_shards=world)
        del words

    # ---- reads: host (pinned) and device copies --------------------------------------------------------
    # read-sharded: every rank classifies its own batch; bin-sharded: all ranks see the same batch
    seed = 1234 if bin_sharded else 1234 + rank
    bases_np, off_np, from_ref = synth.sample_reads(read_src, n_reads, chunk, seed=seed, frac_from_ref=args.from_ref)
    h_bases = torch.empty(bases_np.size, dtype=torch.uint8, pin_memory=True)
    h_bases.numpy()[:] = bases_np
    h_off = torch.empty(off_np.size, dtype=torch.int64, pin_memory=True)
    h_off.numpy()[:] = off_np.astype(np.int64)
    luts_np = np.stack([rb.threshold_lut(ERROR_RATE, k, SIGNIFICANCE),
                        rb.threshold_lut(ERROR_RATE - 0.02, k, SIGNIFICANCE)])
    n_lut = 2                                     # both thresholds of check_unblock in one pass
    d_bases = h_bases.to(dev, non_blocking=True)
    d_off = h_off.to(dev, non_blocking=True)
    d_lut = torch.from_numpy(luts_np.view(np.int16)).to(dev)
    d_keys = torch.zeros(n_lut * n_reads, dtype=torch.int64, device=dev)
    d_max = torch.zeros(n_lut * n_reads, dtype=torch.int16, device=dev)
    d_hit = torch.zeros(n_lut * n_reads, dtype=torch.uint8, device=dev)
    d_amax = torch.zeros(n_lut * n_reads, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()

    def count(keys=None, handle=None):
        (handle or gf).count_batch_dev(d_bases, d_off, n_reads, d_lut, n_lut, d_keys if keys is None else keys,
                                       max_read_len=chunk, stream=stream)

    def decode():
        rb.capi._check(lib.rb_keys_decode_dev(rb.capi._dev_ptr(d_keys), n_lut * n_reads, rb.capi._dev_ptr(d_max),
                                              rb.capi._dev_ptr(d_hit), rb.capi._dev_ptr(d_amax), local,
                                              rb.capi._stream_ptr(stream)))

    def combine():
        if bin_sharded and world > 1:
            dist.all_reduce(d_keys, op=dist.ReduceOp.MAX)       # keys < 2^49: int64 MAX == uint64 MAX

    # ---- one-off costs a caller sees: the k-mer table build and the first call --------------------------------------
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    table_build_ms = None
    if args.kernel in (0, 3, 4, 5):
        try:
            gf.enable_kmer_table(0, stream=stream)             # what the first large count call would do by itself
            torch.cuda.synchronize()
            table_build_ms = 1000 * (time.perf_counter() - t0)
        except rb.RBError:
            table_build_ms = None                              # no table applies (k > 16 ...): hashed probes
    count()
    torch.cuda.synchronize()
    cold_first_call_ms = 1000 * (time.perf_counter() - t0)

    # ---- device-resident timing (value) ------------------------------------------------------------------
    for _ in range(max(warmup, 3)):
        count(); combine(); decode()
    env.barrier()
    # clocks / throttle reasons: nvidia-smi every 100 ms from here to the end of the timed regions (device-resident and
    # end-to-end).  A timed region of a few tens of ms would see at most one sample, so the GPU is kept under the same
    # load (untimed steps) until the first samples have arrived.
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        t_wait = time.time()
        while sampler.n_samples() < 3 and time.time() - t_wait < 3.0:
            count()
            torch.cuda.synchronize()              # no collective here: the other ranks wait at the barrier below
    env.barrier()
    count(); combine(); decode()
    torch.cuda.synchronize()

    # ---- parity of this very batch (before anything is timed) ----------------------------------------------------------
    parity = {}
    pick = np.sort(np.random.default_rng(9).choice(n_reads, min(ORACLE_SAMPLE, n_reads), replace=False))
    sb = bases_np.reshape(n_reads, chunk)[pick].reshape(-1)
    so = np.arange(len(pick) + 1, dtype=np.uint64) * np.uint64(chunk)
    g_keys = d_keys.cpu().numpy().view(np.uint64).reshape(n_lut, n_reads)
    g_max, g_hit, g_amax = rb.keys_decode(g_keys)
    assert np.array_equal(g_max, d_max.cpu().numpy().view(np.uint16).reshape(n_lut, n_reads))
    assert np.array_equal(g_hit, d_hit.cpu().numpy().reshape(n_lut, n_reads))
    assert np.array_equal(g_amax, d_amax.cpu().numpy().view(np.uint32).reshape(n_lut, n_reads))
    oracle_s = None
    if not args.no_cpu_baseline or not primary:
        import oracle
        if bin_sharded:
            # every rank checks ITS shard against the oracle holding the same column slice (local bin ids), then the oracle's
            # keys are combined the same way and must equal the all-reduced GPU keys
            d_own = torch.zeros_like(d_keys)
            count(keys=d_own)
            torch.cuda.synchronize()
            own = d_own.cpu().numpy().view(np.uint64).reshape(n_lut, n_reads)
            del d_own
            nb_local = gf.n_bins_local
            of = oracle.OracleIBF.create(nb_local, 3, k, gf.n_blocks * 64 * gf.col_words)
            assert of.bin_width == gf.col_words and of.n_blocks == gf.n_blocks
            nw = gf.n_blocks * gf.col_words                  # (an unsharded handle also holds the words behind the last whole row)
            of.words()[:nw] = gf.download()[:nw]
            exp_keys = np.zeros((n_lut, len(pick)), np.uint64)
            t0 = time.time()
            for t in range(n_lut):
                exp = of.count_batch(sb, so, luts_np[t], dense=False, n_threads=env.cores)
                glob = (exp["argmax_bin"].astype(np.uint64) + np.uint64(gf.bin_begin)) & np.uint64(0xFFFFFFFF)
                exp_keys[t] = np.where(exp["hit"] > 0, (np.uint64(1) << np.uint64(48)) | (exp["max_count"].astype(np.uint64) << np.uint64(32))
                                       | (~glob & np.uint64(0xFFFFFFFF)), np.uint64(0))
            oracle_s = (time.time() - t0) / n_lut
            assert np.array_equal(own[:, pick], exp_keys), "this shard's GPU keys differ from the oracle's on the sampled chunks"
            comb = torch.from_numpy(exp_keys.view(np.int64)).to(dev)
            if world > 1:
                dist.all_reduce(comb, op=dist.ReduceOp.MAX)
            assert np.array_equal(comb.cpu().numpy().view(np.uint64), g_keys[:, pick]), "combined keys differ from the combined oracle keys"
            parity["oracle"] = "max_count, hit, argmax_bin of %d sampled chunks x %d thresholds: shard-local GPU keys == oracle on the same " \
                               "column slice, and all-reduced GPU keys == MAX over ranks of the oracle keys" % (len(pick), n_lut)
            del of
        else:
            of = oracle.OracleIBF.create(plan["n_bins"], 3, k, plan["n_bits"])
            words = gf.download()
            of.words()[:plan["n_bits"] // 64] = words
            t0 = time.time()
            for t in range(n_lut):
                exp = of.count_batch(sb, so, luts_np[t], dense=False, n_threads=env.cores)
                assert np.array_equal(g_max[t][pick], exp["max_count"]), "max_count differs from the oracle"
                assert np.array_equal(g_hit[t][pick], exp["hit"]), "hit differs from the oracle"
                assert np.array_equal(g_amax[t][pick], exp["argmax_bin"]), "argmax_bin differs from the oracle"
            oracle_s = (time.time() - t0) / n_lut
            parity["oracle"] = "max_count, hit, argmax_bin of %d sampled chunks x %d thresholds == CPU oracle on the downloaded filter" % (
                len(pick), n_lut)
            if w.get("cpu_build_frags"):
                # config #4: the oracle builds the first M fragments on the host cores; those bin columns of the GPU-built
                # matrix must equal the oracle's (M is a multiple of 64: whole row words)
                m = min(int(w["cpu_build_frags"]), len(plan["frag_bin"])) // 64 * 64
                end = int(plan["frag_end"][m - 1])
                hb = ref.host(end)
                ob = oracle.OracleIBF.create(plan["n_bins"], 3, k, plan["n_bits"])
                t0 = time.time()
                ob.insert_batch(hb, plan["frag_begin"][:m], plan["frag_end"][:m], plan["frag_bin"][:m], n_threads=env.cores)
                cpu_build_s = time.time() - t0
                bw = gf.bin_width
                gcols = words.reshape(-1, bw)[:, :m // 64]
                ocols = ob.words()[:plan["n_bits"] // 64].reshape(-1, bw)[:, :m // 64]
                assert np.array_equal(gcols, ocols), "GPU-built bin columns differ from the oracle-built ones"
                km = int(np.maximum(plan["frag_end"][:m] - plan["frag_begin"][:m], k - 1).sum() - (k - 1) * m)
                parity["build"] = "bin columns 0..%d of the GPU-built matrix (all %d rows) == oracle-built" % (m - 1, gf.n_blocks)
                parity["cpu_build"] = {"kmers_per_s": km / cpu_build_s, "cores": env.cores, "kind": "port",
                                       "sample": "first %d of %d fragments (%.0f Mb), %.1f s" % (m, len(plan["frag_bin"]), end / 1e6, cpu_build_s)}
                del ob, hb
            del of, words
    if per_rank:
        # nobody holds the whole filter: the combined answer of a reference-derived chunk must name a bin of the sampling
        # windows (the first window/fragment + 1 bins of a rank's range; 1.7 % of the bins by chance)
        sel = (g_hit[0] > 0) & from_ref
        in_window = (g_amax[0][sel].astype(np.int64) % w["bins_per_rank"]) <= w["window"] // w["fragment"] + 1
        assert sel.sum() > 0.9 * from_ref.sum() and in_window.mean() > 0.99, (sel.sum(), from_ref.sum(), in_window.mean())
    elif bin_sharded:
        # the all-reduced shard keys must equal the keys of the whole (replicated) filter
        d_full = torch.zeros_like(d_keys)
        count(keys=d_full, handle=gf_full)
        torch.cuda.synchronize()
        assert torch.equal(d_full, d_keys), "bin-sharded combine differs from the whole-filter result"
        parity["combine"] = "all-reduced shard keys == keys of the whole filter, all %d chunks" % n_reads
        del d_full
        gf_full.close()
        torch.cuda.empty_cache()

    # ---- what the launch has to fetch (table geometry; the ncu check of the same number lives in profiles/) -------------
    table_bytes, table_reqs, io_bytes = gf.count_traffic_dev(d_bases, d_off, n_reads, n_lut, stream=stream)

    env.barrier()
    launches0 = rb.kernel_launches()
    k_ev = [env.events() for _ in range(steps)]
    c_ev = [env.events() for _ in range(steps)]
    t_ev0, t_ev1 = env.events()
    t_ev0.record(stream)
    for i in range(steps):
        k_ev[i][0].record(stream)
        count()
        k_ev[i][1].record(stream)
        if bin_sharded and world > 1:
            c_ev[i][0].record(stream)
            combine()
            c_ev[i][1].record(stream)
        decode()
    t_ev1.record(stream)
    env.barrier()
    launches = rb.kernel_launches() - launches0
    total_ms = t_ev0.elapsed_time(t_ev1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in k_ev]))
    coll_ms = float(np.mean([a.elapsed_time(b) for a, b in c_ev])) if bin_sharded and world > 1 else 0.0
    total_ms, kernel_ms, coll_ms = env.max_over_ranks(total_ms, kernel_ms, coll_ms)
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
        for _ in range(max(warmup, 5)):       # the library times both ways in (2 calls each) and keeps the faster
            e2e_step()
        env.barrier()
        xfer0 = rb.transfer_bytes()
        t0 = time.perf_counter()
        for _ in range(steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        xfer1 = rb.transfer_bytes()
        e2e_s, = env.max_over_ranks(e2e_s)
        # the host-buffer call must return what the device call returned: every array, every chunk, both thresholds
        assert np.array_equal(res_max.reshape(n_lut, n_reads), g_max), "host-API max_count differs from the device API's"
        assert np.array_equal(res_hit.reshape(n_lut, n_reads), g_hit), "host-API hit differs from the device API's"
        assert np.array_equal(res_am.reshape(n_lut, n_reads), g_amax), "host-API argmax_bin differs from the device API's"
        assert not res_flag.any()
        parity["host_call"] = "rb_ibf_count_batch max_count / hit / argmax_bin == device API, all %d chunks x %d thresholds" % (n_reads, n_lut)
        # bytes that crossed PCIe, counted by the library at its copy calls: the host threads turn the ASCII bases into
        # bit planes before the transfer, so this is less than the size of the host input
        h2d_step = (xfer1[0] - xfer0[0]) // steps
        e2e = {"value": world * n_reads * steps / e2e_s, "unit": UNIT,
               "h2d_bytes_per_step": h2d_step,
               "d2h_bytes_per_step": (xfer1[1] - xfer0[1]) // steps,
               "host_input_bytes_per_step": int(hb.nbytes + ho.nbytes + luts_np.nbytes),
               "host_result_bytes_per_step": int(res_max.nbytes + res_hit.nbytes + res_am.nbytes + res_flag.nbytes),
               "ms_per_step": 1000 * e2e_s / steps, "api": "rb_ibf_count_batch (host buffers, pinned)",
               "host_pack": dict(rb.host_pack_info(), enabled=os.environ.get("RB_HOST_PACK", "1") != "0"),
               "transfer_policy": gf.transfer_policy()}
        if primary:
            # the platform's ceiling for the host feed, measured in this run: bare pinned H2D copies on all ranks at once.
            # ascii_bound = chunks/s if the ASCII input crossed at that rate with nothing else in the way.
            ceil_gbs = h2d_ceiling(env)
            e2e["h2d_ceiling_gbs"] = ceil_gbs
            # ... and how fast this rank's packer threads can stream-read the input at all (all ranks at the same time)
            env.barrier()
            g, = env.max_over_ranks(-rb.capi.host_read_gbs(hb))
            e2e["host_read_gbs_per_rank_min"] = -g
            e2e["host_read_note"] = "packer threads stream-reading the step's input, no work, no stores; the packed path reads the input " \
                                    "once and writes + DMA-reads 3/8 of it again"
            e2e["h2d_achieved_gbs"] = world * h2d_step / (e2e_s / steps) / 1e9
            e2e["host_input_gbs"] = world * e2e["host_input_bytes_per_step"] / (e2e_s / steps) / 1e9
            e2e["ascii_bound_chunks_per_s"] = ceil_gbs * 1e9 / (e2e["host_input_bytes_per_step"] / n_reads)
            e2e["frac_of_ascii_bound"] = e2e["value"] / e2e["ascii_bound_chunks_per_s"]

    clocks = sampler.stop() if sampler else None
    if rank != 0:
        gf.close()
        return None

    # ---- roofline of the dominant kernel (count kernel) ----------------------------------------------------
    lookups, bytes_per_chunk = algorithmic_bytes_per_chunk(w, gf.col_words)
    peak, peak_src = measured_peaks()
    span = gf.kmer_table_span()
    kind = gf.kmer_table_kind()
    group_ok = span >= 2 and chunk - k + 1 <= 127 * span and chunk <= 545   # ibf_wtable.cu: wgroup_applicable
    tb_ = gf.kmer_table_bytes()
    # the library picks the lookup kernel of a postings table by its geometry (ibf_postings.cu): slots of one or two lines and
    # lists averaging up to 16 units are walked by groups of lanes, the rest by whole warps
    slots_direct = kind == 3 and (tb_ // 4 ** k) // 128 * 128 <= 256 and os.environ.get("RB_SLOTS_SUB", "1") != "0"
    lists_short = kind == 2 and (tb_ - 4 * (4 ** k + 1)) / 16 / 4 ** k <= 16.0 and os.environ.get("RB_POSTINGS_SUB", "") != "0"
    kernel_name = ("count_slots_sub_kernel" if slots_direct else "count_slots_kernel" if kind == 3
                   else "count_postings_sub_kernel" if lists_short else "count_postings_kernel" if kind == 2 else "count_ctable_kernel" if kind == 4
                   else "count_wgroup_kernel" if group_ok else "count_wtable_kernel" if span >= 2
                   else "count_table_kernel" if span == 1 else
                   "count_stream_kernel" if (gf.col_words > 4 or args.kernel == 2) else "count_tile_kernel")
    streamed = kernel_name == "count_stream_kernel"
    # DRAM bytes of one launch.  Table kernels: the table data the launch must fetch at 128-byte line granularity (HBM delivers
    # whole lines; tables of tens of GB are not found in L2) + bases/offsets in + keys out.  The streaming kernel reads whole row
    # segments: its traffic is the reference's algorithmic bytes.
    traffic = (n_reads * bytes_per_chunk if streamed else table_bytes) + io_bytes
    achieved = traffic / (kernel_ms * 1e-3) / 1e9
    alg = n_reads * bytes_per_chunk / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic,
                "traffic_source": "table geometry (rb_ibf_count_traffic_dev): %d table accesses, %d table bytes at 128-byte line "
                                  "granularity + %d bytes of reads / offsets / keys; ncu dram__bytes cross-check: profiles/traffic.json"
                                  % (table_reqs, table_bytes, io_bytes),
                "kernel": kernel_name, "kernel_ms": kernel_ms, "peak_source": peak_src,
                "algorithmic_bytes_per_chunk": bytes_per_chunk, "algorithmic_achieved": alg, "algorithmic_frac": alg / peak,
                "algorithmic_note": "SURVEY 8(d): lookups x h x row bytes, the reference's row probes; the k-mer table / postings lists "
                                    "replace those probes, so this is not what the kernel moves (frac is)" if not streamed else
                                    "the streaming kernel reads exactly these row segments",
                "kmer_lookups_per_s": (n_reads if bin_sharded else world * n_reads) * lookups / (total_ms * 1e-3 / steps)}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):          # measured under ncu --set full for this workload and kernel: per-chunk figure for comparison
        ent = json.load(open(tp)).get(name, {})
        if ent.get("kernel") == kernel_name:
            roofline["traffic_ncu_per_chunk"] = ent.get("dram_bytes_per_launch") / ent.get("chunks_per_launch")
            roofline["traffic_per_chunk"] = traffic / n_reads
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

    npos = chunk - k + 1
    if kind in (2, 3):
        pass                # postings lists are streamed: the HBM-bandwidth roofline above applies
    elif kind == 4:     # one k-mer per entry of 64 / 128 / 256 bytes, 16 bytes per lane, adjacent lanes
        entry_bytes = gf.kmer_table_bytes() // 4 ** k
        lanes = entry_bytes // 16
        ppg = 64 * lanes
        g_ms = time_gather(lambda: rb.microbench_gather_coop(gf.device_kmer_table_ptr(), 4 ** k, entry_bytes, 16, ppg, blocks, sink,
                                                              stream=stream))
        roofline["requests"] = {"what": "one %d-byte table entry (1 k-mer position, both strands, %d lanes) per request" % (entry_bytes, lanes),
                                "per_chunk": npos, "peak_per_s": blocks * 256 // lanes * ppg / (g_ms * 1e-3),
                                "achieved_per_s": n_reads * npos / (kernel_ms * 1e-3),
                                "how": "rb_microbench_gather_coop over the k-mer table itself (%d entries)" % 4 ** k}
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
    if world == 1 and not args.no_cpu_baseline and primary:
        import oracle
        cores = env.cores
        of = oracle.OracleIBF.create(plan["n_bins"], 3, k, plan["n_bits"])
        of.words()[:plan["n_bits"] // 64] = gf.download()
        lut0 = luts_np[0]
        cal = min(2000, n_reads)
        t0 = time.time()
        exp = of.count_batch(bases_np[:cal * chunk], off_np[:cal + 1], lut0, dense=False, n_threads=cores)
        rate = cal / max(time.time() - t0, 1e-6)
        sample = int(max(cal, min(n_reads, rate * 15.0)))
        sb2, so2 = bases_np[:sample * chunk], off_np[:sample + 1]
        # 10-20 s of CPU work: the sample is classified again until 12 s have passed (a batch of 1 M chunks takes ~4 s)
        passes, dt = 0, 0.0
        while passes == 0 or (dt < 12.0 and passes < 64):
            t0 = time.time()
            exp = of.count_batch(sb2, so2, lut0, dense=False, n_threads=cores)
            dt += time.time() - t0
            passes += 1
        assert np.array_equal(g_max[0][:sample], exp["max_count"])
        assert np.array_equal(g_hit[0][:sample], exp["hit"])
        assert np.array_equal(g_amax[0][:sample], exp["argmax_bin"])
        parity["cpu_baseline_sample"] = "max_count, hit, argmax_bin of the first %d chunks == CPU oracle" % sample
        # one host thread, and the reference's call shape (one read per call: IBFClassify.cpp:138-171 is entered per read)
        n1 = int(max(64, min(sample, rate / cores * 3.0)))
        t0 = time.time()
        of.count_batch(sb2[:n1 * chunk], so2[:n1 + 1], lut0, dense=False, n_threads=1)
        dt1 = time.time() - t0
        n2 = min(n1, 2000)
        t0 = time.time()
        for i in range(n2):
            of.count_batch(sb2[i * chunk:(i + 1) * chunk], so2[:2], lut0, dense=False, n_threads=1)
        dt2 = time.time() - t0
        cpu_baseline = {"value": sample * passes / dt, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "first %d chunks of the step's batch x %d passes, %d threads, %.1f s; results equal the GPU's" % (
                            sample, passes, cores, dt),
                        "single_thread": {"value": n1 / dt1, "chunks": n1},
                        "one_read_per_call_single_thread": {"value": n2 / dt2, "chunks": n2}}
        del of
    elif oracle_s:
        cpu_baseline = {"value": len(pick) / oracle_s, "unit": UNIT, "cores": env.cores, "kind": "port",
                        "sample": "the %d sampled chunks of the parity check, %d threads, %.1f s per threshold" % (len(pick), env.cores, oracle_s)}

    # read-sharded: every rank classifies its own batch; bin-sharded: all ranks share ONE batch
    units = n_reads if bin_sharded else world * n_reads
    value = units * steps / (total_ms * 1e-3)
    table_b = gf.kmer_table_bytes()
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": max(warmup, 3),
        "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "weak" if (not bin_sharded or per_rank) else "strong",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": name,
                   "mode": ("bin_sharded, one genome group and one column slice per GPU, whole filter never in one place: %d bins / %.1f Gb"
                            % (plan["n_bins"], world * w["seq_len"] / 1e9)) if per_rank else
                           ("bin_sharded" if bin_sharded else "read_sharded") if world > 1 else "single_gpu",
                   "chunks_per_gpu_per_step": n_reads, "chunk_length": chunk, "kmer_size": k,
                   "bins": plan["n_bins"], "row_bytes": int(gf.bin_width * 8), "filter_bytes": plan["n_bits"] // 8,
                   "thresholds_per_pass": n_lut, "l2_fetch_granularity": rb.get_l2_fetch_granularity(local), "error_rate": ERROR_RATE,
                   "read_mix": read_mix(args),
                   "l2": "no flush: inputs (%.0f MB reads + %.0f MB filter + %.0f MB k-mer table) exceed the 126 MB L2" % (
                       bases_np.nbytes / 1e6, plan["n_bits"] / 8e6, table_b / 1e6),
                   "hit_fraction": hits_dev / n_reads, "kmer_table_bytes": table_b,
                   "kmer_table_over_filter_bytes": table_b / (gf.n_local_words * 8),
                   "kmer_table_span": span, "kmer_table_kind": kind,
                   "kmer_table_build_ms": table_build_ms, "cold_first_call_ms": cold_first_call_ms,
                   "ibf_build_ms_gpu": build_ms, "ibf_build_ms_gpu_first_call": build_cold_ms,
                   "ibf_build_kmers_per_s": n_kmers_ref / (build_ms * 1e-3),
                   "reference_generated": "in HBM by rb_synth_bases_dev (hash of seed and position); host regenerates the sampled windows"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
        "parity": parity,
    }
    if bin_sharded and world > 1:
        out["collective"] = {"op": "NCCL all_reduce(MAX) of the packed per-read keys (int64, %d bytes per rank and step)" % (8 * n_lut * n_reads),
                             "ms_per_step": coll_ms, "inside_timed_region": True, "count_kernel_ms": kernel_ms}
    gf.close()
    return out


def release(env):
    import gc
    gc.collect()
    env.torch.cuda.synchronize()
    env.torch.cuda.empty_cache()


def main():
    args = parse_args()
    explicit = args.workload is not None
    name = args.workload or DEFAULT_WORKLOAD
    w = WORKLOADS[name]
    n_reads = args.reads or w["reads"]
    if args.impl == "reference":
        run_reference(args, name, w, n_reads)
        return
    env = Env(args)
    out = run_workload(env, args, name, n_reads, args.steps, args.warmup, "primary")
    release(env)
    sec_names = []
    secondary = []
    if args.secondary is not None:
        sec_names = [s for s in args.secondary.split(",") if s]
    elif not explicit and not args.no_secondary and args.mode == "read_sharded" and args.kernel == 0:
        sec_names = SECONDARY_1GPU if env.world == 1 else SECONDARY_NGPU
    for s in sec_names:
        if s.startswith("live_"):
            # usage="target" without a sequencer: micro-batches of 250-base chunks through rblive::LiveClassifier (check_unblock
            # against 3 target + 1 depletion filter, both thresholds of the retry in one pass per filter); latency per batch
            if env.rank == 0:
                t0 = time.time()
                exe = os.path.join(ROOT, "readbouncer_b200", "bin", "rb_live_bench")
                try:
                    p = subprocess.run([exe, "3", "1", "4000000", "100", "64", "512", "4096"], capture_output=True, text=True, timeout=300)
                    rows = [json.loads(ln) for ln in p.stdout.splitlines() if ln.startswith("{")]
                    r = {"workload": s, "what": "time from handing a micro-batch to rblive::LiveClassifier::classify_batch to having its "
                                              "decisions (adaptive_sampling.hpp:214-356 batched; 4 filters, k = 13, fragment_size 100 000)",
                         "micro_batches": rows}
                    if p.returncode != 0 or not rows:
                        r["error"] = p.stderr[-500:]
                except Exception as e:
                    r = {"workload": s, "error": "%s: %s" % (type(e).__name__, e)}
                r["wall_s"] = time.time() - t0
                secondary.append(r)
            continue
        if s.startswith("readme_"):
            # the reference's published use case: 3 target + 1 depletion filter, FASTA in -> files out through the C++ driver
            if env.rank == 0:
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import readme_bench
                n = int(s.split(":")[1]) if ":" in s else README_READS
                t0 = time.time()
                try:
                    r = readme_bench.run(n)
                except Exception as e:                      # a failed secondary must not cost the primary line
                    r = {"workload": s, "error": "%s: %s" % (type(e).__name__, e)}
                r["wall_s"] = time.time() - t0
                secondary.append(r)
            continue
        sw = WORKLOADS[s]
        t0 = time.time()
        try:
            r = run_workload(env, args, s, sw["reads"], max(3, min(args.steps, 10)), 3, "secondary")
        except Exception as e:
            if env.world > 1:
                raise                                       # the other ranks would wait in a collective
            r = {"config": {"workload": s}, "error": "%s: %s" % (type(e).__name__, e)}
        release(env)
        if r is not None:
            r["wall_s"] = time.time() - t0
            for key in ("metric", "unit", "higher_is_better", "vs_baseline", "dtype", "data"):
                r.pop(key, None)
            secondary.append(r)
    if env.rank == 0:
        if sec_names:
            out["secondary"] = secondary
        print(json.dumps(out))
    if env.world > 1:
        env.dist.destroy_process_group()


if __name__ == "__main__":
    main()
