"""Deterministic synthetic references and read chunks plus the host-side
build plan (N-cut, bins per sequence, fragment schedule) that feeds rb_ibf_insert_batch.

Everything here is host-side numpy; the scalar rules (cutOutNNNs, fragment loop, sizing) are the
C-ABI's own host functions, not re-implemented in Python.
"""
import numpy as np

from . import capi

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
_COMP[:] = ord("N")
for _a, _b in zip(b"ACGTacgt", b"TGCAtgca"):
    _COMP[_a] = _b


def random_bases(n, seed):
    """n iid-uniform ACGT bases as ASCII uint8."""
    rng = np.random.default_rng(seed)
    return ACGT[rng.integers(0, 4, size=int(n), dtype=np.uint8)]


_M64 = (1 << 64) - 1


def _mix64(z):
    """splitmix64 finaliser on a uint64 array (same as mix64 in csrc/ibf_microbench.cu)."""
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def hash_bases(start, n, seed):
    """Bases [start, start + n) of the synthetic stream `seed`: the host twin of rb_synth_bases_dev (a pure function of
    (seed, position), so multi-Gb references are generated on the GPU and only the windows reads are sampled from are
    regenerated here)."""
    start, n = int(start), int(n)
    b0, b1 = start >> 5, (start + n + 31) >> 5
    base = np.uint64((int(seed) * 0xD1342543DE82EF95) & _M64)
    with np.errstate(over="ignore"):
        h = _mix64(np.arange(b0, b1, dtype=np.uint64) + base)
    codes = ((h[:, None] >> (np.uint64(2) * np.arange(32, dtype=np.uint64))[None, :]) & np.uint64(3)).astype(np.uint8)
    return ACGT[codes].reshape(-1)[start - 32 * b0: start - 32 * b0 + n]


class HashReference:
    """A synthetic reference of several sequences, sequence i = hash_bases(0, lengths[i], seed0 + i) -- what is left of a
    raw record of lengths[i] + 1 bases after cutOutNNNs dropped its last base (IBFBuild.cpp:121-125, quirk Q1).  Never
    materialised on the host at full size: `to_device` generates it in HBM, `windows` regenerates what reads are sampled
    from, `plan` is IBF::create_filter's host logic on the lengths alone."""

    def __init__(self, lengths, seed0):
        self.lengths = [int(x) for x in lengths]
        self.seeds = [int(seed0) + i for i in range(len(self.lengths))]
        self.offsets = np.concatenate([[0], np.cumsum(self.lengths)]).astype(np.int64)

    def __len__(self):
        return int(self.offsets[-1])

    def host(self, end=None):
        """The concatenated bases on the host, or their first `end`."""
        end = len(self) if end is None else int(end)
        parts = []
        for o, n, s in zip(self.offsets[:-1], self.lengths, self.seeds):
            if o >= end:
                break
            parts.append(hash_bases(0, min(n, end - int(o)), s))
        return np.concatenate(parts) if parts else np.zeros(0, np.uint8)

    def to_device(self, device, stream=None):
        import torch
        d = torch.empty(len(self) + 32, dtype=torch.uint8, device=device)     # kernels fetch aligned 16-byte blocks
        for o, n, s in zip(self.offsets[:-1], self.lengths, self.seeds):
            capi.synth_bases_dev(d.data_ptr() + int(o), n, s, 0, stream)
        return d

    def windows(self, pos, win):
        """(len(pos), win) bases starting at the global positions `pos` of the concatenated reference; a window that
        would run over the end of its sequence is moved back to end there."""
        pos = np.asarray(pos, dtype=np.int64)
        win = int(win)
        seq = np.searchsorted(self.offsets, pos, side="right") - 1
        lens = np.asarray(self.lengths, dtype=np.int64)[seq]
        assert (lens >= win).all()
        local0 = np.minimum(pos - self.offsets[seq], lens - win)
        nblk = win // 32 + 2
        base = np.array([(s * 0xD1342543DE82EF95) & _M64 for s in self.seeds], dtype=np.uint64)[seq]
        with np.errstate(over="ignore"):
            h = _mix64((local0 >> 5).astype(np.uint64)[:, None] + np.arange(nblk, dtype=np.uint64)[None, :] + base[:, None])
        codes = ((h[:, :, None] >> (np.uint64(2) * np.arange(32, dtype=np.uint64))[None, None, :]) & np.uint64(3)).astype(np.uint8)
        idx = (local0 & 31)[:, None] + np.arange(win, dtype=np.int64)[None, :]
        return ACGT[np.take_along_axis(codes.reshape(len(pos), nblk * 32), idx, axis=1)]

    def plan(self, fragment_length, kmer_size=13, n_hash=3, max_fp=0.01, bin0=0, n_bins=None):
        total_bins = sum(n // fragment_length + 1 for n in self.lengths)        # IBFBuild.cpp:90
        fb, fe = [], []
        for o, n in zip(self.offsets[:-1], self.lengths):
            b, e = capi.fragment_schedule(n, fragment_length, kmer_size)
            fb.append(b + np.uint64(o))
            fe.append(e + np.uint64(o))
        fb, fe = np.concatenate(fb), np.concatenate(fe)
        nb = int(n_bins if n_bins is not None else total_bins)
        return {"frag_begin": fb, "frag_end": fe, "frag_bin": np.arange(bin0, bin0 + len(fb), dtype=np.uint64),
                "n_bins": nb, "n_bits": int(capi.ibf_size_bits(fragment_length, kmer_size, n_hash, max_fp, nb)),
                "sum_seq_len": len(self), "bin_ids_consumed": int(len(fb)), "bins_own": int(total_bins),
                "kmer_size": kmer_size, "n_hash": n_hash}


def revcomp(a):
    return _COMP[a[::-1]]


def build_plan(seqs, fragment_length, kmer_size=13, n_hash=3, max_fp=0.01):
    """IBF::create_filter's host logic (src/IBF/IBFBuild.cpp:421-521) for in-memory records.

    seqs: list of uint8/bytes sequences in file order.  Returns dict with the concatenated
    N-cut bases, fragment begin/end/bin arrays, total bins and the filter size in bits.
    Sequences shorter than k are skipped (IBFBuild.cpp:70-74).  Fragments whose bin id runs past
    the reserved bins (quirk Q3) are kept in the plan; the insert call reports them.
    """
    cut = []
    invalid = 0
    for s in seqs:
        b = s.tobytes() if isinstance(s, np.ndarray) else bytes(s)
        if len(b) < kmer_size:
            invalid += 1
            continue
        cut.append(np.frombuffer(capi.cut_out_nnns(b), dtype=np.uint8))
    total_bins = sum(len(c) // fragment_length + 1 for c in cut)       # IBFBuild.cpp:90
    n_bits = capi.ibf_size_bits(fragment_length, kmer_size, n_hash, max_fp, total_bins)
    fb, fe, fbin = [], [], []
    off = 0
    binid = 0
    for c in cut:
        b, e = capi.fragment_schedule(len(c), fragment_length, kmer_size)
        fb.append(b + np.uint64(off))
        fe.append(e + np.uint64(off))
        fbin.append(np.arange(binid, binid + len(b), dtype=np.uint64))
        binid += len(b)
        off += len(c)
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
    return {
        "bases": cat(cut, np.uint8), "frag_begin": cat(fb, np.uint64), "frag_end": cat(fe, np.uint64),
        "frag_bin": cat(fbin, np.uint64), "n_bins": int(total_bins), "n_bits": int(n_bits),
        "sum_seq_len": int(off), "invalid_seqs": invalid, "bin_ids_consumed": int(binid),
        "kmer_size": kmer_size, "n_hash": n_hash,
    }


def sample_reads(ref, n_reads, read_len, seed, frac_from_ref=0.5, error_rate=0.1, block=65536):
    """n_reads chunks of exactly read_len bases: a fraction sampled from `ref` (uniform position and
    strand) with `error_rate` errors (1/3 substitution, 1/3 insertion, 1/3 deletion), the rest iid.

    Returns (bases uint8 [n_reads*read_len], read_off uint64 [n_reads+1], from_ref bool [n_reads]).
    """
    rng = np.random.default_rng(seed)
    n_reads = int(n_reads)
    out = np.empty((n_reads, read_len), dtype=np.uint8)
    from_ref = rng.random(n_reads) < frac_from_ref
    margin = 48 + read_len // 8
    win = read_len + margin
    for b0 in range(0, n_reads, block):
        b1 = min(n_reads, b0 + block)
        nb = b1 - b0
        blk = ACGT[rng.integers(0, 4, size=(nb, read_len), dtype=np.uint8)]
        idx = np.nonzero(from_ref[b0:b1])[0]
        if idx.size and len(ref) > win:
            pos = rng.integers(0, len(ref) - win, size=idx.size)
            w = ref.windows(pos, win) if hasattr(ref, "windows") else ref[pos[:, None] + np.arange(win)[None, :]]
            strand = rng.random(idx.size) < 0.5
            w[strand] = _COMP[w[strand][:, ::-1]]
            u = rng.random((idx.size, win))
            sub = u < error_rate / 3
            ins = (u >= error_rate / 3) & (u < 2 * error_rate / 3)
            dele = (u >= 2 * error_rate / 3) & (u < error_rate)
            rnd = ACGT[rng.integers(0, 4, size=(idx.size, win), dtype=np.uint8)]
            w = np.where(sub, ACGT[(np.searchsorted(ACGT, w) + 1 + rng.integers(0, 3, size=w.shape)) % 4], w)
            # every window slot emits `rep` symbols: 0 (deleted), 1, or 2 (base + inserted base)
            rep = np.ones((idx.size, win), dtype=np.int64)
            rep[dele] = 0
            rep[ins] = 2
            pair = np.stack([w, rnd], axis=2).reshape(idx.size, 2 * win)
            keep = np.stack([rep >= 1, rep == 2], axis=2).reshape(idx.size, 2 * win)
            rank = np.cumsum(keep, axis=1)
            take = keep & (rank <= read_len)
            enough = rank[:, -1] >= read_len
            rows = np.nonzero(enough)[0]
            vals = pair[take & enough[:, None]].reshape(rows.size, read_len)
            blk[idx[rows]] = vals
            from_ref[b0 + idx[~enough]] = False
        out[b0:b1] = blk
    read_off = (np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(read_len))
    return out.reshape(-1), read_off, from_ref


def ragged_reads(ref, lengths, seed, frac_from_ref=0.5, error_rate=0.05, n_frac=0.0, lower_frac=0.0):
    """Reads of the given lengths (list of ints, 0 allowed) for parity tests; optional N and
    lowercase sprinkling.  Returns (bases, read_off)."""
    rng = np.random.default_rng(seed)
    parts = []
    for L in lengths:
        L = int(L)
        if L and rng.random() < frac_from_ref and len(ref) > L + 1:
            p = int(rng.integers(0, len(ref) - L))
            r = ref[p:p + L].copy()
            if rng.random() < 0.5:
                r = revcomp(r)
            m = rng.random(L) < error_rate
            r[m] = ACGT[rng.integers(0, 4, size=int(m.sum()), dtype=np.uint8)]
        else:
            r = ACGT[rng.integers(0, 4, size=L, dtype=np.uint8)]
        if n_frac > 0 and L:
            r[rng.random(L) < n_frac] = ord("N")
        if lower_frac > 0 and L:
            m = rng.random(L) < lower_frac
            r[m] = r[m] | 0x20
        parts.append(r)
    read_off = np.zeros(len(lengths) + 1, dtype=np.uint64)
    read_off[1:] = np.cumsum([len(p) for p in parts])
    bases = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
    return bases.astype(np.uint8), read_off
