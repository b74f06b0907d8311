"""ctypes binding of the CPU oracle (oracle/ibf_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(readbouncer_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libibf_oracle.so")
_lib = None

u64p = C.POINTER(C.c_uint64)
u16p = C.POINTER(C.c_uint16)
u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)

STATUS_NAMES = {0: "OK", 1: "NullFilter", 2: "ShortRead", 3: "CountKmer", 4: "ParseIBFFile",
                5: "MissingIBFFile", 6: "StoreFilter", 7: "InsertSequence", 8: "InvalidConfig", 9: "Alloc"}


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("ibf_oracle.c", "ibf_oracle.h", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)
             or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B" if force else "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    vp = C.c_void_p
    sig = {
        "orc_dna5": (C.c_uint8, [C.c_char]),
        "orc_ibf_create": (vp, [C.c_uint64] * 4),
        "orc_ibf_load": (vp, [C.c_char_p, C.POINTER(C.c_int)]),
        "orc_ibf_store": (C.c_int, [vp, C.c_char_p]),
        "orc_ibf_free": (None, [vp]),
        "orc_ibf_resize_bins": (C.c_int, [vp, C.c_uint64]),
        "orc_ibf_words": (u64p, [vp]),
        "orc_ibf_info": (None, [vp, u64p, u64p, u64p, u64p, u64p]),
        "orc_kmer_hash": (C.c_uint64, [C.c_char_p, C.c_uint64]),
        "orc_hash_row": (C.c_uint64, [vp, C.c_uint64, C.c_uint]),
        "orc_insert": (None, [vp, C.c_char_p, C.c_uint64, C.c_uint64]),
        "orc_count": (None, [vp, C.c_char_p, C.c_uint64, C.c_int, u16p]),
        "orc_filter_size_bits": (C.c_uint64, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, C.c_uint64]),
        "orc_calculate_ci": (None, [C.c_double, C.c_uint8, C.c_uint32, C.c_double, u16p, u16p]),
        "orc_threshold": (C.c_uint16, [C.c_double, C.c_uint64, C.c_uint64, C.c_double]),
        "orc_threshold_lut": (None, [C.c_double, C.c_uint64, C.c_double, u16p]),
        "orc_cut_out_nnns": (C.c_uint64, [C.c_char_p, C.c_uint64, C.c_char_p]),
        "orc_bins_for_sequence": (C.c_uint64, [C.c_uint64, C.c_uint64]),
        "orc_fragment_schedule": (C.c_uint64, [C.c_uint64, C.c_uint64, C.c_uint64, u64p, u64p, C.c_uint64]),
        "orc_select_matches": (C.c_int, [u16p, u16p, C.c_uint64, C.c_uint16]),
        "orc_max_matches": (C.c_uint64, [u16p, u16p, C.c_uint64, C.c_uint16]),
        "orc_count_matches": (C.c_uint64, [vp, C.c_char_p, C.c_uint64, C.c_double, C.c_double]),
        "orc_classify_any": (C.c_int, [C.POINTER(vp), C.c_uint64, C.c_char_p, C.c_uint64, C.c_double,
                                       C.c_double, C.POINTER(C.c_int)]),
        "orc_classify_best": (C.c_int, [C.POINTER(vp), C.c_uint64, C.c_char_p, C.c_uint64, C.c_double,
                                        C.c_double, C.POINTER(C.c_int)]),
        "orc_classify_pair": (C.c_int, [C.POINTER(vp), C.c_uint64, C.POINTER(vp), C.c_uint64, C.c_char_p,
                                        C.c_uint64, C.c_double, C.c_double, u64p, u64p]),
        "orc_check_unblock": (C.c_int, [C.POINTER(vp), C.c_uint64, C.POINTER(vp), C.c_uint64, C.c_char_p,
                                        C.c_uint64, C.c_double, C.c_double, C.POINTER(C.c_int)]),
        "orc_count_batch": (C.c_int, [vp, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
        "orc_insert_batch": (C.c_int, [vp, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                       C.c_int]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _b(s):
    if isinstance(s, (bytes, bytearray)):
        return s
    if isinstance(s, np.ndarray):
        return s.tobytes()
    return s.encode()


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleError(Exception):
    def __init__(self, status):
        super().__init__(STATUS_NAMES.get(status, str(status)))
        self.status = status


class OracleIBF:
    """Host-memory IBF with the SeqAn layout (see ibf_oracle.h)."""

    def __init__(self, handle):
        self._h = handle
        v = [C.c_uint64() for _ in range(5)]
        lib().orc_ibf_info(self._h, *[C.byref(x) for x in v])
        self.n_bins, self.n_hash, self.k, self.n_bits, self.n_words = [int(x.value) for x in v]
        self.bin_width = (self.n_bins + 63) // 64
        self.n_blocks = self.n_bits // (64 * self.bin_width)

    @classmethod
    def create(cls, n_bins, n_hash, k, n_bits):
        h = lib().orc_ibf_create(n_bins, n_hash, k, n_bits)
        if not h:
            raise OracleError(8)
        return cls(h)

    @classmethod
    def load(cls, path):
        st = C.c_int(0)
        h = lib().orc_ibf_load(_b(str(path)), C.byref(st))
        if not h:
            raise OracleError(st.value)
        return cls(h)

    def store(self, path):
        st = lib().orc_ibf_store(self._h, _b(str(path)))
        if st:
            raise OracleError(st)

    def resize_bins(self, new_n_bins):
        st = lib().orc_ibf_resize_bins(self._h, new_n_bins)
        if st:
            raise OracleError(st)
        self.__init__(self._h)

    def words(self):
        """numpy view (no copy) of all words including the 4-word metadata tail."""
        p = lib().orc_ibf_words(self._h)
        return np.ctypeslib.as_array(p, shape=(self.n_words,))

    def insert(self, text, bin_id):
        t = _b(text)
        lib().orc_insert(self._h, t, len(t), bin_id)

    def count(self, text, revcomp=False):
        t = _b(text)
        out = np.zeros(self.n_bins, dtype=np.uint16)
        lib().orc_count(self._h, t, len(t), int(revcomp), out.ctypes.data_as(u16p))
        return out

    def count_matches(self, text, error_rate=0.1, significance=0.95):
        t = _b(text)
        return int(lib().orc_count_matches(self._h, t, len(t), error_rate, significance))

    def hash_row(self, kmer_value, i):
        return int(lib().orc_hash_row(self._h, kmer_value, i))

    def count_batch(self, bases, read_off, thr_lut, dense=True, n_threads=1):
        """bases: uint8 array; read_off: uint64[n+1]; returns dict of arrays."""
        n = len(read_off) - 1
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        read_off = np.ascontiguousarray(read_off, dtype=np.uint64)
        thr_lut = np.ascontiguousarray(thr_lut, dtype=np.uint16)
        res = {
            "counts_fwd": np.zeros((n, self.n_bins), np.uint16) if dense else None,
            "counts_rev": np.zeros((n, self.n_bins), np.uint16) if dense else None,
            "max_count": np.zeros(n, np.uint16), "hit": np.zeros(n, np.uint8),
            "argmax_bin": np.zeros(n, np.uint32), "short_read": np.zeros(n, np.uint8),
        }
        st = lib().orc_count_batch(self._h, _ptr(bases), _ptr(read_off), n, _ptr(thr_lut),
                                   _ptr(res["counts_fwd"]), _ptr(res["counts_rev"]), _ptr(res["max_count"]),
                                   _ptr(res["hit"]), _ptr(res["argmax_bin"]), _ptr(res["short_read"]),
                                   int(n_threads))
        if st:
            raise OracleError(st)
        return res

    def insert_batch(self, bases, frag_begin, frag_end, frag_bin, n_threads=1):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        fb = np.ascontiguousarray(frag_begin, dtype=np.uint64)
        fe = np.ascontiguousarray(frag_end, dtype=np.uint64)
        fbin = np.ascontiguousarray(frag_bin, dtype=np.uint64)
        st = lib().orc_insert_batch(self._h, _ptr(bases), _ptr(fb), _ptr(fe), _ptr(fbin), len(fb),
                                    int(n_threads))
        if st:
            raise OracleError(st)

    def close(self):
        if self._h:
            lib().orc_ibf_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _handles(filters):
    arr = (C.c_void_p * max(1, len(filters)))(*[f._h for f in filters])
    return arr


def dna5(c):
    return int(lib().orc_dna5(_b(c)[:1]))


def kmer_hash(text, k):
    return int(lib().orc_kmer_hash(_b(text), k))


def filter_size_bits(fragment_length, k, n_hash, max_fp, n_bins):
    return int(lib().orc_filter_size_bits(fragment_length, k, n_hash, max_fp, n_bins))


def calculate_ci(r, k, readlen, confidence):
    lo, hi = C.c_uint16(), C.c_uint16()
    lib().orc_calculate_ci(r, k, readlen, confidence, C.byref(lo), C.byref(hi))
    return lo.value, hi.value


def threshold(r, k, readlen, confidence=0.95):
    return int(lib().orc_threshold(r, k, readlen, confidence))


def threshold_lut(r, k, confidence=0.95):
    out = np.zeros(65536, np.uint16)
    lib().orc_threshold_lut(r, k, confidence, out.ctypes.data_as(u16p))
    return out


def cut_out_nnns(seq):
    s = _b(seq)
    out = C.create_string_buffer(len(s) + 1)
    n = lib().orc_cut_out_nnns(s, len(s), out)
    return out.raw[:n]


def bins_for_sequence(cut_len, fragment_length):
    return int(lib().orc_bins_for_sequence(cut_len, fragment_length))


def fragment_schedule(seqlen, fragment_length, k):
    n = int(lib().orc_fragment_schedule(seqlen, fragment_length, k, None, None, 0))
    b = np.zeros(max(n, 1), np.uint64)
    e = np.zeros(max(n, 1), np.uint64)
    lib().orc_fragment_schedule(seqlen, fragment_length, k, b.ctypes.data_as(u64p), e.ctypes.data_as(u64p), n)
    return b[:n], e[:n]


def classify_any(filters, read, error_rate=0.1, significance=0.95):
    st = C.c_int(0)
    t = _b(read)
    r = lib().orc_classify_any(_handles(filters), len(filters), t, len(t), error_rate, significance, C.byref(st))
    if st.value:
        raise OracleError(st.value)
    return bool(r)


def classify_best(filters, read, error_rate=0.1, significance=0.95):
    st = C.c_int(0)
    t = _b(read)
    r = lib().orc_classify_best(_handles(filters), len(filters), t, len(t), error_rate, significance, C.byref(st))
    if st.value:
        raise OracleError(st.value)
    return int(r)


def classify_pair(filt1, filt2, read, error_rate=0.1, significance=0.95):
    a, b = C.c_uint64(), C.c_uint64()
    t = _b(read)
    st = lib().orc_classify_pair(_handles(filt1), len(filt1), _handles(filt2), len(filt2), t, len(t),
                                 error_rate, significance, C.byref(a), C.byref(b))
    if st:
        raise OracleError(st)
    return int(a.value), int(b.value)


def check_unblock(deplete, target, read, error_rate=0.1, significance=0.95):
    st = C.c_int(0)
    t = _b(read)
    r = lib().orc_check_unblock(_handles(deplete), len(deplete), _handles(target), len(target), t, len(t),
                                error_rate, significance, C.byref(st))
    if st.value:
        raise OracleError(st.value)
    return int(r)


def classify_reads_serial(reads, dep, tgt, chunk, max_chunks, err):
    """The serial loop of classify_reads (src/main/classify.hpp:229-303) read by read, chunk by chunk, with the oracle's
    classify overloads: per read the index of the target filter it is assigned to, -2 a depletion-mode hit, -1 unclassified,
    -3 failed (ShortReadException), -4 too short (len < chunk_length).  `dep` / `tgt` are lists of OracleIBF."""
    assign = []
    for seq in reads:
        if len(seq) < chunk:
            assign.append(-4)
            continue
        a = -1
        for i in range(max_chunks):
            if i * chunk >= len(seq):
                break
            frag = seq[i * chunk:min((i + 1) * chunk, len(seq))]
            if dep and tgt:                                   # classify_deplete_target, classify.hpp:58-111
                t0, d0 = classify_pair(tgt, dep, frag, err)
                ok = False
                if t0 > 0:
                    if d0 > 0:
                        t1, d1 = classify_pair(tgt, dep, frag, err - 0.02)
                        ok = t1 > 0 and d1 == 0
                    else:
                        ok = True
                if ok:
                    a = classify_best(tgt, frag, err)
            elif dep:
                if len(frag) < dep[0].k:
                    a = -3
                    break
                if classify_best(dep, frag, err) > -1:
                    a = -2
            else:
                if len(frag) < tgt[0].k:
                    a = -3
                    break
                b = classify_best(tgt, frag, err)
                if b != -1:
                    a = b
            if a != -1:
                break
        assign.append(a)
    return assign


def build_from_sequences(seqs, fragment_length, k=13, n_hash=3, max_fp=0.01, passes=1, n_threads=1):
    """IBF::create_filter restated (src/IBF/IBFBuild.cpp:421-521) on in-memory records.

    seqs: list of raw sequence strings/bytes in file order.  `passes=2` replays
    the queue twice, as the test build that produced the golden .ibf files did
    (SURVEY Appendix B).  Returns (OracleIBF, stats dict).
    """
    cut = []
    invalid = 0
    for s in seqs:
        s = _b(s)
        if len(s) < k:          # IBFBuild.cpp:70-74
            invalid += 1
            continue
        cut.append(cut_out_nnns(s))
    queue = cut * passes
    total_bins = sum(bins_for_sequence(len(s), fragment_length) for s in queue)
    n_bits = filter_size_bits(fragment_length, k, n_hash, max_fp, total_bins)
    f = OracleIBF.create(total_bins, n_hash, k, n_bits)
    binid = 0
    dropped = 0
    bases, fb, fe, fbin = [], [], [], []
    off = 0
    for s in queue:
        b, e = fragment_schedule(len(s), fragment_length, k)
        for bb, ee in zip(b, e):
            if binid >= total_bins:
                dropped += 1
            fb.append(off + int(bb)); fe.append(off + int(ee)); fbin.append(binid)
            binid += 1
        bases.append(np.frombuffer(s, np.uint8))
        off += len(s)
    if fb:
        f.insert_batch(np.concatenate(bases) if bases else np.zeros(0, np.uint8), fb, fe, fbin, n_threads)
    stats = {"totalBinsBinId": total_bins, "sumSeqLen": sum(len(s) for s in queue), "invalidSeqs": invalid,
             "filter_size_bits": n_bits, "bin_ids_consumed": binid, "dropped_fragments": dropped}
    return f, stats
