"""ctypes binding of librb_ibf.so (the C ABI in include/rb_ibf.h).

Fails loudly when the library is missing or no CUDA device is usable: there is no
CPU or PyTorch fallback anywhere in this package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_LIB = os.path.join(_HERE, "lib", "librb_ibf.so")
_lib = None

RB_OK = 0
STATUS = {0: "ok", 1: "NullFilterException", 2: "ShortReadException", 3: "CountKmerException",
          4: "ParseIBFFileException", 5: "MissingIBFFileException", 6: "StoreFilterException",
          7: "InsertSequenceException", 8: "InvalidConfigException", 9: "out of memory", 10: "CUDA error",
          11: "no CUDA device", 12: "invalid argument"}
MAX_LUT = 4
LUT_SIZE = 65536

EXPORTS = [
    "rb_status_string", "rb_last_error", "rb_device_count", "rb_ibf_size_bits", "rb_calculate_ci",
    "rb_threshold_lut", "rb_cut_out_nnns", "rb_fragment_schedule", "rb_ibf_create", "rb_ibf_load",
    "rb_ibf_load_shard", "rb_ibf_create_shard", "rb_ibf_from_words", "rb_ibf_store", "rb_ibf_download", "rb_ibf_free", "rb_ibf_info",
    "rb_ibf_device_words", "rb_ibf_device_kmer_table", "rb_ibf_insert_batch", "rb_ibf_insert_batch_dev", "rb_ibf_count_batch",
    "rb_ibf_count_batch_dev", "rb_keys_decode_dev", "rb_set_count_kernel", "rb_set_insert_kernel", "rb_kernel_launches",
    "rb_microbench_gather", "rb_microbench_gather_coop", "rb_set_l2_fetch_granularity", "rb_get_l2_fetch_granularity",
    "rb_ibf_enable_kmer_table", "rb_ibf_resize_bins", "rb_host_pack_info", "rb_transfer_bytes",
    "rb_ibf_transfer_policy", "rb_ibf_count_traffic_dev", "rb_synth_bases_dev",
    "rb_ibf_enable_kmer_tables", "rb_threshold_lut_raw", "rb_ibf_count_batch_sharded", "rb_keys_combine_nccl",
    "rb_microbench_host_read",
]


class RBError(RuntimeError):
    def __init__(self, status, message=""):
        super().__init__("%s (%d)%s" % (STATUS.get(status, "?"), status, ": " + message if message else ""))
        self.status = status


class _Info(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_bins", "n_hash", "kmer_size", "n_bits", "bin_width", "n_blocks",
                                          "col_begin", "col_words", "bin_begin", "n_bins_local", "device_bytes")] + \
               [("device", C.c_int32), ("shard", C.c_int32), ("n_shards", C.c_int32), ("kmer_table_span", C.c_int32),
                ("kmer_table_bytes", C.c_uint64), ("kmer_table_kind", C.c_int32)]


def lib_path():
    return _LIB


def build_library(force=False):
    """Compile librb_ibf.so in-tree with nvcc for sm_100a (see csrc/Makefile)."""
    srcs = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cu", ".cuh", ".cpp", ".hpp", "Makefile"))]
    root = os.path.dirname(_HERE)
    srcs += [os.path.join(root, "include", f) for f in os.listdir(os.path.join(root, "include")) if f.endswith((".h", ".hpp"))]
    srcs.append(os.path.join(root, "tools", "rb_readbouncer.cpp"))
    exe = os.path.join(_HERE, "bin", "rb_readbouncer")
    stale = (not os.path.exists(_LIB) or not os.path.exists(exe)
             or any(os.path.getmtime(s) > min(os.path.getmtime(_LIB), os.path.getmtime(exe)) for s in srcs))
    if force or stale:
        cmd = ["make", "-C", _CSRC, "-j4"] + (["-B"] if force else [])
        subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    return _LIB


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB):
        raise RuntimeError("librb_ibf.so is not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "or `make -C readbouncer_b200/csrc`.  There is no CPU fallback." % _LIB)
    L = C.CDLL(_LIB)
    vp, u64, u32, i32, dbl = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_double
    ip = C.POINTER(C.c_int)
    sig = {
        "rb_status_string": (C.c_char_p, [i32]),
        "rb_last_error": (C.c_char_p, []),
        "rb_device_count": (i32, []),
        "rb_ibf_size_bits": (u64, [u64, u32, u32, dbl, u64]),
        "rb_calculate_ci": (i32, [dbl, u32, u32, dbl, vp, vp]),
        "rb_threshold_lut": (i32, [dbl, dbl, u32, vp]),
        "rb_cut_out_nnns": (u64, [C.c_char_p, u64, C.c_char_p]),
        "rb_fragment_schedule": (u64, [u64, u64, u32, vp, vp, u64]),
        "rb_ibf_create": (vp, [u64, u32, u32, u64, i32, ip]),
        "rb_ibf_load": (vp, [C.c_char_p, i32, ip]),
        "rb_ibf_load_shard": (vp, [C.c_char_p, i32, i32, i32, ip]),
        "rb_ibf_create_shard": (vp, [u64, u32, u32, u64, i32, i32, i32, ip]),
        "rb_ibf_from_words": (vp, [vp, u64, u32, u32, u64, i32, i32, i32, ip]),
        "rb_ibf_store": (i32, [vp, C.c_char_p]),
        "rb_ibf_download": (i32, [vp, vp, u64]),
        "rb_ibf_free": (None, [vp]),
        "rb_ibf_info": (i32, [vp, C.POINTER(_Info)]),
        "rb_ibf_device_words": (vp, [vp]),
        "rb_ibf_device_kmer_table": (vp, [vp]),
        "rb_ibf_insert_batch": (i32, [vp, vp, u64, vp, vp, vp, u64, vp]),
        "rb_ibf_insert_batch_dev": (i32, [vp, vp, vp, vp, vp, u64, u64, vp]),
        "rb_ibf_count_batch": (i32, [vp, vp, vp, u64, vp, u32, vp, vp, vp, vp, vp, vp, vp]),
        "rb_ibf_count_batch_dev": (i32, [vp, vp, vp, u64, u32, vp, u32, vp, vp, vp, vp, vp]),
        "rb_keys_decode_dev": (i32, [vp, u64, vp, vp, vp, i32, vp]),
        "rb_set_count_kernel": (i32, [i32]),
        "rb_set_insert_kernel": (i32, [i32]),
        "rb_kernel_launches": (u64, []),
        "rb_microbench_gather": (i32, [vp, u64, u32, u64, u32, vp, vp]),
        "rb_microbench_gather_coop": (i32, [vp, u64, u32, u32, u64, u32, vp, vp]),
        "rb_set_l2_fetch_granularity": (i32, [i32, u32]),
        "rb_get_l2_fetch_granularity": (i32, [i32, vp]),
        "rb_ibf_enable_kmer_table": (i32, [vp, u64, vp]),
        "rb_ibf_resize_bins": (i32, [vp, u64, vp]),
        "rb_host_pack_info": (i32, [vp, vp]),
        "rb_transfer_bytes": (i32, [vp, vp]),
        "rb_ibf_transfer_policy": (i32, [vp, vp, vp, vp]),
        "rb_ibf_count_traffic_dev": (i32, [vp, vp, vp, u64, u32, vp, vp, vp, vp]),
        "rb_synth_bases_dev": (i32, [vp, u64, u64, u64, vp]),
        "rb_ibf_enable_kmer_tables": (i32, [vp, u32, u64, vp]),
        "rb_threshold_lut_raw": (i32, [dbl, dbl, u32, vp]),
        "rb_ibf_count_batch_sharded": (i32, [vp, u32, vp, vp, u64, vp, u32, vp, vp, vp, vp]),
        "rb_keys_combine_nccl": (i32, [vp, vp, u64, vp]),
        "rb_microbench_host_read": (i32, [vp, u64, u32, vp]),
    }
    assert sorted(sig) == sorted(EXPORTS)
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _check(status):
    if status != RB_OK:
        raise RBError(status, lib().rb_last_error().decode(errors="replace"))


def _np_ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _dev_ptr(t):
    """Device pointer of a torch CUDA tensor (or an int / None passed through)."""
    if t is None:
        return None
    if isinstance(t, int):
        return C.c_void_p(t)
    assert t.is_cuda and t.is_contiguous()
    return C.c_void_p(t.data_ptr())


def _stream_ptr(stream):
    if stream is None:
        return None
    if isinstance(stream, int):
        return C.c_void_p(stream)
    return C.c_void_p(stream.cuda_stream)      # torch.cuda.Stream


# ---- scalar helpers -------------------------------------------------------------------------------
def device_count():
    return int(lib().rb_device_count())


def kernel_launches():
    return int(lib().rb_kernel_launches())


def set_count_kernel(which):
    _check(lib().rb_set_count_kernel(int(which)))


def set_insert_kernel(which):
    _check(lib().rb_set_insert_kernel(int(which)))


def microbench_gather(d_buf, n_rows, row_bytes, probes_per_thread, n_blocks, d_sink, stream=None):
    _check(lib().rb_microbench_gather(_dev_ptr(d_buf), n_rows, row_bytes, probes_per_thread, n_blocks,
                                      _dev_ptr(d_sink), _stream_ptr(stream)))


def microbench_gather_coop(d_buf, n_rows, row_bytes, lane_bytes, probes_per_group, n_blocks, d_sink, stream=None):
    _check(lib().rb_microbench_gather_coop(_dev_ptr(d_buf), n_rows, row_bytes, lane_bytes, probes_per_group, n_blocks,
                                           _dev_ptr(d_sink), _stream_ptr(stream)))


def synth_bases_dev(d_out, n, seed, start=0, stream=None):
    """n synthetic bases of stream `seed` from position `start` into a device buffer (see synth.hash_bases for the host twin)."""
    _check(lib().rb_synth_bases_dev(_dev_ptr(d_out), int(n), int(seed), int(start), _stream_ptr(stream)))


def host_read_gbs(buf, reps=4):
    """GB/s at which the packer's host threads stream-read a numpy buffer (rb_microbench_host_read)."""
    g = C.c_double(0)
    _check(lib().rb_microbench_host_read(_np_ptr(buf), buf.nbytes, reps, C.byref(g)))
    return float(g.value)


def host_pack_info():
    t, a = C.c_int32(0), C.c_int32(0)
    _check(lib().rb_host_pack_info(C.byref(t), C.byref(a)))
    return {"threads": int(t.value), "isa": {0: "scalar", 2: "avx2", 5: "avx512vbmi"}.get(int(a.value), str(a.value))}


def transfer_bytes():
    h, d = C.c_uint64(0), C.c_uint64(0)
    _check(lib().rb_transfer_bytes(C.byref(h), C.byref(d)))
    return int(h.value), int(d.value)


def set_l2_fetch_granularity(nbytes, device=0):
    _check(lib().rb_set_l2_fetch_granularity(device, nbytes))


def get_l2_fetch_granularity(device=0):
    v = C.c_uint32(0)
    _check(lib().rb_get_l2_fetch_granularity(device, C.addressof(v)))
    return int(v.value)


def ibf_size_bits(fragment_length, kmer_size=13, n_hash=3, max_fp=0.01, n_bins=1):
    return int(lib().rb_ibf_size_bits(fragment_length, kmer_size, n_hash, max_fp, n_bins))


def calculate_ci(error_rate, kmer_size, readlen, significance=0.95):
    lo, hi = C.c_uint16(), C.c_uint16()
    _check(lib().rb_calculate_ci(error_rate, kmer_size, readlen, significance, C.addressof(lo), C.addressof(hi)))
    return lo.value, hi.value


def threshold_lut(error_rate, kmer_size, significance=0.95, raw=False):
    """uint16[65536] thresholds by read length; raw=True skips the (0, 1) range check (the reference's retry rate)."""
    out = np.zeros(LUT_SIZE, np.uint16)
    fn = lib().rb_threshold_lut_raw if raw else lib().rb_threshold_lut
    _check(fn(error_rate, significance, kmer_size, _np_ptr(out)))
    return out


def count_batch_sharded(shards, bases, read_off, thr_lut):
    """Bin shards on several devices of this process: one call, keys folded over NVLink (rb_ibf_count_batch_sharded)."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    read_off = np.ascontiguousarray(read_off, dtype=np.uint64)
    lut = np.ascontiguousarray(thr_lut, dtype=np.uint16).reshape(-1, LUT_SIZE)
    n, n_lut = read_off.size - 1, lut.shape[0]
    res = {"max_count": np.zeros((n_lut, n), np.uint16), "hit": np.zeros((n_lut, n), np.uint8),
           "argmax_bin": np.zeros((n_lut, n), np.uint32), "read_flag": np.zeros(n, np.uint8)}
    arr = (C.c_void_p * len(shards))(*[f._h for f in shards])
    _check(lib().rb_ibf_count_batch_sharded(arr, len(shards), _np_ptr(bases), _np_ptr(read_off), n, _np_ptr(lut), n_lut,
                                            _np_ptr(res["max_count"]), _np_ptr(res["hit"]), _np_ptr(res["argmax_bin"]),
                                            _np_ptr(res["read_flag"])))
    if n_lut == 1:
        for key in ("max_count", "hit", "argmax_bin"):
            res[key] = res[key][0]
    return res


def enable_kmer_tables(filters, total_bytes=0, stream=None):
    """Joint k-mer table plan for all filters a caller classifies against (rb_ibf_enable_kmer_tables)."""
    arr = (C.c_void_p * len(filters))(*[f._h for f in filters])
    _check(lib().rb_ibf_enable_kmer_tables(arr, len(filters), int(total_bytes), _stream_ptr(stream)))


def cut_out_nnns(seq):
    s = seq if isinstance(seq, (bytes, bytearray)) else seq.encode()
    out = C.create_string_buffer(len(s) + 1)
    n = lib().rb_cut_out_nnns(s, len(s), out)
    return out.raw[:n]


def fragment_schedule(seqlen, fragment_length, kmer_size):
    n = int(lib().rb_fragment_schedule(seqlen, fragment_length, kmer_size, None, None, 0))
    b = np.zeros(max(n, 1), np.uint64)
    e = np.zeros(max(n, 1), np.uint64)
    lib().rb_fragment_schedule(seqlen, fragment_length, kmer_size, _np_ptr(b), _np_ptr(e), n)
    return b[:n], e[:n]


def keys_decode(keys):
    """numpy decode of packed summary keys (see RB_KEY_* in rb_ibf.h)."""
    keys = np.asarray(keys, dtype=np.uint64)
    hit = ((keys >> np.uint64(48)) & np.uint64(1)).astype(np.uint8)
    mx = ((keys >> np.uint64(32)) & np.uint64(0xFFFF)).astype(np.uint16)
    am = np.where(keys != 0, ~(keys & np.uint64(0xFFFFFFFF)) & np.uint64(0xFFFFFFFF), np.uint64(0xFFFFFFFF))
    return mx, hit, am.astype(np.uint32)


# ---- filter handle ------------------------------------------------------------------------------------
class IBF:
    """Device-resident Interleaved Bloom Filter (wraps an rb_ibf*)."""

    def __init__(self, handle):
        self._h = handle
        info = _Info()
        _check(lib().rb_ibf_info(self._h, C.byref(info)))
        for name, _ in _Info._fields_:
            if name not in ("kmer_table_span", "kmer_table_bytes", "kmer_table_kind"):      # the latter changes over time: see method
                setattr(self, name, int(getattr(info, name)))
        self.k = self.kmer_size
        self.n_local_words = self.device_bytes // 8

    @staticmethod
    def _wrap(h, st):
        if not h:
            raise RBError(st.value, lib().rb_last_error().decode(errors="replace"))
        return IBF(h)

    @classmethod
    def create(cls, n_bins, n_hash, kmer_size, n_bits, device=0):
        st = C.c_int(0)
        return cls._wrap(lib().rb_ibf_create(n_bins, n_hash, kmer_size, n_bits, device, C.byref(st)), st)

    @classmethod
    def create_shard(cls, n_bins, n_hash, kmer_size, n_bits, shard, n_shards, device=0):
        """Zero-filled bin shard of a filter that is never held in one place (rb_ibf_create_shard)."""
        st = C.c_int(0)
        return cls._wrap(lib().rb_ibf_create_shard(n_bins, n_hash, kmer_size, n_bits, device, shard, n_shards, C.byref(st)), st)

    @classmethod
    def load(cls, path, device=0, shard=0, n_shards=1):
        st = C.c_int(0)
        return cls._wrap(lib().rb_ibf_load_shard(str(path).encode(), device, shard, n_shards, C.byref(st)), st)

    @classmethod
    def from_words(cls, words, n_bins, n_hash, kmer_size, n_bits, device=0, shard=0, n_shards=1):
        words = np.ascontiguousarray(words, dtype=np.uint64)
        assert words.size >= n_bits // 64
        st = C.c_int(0)
        return cls._wrap(lib().rb_ibf_from_words(_np_ptr(words), n_bins, n_hash, kmer_size, n_bits, device, shard,
                                                 n_shards, C.byref(st)), st)

    def store(self, path):
        _check(lib().rb_ibf_store(self._h, str(path).encode()))

    def download(self):
        out = np.zeros(self.n_local_words, np.uint64)
        _check(lib().rb_ibf_download(self._h, _np_ptr(out), out.size))
        return out

    def resize_bins(self, new_n_bins, stream=None):
        _check(lib().rb_ibf_resize_bins(self._h, new_n_bins, _stream_ptr(stream)))
        self.__init__(self._h)                       # refresh the cached geometry

    def enable_kmer_table(self, max_table_bytes=0, stream=None):
        """Build the direct k-mer table now (0 = automatic budget).  See rb_ibf.h."""
        _check(lib().rb_ibf_enable_kmer_table(self._h, max_table_bytes, _stream_ptr(stream)))

    def disable_kmer_table(self):
        _check(lib().rb_ibf_enable_kmer_table(self._h, 0xFFFFFFFFFFFFFFFF, None))

    def kmer_table_bytes(self):
        info = _Info()
        _check(lib().rb_ibf_info(self._h, C.byref(info)))
        return int(info.kmer_table_bytes)

    def kmer_table_span(self):
        info = _Info()
        _check(lib().rb_ibf_info(self._h, C.byref(info)))
        return int(info.kmer_table_span)

    def kmer_table_kind(self):
        info = _Info()
        _check(lib().rb_ibf_info(self._h, C.byref(info)))
        return int(info.kmer_table_kind)

    def transfer_policy(self):
        """How large host-buffer batches are shipped: packed bit planes or ASCII (rb_ibf_transfer_policy)."""
        c, a, b = C.c_int32(0), C.c_double(0), C.c_double(0)
        _check(lib().rb_ibf_transfer_policy(self._h, C.byref(c), C.byref(a), C.byref(b)))
        return {"choice": {-1: "undecided", 0: "packed", 1: "ascii"}[int(c.value)],
                "ns_per_base_packed": float(a.value), "ns_per_base_ascii": float(b.value)}

    def device_words_ptr(self):
        return int(lib().rb_ibf_device_words(self._h) or 0)

    def device_kmer_table_ptr(self):
        return int(lib().rb_ibf_device_kmer_table(self._h) or 0)

    # ---- build -----------------------------------------------------------------------------------
    def insert_batch(self, bases, frag_begin, frag_end, frag_bin, stream=None):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        fb = np.ascontiguousarray(frag_begin, dtype=np.uint64)
        fe = np.ascontiguousarray(frag_end, dtype=np.uint64)
        fbin = np.ascontiguousarray(frag_bin, dtype=np.uint64)
        assert fb.size == fe.size == fbin.size
        _check(lib().rb_ibf_insert_batch(self._h, _np_ptr(bases), bases.size, _np_ptr(fb), _np_ptr(fe), _np_ptr(fbin),
                                         fb.size, _stream_ptr(stream)))

    def insert_batch_dev(self, d_bases, d_frag_begin, d_frag_end, d_frag_bin, n_frags, max_frag_len=0, stream=None):
        _check(lib().rb_ibf_insert_batch_dev(self._h, _dev_ptr(d_bases), _dev_ptr(d_frag_begin), _dev_ptr(d_frag_end),
                                             _dev_ptr(d_frag_bin), n_frags, max_frag_len, _stream_ptr(stream)))

    # ---- classify --------------------------------------------------------------------------------
    def count_batch(self, bases, read_off, thr_lut, dense=False, stream=None):
        """Host-buffer classify call.  thr_lut: uint16 [n_lut, 65536] (or [65536]).
        Returns dict with max_count/hit/argmax_bin of shape [n_lut, n] (squeezed when n_lut == 1),
        read_flag [n] and, with dense=True, counts_fwd/counts_rev [n, n_bins_local]."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        read_off = np.ascontiguousarray(read_off, dtype=np.uint64)
        lut = np.ascontiguousarray(thr_lut, dtype=np.uint16).reshape(-1, LUT_SIZE)
        n, n_lut = read_off.size - 1, lut.shape[0]
        res = {
            "max_count": np.zeros((n_lut, n), np.uint16), "hit": np.zeros((n_lut, n), np.uint8),
            "argmax_bin": np.zeros((n_lut, n), np.uint32), "read_flag": np.zeros(n, np.uint8),
            "counts_fwd": np.zeros((n, self.n_bins_local), np.uint16) if dense else None,
            "counts_rev": np.zeros((n, self.n_bins_local), np.uint16) if dense else None,
        }
        _check(lib().rb_ibf_count_batch(self._h, _np_ptr(bases), _np_ptr(read_off), n, _np_ptr(lut), n_lut,
                                        _np_ptr(res["counts_fwd"]), _np_ptr(res["counts_rev"]),
                                        _np_ptr(res["max_count"]), _np_ptr(res["hit"]), _np_ptr(res["argmax_bin"]),
                                        _np_ptr(res["read_flag"]), _stream_ptr(stream)))
        if n_lut == 1:
            for key in ("max_count", "hit", "argmax_bin"):
                res[key] = res[key][0]
        return res

    def count_batch_dev(self, d_bases, d_read_off, n_reads, d_thr_lut, n_lut, d_keys, max_read_len=0,
                        d_counts_fwd=None, d_counts_rev=None, d_read_flag=None, stream=None):
        """Device-pointer classify call: enqueues on `stream`, no host synchronisation."""
        _check(lib().rb_ibf_count_batch_dev(self._h, _dev_ptr(d_bases), _dev_ptr(d_read_off), n_reads, max_read_len,
                                            _dev_ptr(d_thr_lut), n_lut, _dev_ptr(d_keys), _dev_ptr(d_counts_fwd),
                                            _dev_ptr(d_counts_rev), _dev_ptr(d_read_flag), _stream_ptr(stream)))

    def count_traffic_dev(self, d_bases, d_read_off, n_reads, n_lut=1, stream=None):
        """(table_bytes at 128-byte line granularity, table_requests, io_bytes) of one count launch over this batch."""
        tb, tr, io = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        _check(lib().rb_ibf_count_traffic_dev(self._h, _dev_ptr(d_bases), _dev_ptr(d_read_off), n_reads, n_lut,
                                              C.byref(tb), C.byref(tr), C.byref(io), _stream_ptr(stream)))
        return int(tb.value), int(tr.value), int(io.value)

    def close(self):
        if getattr(self, "_h", None):
            lib().rb_ibf_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
