"""readbouncer_b200 -- B200-native IBF read classifier (ReadBouncer's hot path).

The product is the C-ABI shared library built from readbouncer_b200/csrc
(include/rb_ibf.h).  This package only binds it with ctypes for the tests, the
benchmark and Python callers; the C++ mirror of ReadBouncer's interleave::
classes lives in include/rb_interleave.hpp.
"""
from .capi import (IBF, count_batch_sharded, enable_kmer_tables, get_l2_fetch_granularity, set_l2_fetch_granularity, RBError, build_library, calculate_ci, cut_out_nnns, device_count, fragment_schedule,
                   ibf_size_bits, kernel_launches, keys_decode, lib, lib_path, host_pack_info, transfer_bytes, microbench_gather, microbench_gather_coop, set_count_kernel, set_insert_kernel,
                   threshold_lut)

__all__ = ["IBF", "count_batch_sharded", "enable_kmer_tables", "get_l2_fetch_granularity", "set_l2_fetch_granularity", "RBError", "build_library", "calculate_ci", "cut_out_nnns", "device_count",
           "fragment_schedule", "ibf_size_bits", "kernel_launches", "keys_decode", "lib", "lib_path",
           "host_pack_info", "transfer_bytes", "microbench_gather", "microbench_gather_coop", "set_count_kernel", "set_insert_kernel", "threshold_lut"]
