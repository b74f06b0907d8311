// host_pack.hpp -- host side of the packed read transfer (internal to the library).
//
// The host-buffer classify call (rb_ibf_count_batch) is bound by PCIe when it ships ASCII bases
// (8 bits per base).  The group-per-read window-table kernel needs the reads as three bit planes anyway
// (low code bit, high code bit, "not ACGT"), so the host makes those planes -- 3 bits per base -- with
// AVX-512 or AVX2 (64 or 32 bases per iteration) on a small pool of threads, straight into pinned staging memory.
#pragma once

#include <cstddef>
#include <cstdint>
#include <functional>

namespace rb {

// Planes of bases[0, n): bit i of lo/hi/bad = base i.  Codes (ascii >> 1) & 3 = A0 C1 T2 G3, U/u = T;
// every other character is "bad" (Dna5 rank 4) with code bits 0.  n need not be a multiple of 32: the last
// word is zero padded.  Writes ceil(n / 32) words per plane.
// streaming_stores: write the planes with non-temporal stores (staging memory that a DMA engine reads next).
void pack_bases(const uint8_t *bases, size_t n, uint32_t *lo, uint32_t *hi, uint32_t *bad, bool streaming_stores = false);

// The same, but the "not ACGT" plane is only written if the range has such a base (returns true then: all ceil(n / 32) words
// of `bad` are valid).  When it returns false `bad` may be untouched: most reads have no such base, and a plane that is
// neither written by the host threads nor read by the DMA engine is an eighth less host-memory traffic per base.
bool pack_bases_lazy(const uint8_t *bases, size_t n, uint32_t *lo, uint32_t *hi, uint32_t *bad, bool streaming_stores = false);

// max over i < n of off[i+1] - off[i] as unsigned 64-bit (a decreasing pair shows up as a value >= 2^63)
uint64_t max_read_length(const uint64_t *off, size_t n);

// Instruction set of the packer in use: 0 scalar, 2 AVX2, 5 AVX-512 BW+VBMI (reported by rb_host_pack_info;
// env RB_HOST_PACK_ISA caps it).
int pack_isa();

// Runs fn(task) for task in [0, n_tasks) on the library's host threads plus the caller; `poll` is called by the
// CALLER between its own tasks and while waiting (it submits finished pieces to the GPU) and receives the number of
// leading tasks known complete.  Falls back to the caller alone when the pool is busy with another call.
void parallel_tasks(size_t n_tasks, const std::function<void(size_t)> &fn, const std::function<void(size_t)> &poll);

int host_threads();   // pool size + 1 (the caller)

// Measurement aid: the pool's threads stream-read buf[0, n) once (the packer's access pattern without its work or its
// stores); returns a checksum so the loads stay.  Tells how much of the packer's time is the host memory system.
uint64_t stream_read(const uint8_t *buf, size_t n);

}  // namespace rb
