// ibf_kernels.cuh -- launch interfaces of the sm_100a kernels (internal to the library).
#pragma once

#include "ibf_common.cuh"

#include <vector>

namespace rb {

// View of one handle's slice of the bit matrix, as the kernels see it.
struct FilterView {
    const uint64_t *words;   // row r starts at words + r * stride
    uint64_t stride;         // local row width in 64-bit words (col_words)
    uint64_t bin_begin;      // global bin index of local bit 0
    uint64_t n_bins_local;   // valid local bins (the last word may be partial)
    HashParams hp;
};

struct CountArgs {
    FilterView fv;
    const uint8_t *bases;       // ASCII
    const uint64_t *read_off;   // [n_reads + 1]
    uint64_t n_reads;
    const uint16_t *lut;        // [n_lut][65536], already cast to uint16
    uint32_t n_lut;
    uint64_t *keys;             // [n_lut][n_reads]
    uint16_t *counts_fwd;       // [n_reads][n_bins_local] or null
    uint16_t *counts_rev;
    uint8_t *read_flag;         // [n_reads] or null
    // optional: the bases as three bit streams made by the host packer (rb_pack.hpp); bit i = base pk_base0 + i
    // of the `bases` coordinate system that read_off uses.  Only the group-per-read window-table kernel reads them.
    const uint32_t *pk_lo, *pk_hi, *pk_bad;
    uint64_t pk_base0;
    // keys_shared: `keys` is ONE array shared by the count kernels of all bin shards (rb_ibf_count_batch_sharded), possibly in
    // a peer device's memory: every kernel folds its key in with a system-scope 64-bit atomicMax instead of storing it, and
    // nobody but the owner zeroes it.  Honoured by the kernels wide shards use (slots, postings, streaming, multi-tile).
    int keys_shared;
};

struct InsertArgs {
    uint64_t *words;
    uint64_t stride;
    uint64_t bin_begin;      // first global bin held by this handle
    uint64_t bin_end;        // one past the last global bin held
    uint64_t n_bins;         // global bin count (bins >= this are an error)
    HashParams hp;
    const uint8_t *bases;
    const uint64_t *frag_begin, *frag_end, *frag_bin;
    uint64_t n_frags;
    unsigned int *error_flag;  // set to 1 when a fragment names a bin >= n_bins
};

// Grid size of the grid-stride count kernels in waves of resident CTAs.  One wave (every CTA resident from start to end, equal
// shares of the reads) leaves the SMs that finish early idle: memory channels are not equally far from every SM.  Several
// waves of smaller shares let the block scheduler hand out work as SMs become free: measured on BASELINE config #2's shape
// 2.38 -> 2.22 ms (k = 13), 6.90 -> 5.46 ms (k = 15), 8.34 -> 6.82 ms (k = 17) per batch with 16 waves
// (profiles/r2_d_grid_waves.jsonl).  Default 16; RB_GRID_WAVES overrides (measurements).
int grid_waves();

// which: 0 auto, 1 tile kernel, 2 streaming kernel.  Returns number of kernel launches or <0.
int launch_count(const CountArgs &a, uint32_t max_read_len, int which, int sm_count, cudaStream_t st);
// direct k-mer table (ibf_table.cu): build 4^k entries of 2*stride words; count through the table
int launch_table_build(const FilterView &fv, uint64_t *table, uint64_t n_entries, int span, int sm_count, cudaStream_t st);
int launch_count_table(const CountArgs &a, const uint64_t *table, int span, uint32_t max_read_len, int variant,
                       int sm_count, cudaStream_t st);
// one-k-mer table for rows of 3..16 words, entries padded to 4 / 8 / 16 words and loaded by that many lanes (ibf_ctable.cu)
int ctable_lanes(uint64_t stride);
int launch_ctable_build(const FilterView &fv, uint64_t *table, uint64_t n_kmers, int sm_count, cudaStream_t st);
int launch_count_ctable(const CountArgs &a, const uint64_t *table, uint32_t max_read_len, int variant, int sm_count, cudaStream_t st);
// window k-mer table with cooperative entry loads (ibf_wtable.cu): span = 2..4 k-mers per entry, rows <= 2 words
bool wtable_geometry(uint64_t stride, uint32_t k, int span, int *lanes, int *canon, uint64_t *n_entries);
int launch_wtable_build(const FilterView &fv, uint64_t *table, int span, int sm_count, cudaStream_t st);
int launch_count_wtable(const CountArgs &a, const uint64_t *table, int span, uint32_t max_read_len, int sm_count,
                        cudaStream_t st);
// k-mer postings table for wide filters (ibf_postings.cu)
bool postings_applicable(const FilterView &fv);
int postings_sample_units(const FilterView &fv, uint32_t *d_scratch, uint32_t n_sample, double *mean_units, int sm_count,
                          cudaStream_t st);
int postings_build_ptr(const FilterView &fv, uint32_t *d_ptr, uint64_t *total_units, int sm_count, cudaStream_t st);
int postings_fill(const FilterView &fv, const uint32_t *d_ptr, uint16_t *d_ids, int sm_count, cudaStream_t st);
int launch_count_postings(const CountArgs &a, const uint32_t *d_ptr, const uint16_t *d_ids, uint32_t max_read_len, double mean_units,
                          int sm_count, cudaStream_t st);
// the same lists as fixed, 128-byte-aligned slots per k-mer, fetched by bulk copies into shared-memory rings (ibf_postings.cu)
bool slots_applicable(const FilterView &fv);
int slots_sample_lengths(const FilterView &fv, uint32_t *d_scratch, uint32_t n_sample, std::vector<uint32_t> *lengths, int sm_count,
                         cudaStream_t st);
bool slots_choose(const std::vector<uint32_t> &lengths, uint32_t k, uint64_t budget, uint32_t *slot_bytes, uint64_t *ovf_units,
                  uint64_t *total_bytes);
int slots_fill(const FilterView &fv, uint8_t *d_slots, uint32_t slot_bytes, uint16_t *d_ovf, uint64_t ovf_units, unsigned int *d_ctr,
               int sm_count, cudaStream_t st);
int launch_count_slots(const CountArgs &a, const uint8_t *d_slots, uint32_t slot_bytes, const uint16_t *d_ovf, uint32_t max_read_len,
                       int sm_count, unsigned int *d_err, cudaStream_t st);
bool wgroup_applicable(uint32_t k, int span, uint32_t max_read_len);
int get_wtable_variant();
void set_wtable_variant(int v);   // 0 auto (group-per-read kernel for short reads), 1 always warp-per-read
// device memory the column build works in (fragment lists, scan storage, bit columns); owned by the caller and kept
// between calls so that only the first build of a handle allocates
struct ScratchBuf { void *p = nullptr; size_t cap = 0; };
int launch_insert(const InsertArgs &a, uint64_t max_frag_len, int sm_count, ScratchBuf *scr, cudaStream_t st);
int get_insert_variant();
void set_insert_variant(int v);   // 0 auto, 1 always 64-bit RED.OR per (k-mer, hash), 2 column build whenever the column fits
int launch_keys_decode(const uint64_t *keys, uint64_t n, uint16_t *max_count, uint8_t *hit,
                       uint32_t *argmax_bin, cudaStream_t st);
int launch_keys_decode_piece(const uint64_t *keys, const uint8_t *flag_in, uint64_t n, uint32_t n_lut, uint64_t stride,
                             uint64_t out_off, uint16_t *max_count, uint8_t *hit, uint32_t *argmax_bin, uint8_t *flag_out,
                             cudaStream_t st);

// measurement aid (ibf_traffic.cu): DRAM bytes at 128-byte line granularity, table accesses and base bytes of one count launch
int launch_traffic(const uint8_t *bases, const uint64_t *read_off, uint64_t n_reads, uint32_t k, uint32_t n_hash, int kind, int span,
                   uint32_t entry_bytes, const uint32_t *ptr, unsigned long long *d_out, int sm_count, cudaStream_t st);

}  // namespace rb
