/*
 * rb_ibf.h -- C ABI of the B200-native Interleaved Bloom Filter (IBF) engine.
 *
 * This is the drop-in boundary for ReadBouncer's IBF hot path.  The reference
 * has no FFI layer: src/IBF (interleave::IBF, interleave::Read) calls seven
 * entry points of an external SeqAn-2 `BinningDirectory<InterleavedBloomFilter>`
 * directly.  Each function below names the reference interface it replaces
 * (paths relative to the ReadBouncer tree).  The C++ shim that re-creates the
 * interleave:: classes on top of this ABI is include/rb_interleave.hpp; the
 * binding a ReadBouncer maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *  - plain pointers and sizes only; no exceptions cross the boundary; every
 *    call returns an rb_status (0 = ok) or reports one through `status`.
 *  - rb_last_error() returns a thread-local message for the last failure.
 *  - "host" entry points take host buffers and include the H2D/D2H copies;
 *    "_dev" entry points take device pointers (same device as the handle) and
 *    only enqueue work on `stream` (a cudaStream_t passed as void*; NULL = the
 *    legacy default stream).  Nothing here falls back to the CPU: without a
 *    CUDA device the calls fail with RB_ERR_NO_DEVICE.
 *  - handles are immutable after load/insert, so count calls on one handle may
 *    run concurrently from several host threads on different streams.
 *  - reads are ASCII (A/C/G/T/U any case, everything else = N), concatenated in
 *    `bases`; read i is bases[read_off[i] .. read_off[i+1]).
 */
#ifndef RB_IBF_H_
#define RB_IBF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define RB_API __declspec(dllexport)
#else
#define RB_API __attribute__((visibility("default")))
#endif

typedef struct rb_ibf rb_ibf;      /* opaque device-resident filter            */
typedef void *rb_stream;           /* cudaStream_t                              */

/* Status codes; 1..8 map 1:1 onto the reference's exception classes
 * (src/IBF/IBFExceptions.hpp). */
typedef enum rb_status {
    RB_OK = 0,
    RB_ERR_NULL_FILTER = 1,       /* NullFilterException      :178 */
    RB_ERR_SHORT_READ = 2,        /* ShortReadException       :96  */
    RB_ERR_COUNT_KMER = 3,        /* CountKmerException       :123 */
    RB_ERR_PARSE_IBF_FILE = 4,    /* ParseIBFFileException    :344 */
    RB_ERR_MISSING_IBF_FILE = 5,  /* MissingIBFFileException  :317 */
    RB_ERR_STORE_FILTER = 6,      /* StoreFilterException     :234 */
    RB_ERR_INSERT_SEQUENCE = 7,   /* InsertSequenceException  :206 */
    RB_ERR_INVALID_CONFIG = 8,    /* InvalidConfigException   :150 */
    RB_ERR_ALLOC = 9,
    RB_ERR_CUDA = 10,
    RB_ERR_NO_DEVICE = 11,
    RB_ERR_INVALID_ARG = 12
} rb_status;

RB_API const char *rb_status_string(int status);
RB_API const char *rb_last_error(void);
RB_API int rb_device_count(void);

/* ---- filter metadata ----------------------------------------------------- */
typedef struct rb_ibf_info_t {
    uint64_t n_bins;        /* noOfBins (global)                                */
    uint64_t n_hash;        /* noOfHashFunc                                     */
    uint64_t kmer_size;     /* kmerSize                                         */
    uint64_t n_bits;        /* payload bits (file bit length - 256)             */
    uint64_t bin_width;     /* 64-bit words per row (global) = ceil(bins/64)    */
    uint64_t n_blocks;      /* rows = n_bits / (64*bin_width)                   */
    uint64_t col_begin;     /* first row word held by this handle (bin shard)   */
    uint64_t col_words;     /* row words held by this handle                    */
    uint64_t bin_begin;     /* first global bin held  = 64*col_begin            */
    uint64_t n_bins_local;  /* bins held by this handle                         */
    uint64_t device_bytes;  /* bytes of HBM used by the bit matrix              */
    int32_t device;
    int32_t shard, n_shards;
    int32_t kmer_table_span;   /* consecutive k-mers per table entry (1..4), 0 if not built */
    uint64_t kmer_table_bytes; /* bytes of the k-mer table, 0 if not built */
    int32_t kmer_table_kind;   /* 0 none, 1 dense k-mer / window table (rows <= 2 words; 3-4 words when kind 4 does not fit),
                                  4 one k-mer per entry, rows of 3..32 words padded to 4/8/16/32 and loaded by as many lanes,
                                  wider rows: 2 postings as pointer + lists (default), 3 postings in fixed slots per k-mer
                                  (RB_POSTINGS_LAYOUT=slots) */
} rb_ibf_info_t;

/* ---- host-side scalar helpers (FP64, bit-exact with the reference) ------- */
/* IBF::calculate_filter_size_bits, src/IBF/IBFBuild.cpp:404-413 */
RB_API uint64_t rb_ibf_size_bits(uint64_t fragment_length, uint32_t kmer_size, uint32_t n_hash,
                                 double max_fp, uint64_t n_bins);
/* interleave::calculateCI, src/IBF/IBF.hpp:320-338 */
RB_API int rb_calculate_ci(double error_rate, uint32_t kmer_size, uint32_t readlen, double significance,
                           uint16_t *low, uint16_t *high);
/* threshold of Read::find_matches / count_matches incl. the int16 -> uint16 wrap,
 * src/IBF/IBFClassify.cpp:102-109,154-159; lut[len] for len in [0, 65536) */
RB_API int rb_threshold_lut(double error_rate, double significance, uint32_t kmer_size,
                            uint16_t *lut65536);
/* IBF::cutOutNNNs + concatenation, src/IBF/IBFBuild.cpp:81-88,112-132.
 * `out` needs `len` bytes; returns the new length. */
RB_API uint64_t rb_cut_out_nnns(const char *seq, uint64_t len, char *out);
/* fragment loop of add_sequences_to_filter, src/IBF/IBFBuild.cpp:165-202.
 * Returns the number of fragments (= bin ids consumed); fills at most `cap`. */
RB_API uint64_t rb_fragment_schedule(uint64_t seqlen, uint64_t fragment_length, uint32_t kmer_size,
                                     uint64_t *begin, uint64_t *end, uint64_t cap);

/* ---- filter life cycle ---------------------------------------------------- */
/* TIbf(bins, hashes, k, bits) ctor, src/IBF/IBFBuild.cpp:465: zero-filled matrix in HBM */
RB_API rb_ibf *rb_ibf_create(uint64_t n_bins, uint32_t n_hash, uint32_t kmer_size, uint64_t n_bits,
                             int device, int *status);
/* The same ctor for ONE bin shard of a filter that is never held in one place (BASELINE config #5): a zero-filled
 * column slice [shard*W/n_shards, (shard+1)*W/n_shards) of every row of the (n_bins, n_bits) filter.  Inserts take
 * global bin ids and skip the bins of other shards. */
RB_API rb_ibf *rb_ibf_create_shard(uint64_t n_bins, uint32_t n_hash, uint32_t kmer_size, uint64_t n_bits,
                                   int device, int shard, int n_shards, int *status);
/* seqan::retrieve, src/IBF/IBFBuild.cpp:343,360 and src/config/configReader.cpp:216.
 * Validates the sdsl header and metadata tail; a FASTA or truncated file fails
 * with RB_ERR_PARSE_IBF_FILE, a missing one with RB_ERR_MISSING_IBF_FILE. */
RB_API rb_ibf *rb_ibf_load(const char *path, int device, int *status);
/* Bin-sharded load: this handle keeps only row words
 * [shard*W/n_shards, (shard+1)*W/n_shards) of every row (W = bin_width). */
RB_API rb_ibf *rb_ibf_load_shard(const char *path, int device, int shard, int n_shards, int *status);
/* Upload from host memory: `words` = the n_bits/64 payload words in file order. */
RB_API rb_ibf *rb_ibf_from_words(const uint64_t *words, uint64_t n_bins, uint32_t n_hash,
                                 uint32_t kmer_size, uint64_t n_bits, int device, int shard,
                                 int n_shards, int *status);
/* seqan::store, src/IBF/IBFBuild.cpp:307,505: byte-identical sdsl bit_vector file (unsharded handles) */
RB_API int rb_ibf_store(const rb_ibf *f, const char *path);
/* D2H copy of the local payload words (n_blocks*col_words for a shard, n_bits/64 otherwise) */
RB_API int rb_ibf_download(const rb_ibf *f, uint64_t *words, uint64_t n_words);
RB_API void rb_ibf_free(rb_ibf *f);
/* seqan::getNumberOfBins / getKmerSize, src/IBF/IBFBuild.cpp:380-381,466 */
RB_API int rb_ibf_info(const rb_ibf *f, rb_ibf_info_t *out);
/* Raw device pointer of the bit matrix (for zero-copy interop, e.g. torch.from_blob) */
RB_API uint64_t *rb_ibf_device_words(const rb_ibf *f);
/* Device pointer of the k-mer table (NULL if not built); measurement aid for the gather microbenchmarks. */
RB_API const uint64_t *rb_ibf_device_kmer_table(const rb_ibf *f);

/* filter.resizeBins(n) as used by IBF::update_filter (src/IBF/IBFBuild.cpp:269-279): grow the bin count; the
 * number of rows is kept and rows are widened when n crosses a multiple of 64 (n_bits = rows * 64 * ceil(n/64)).
 * Unsharded handles only.  (SeqAn's resizeBins is not pinned by any reference fixture.) */
RB_API int rb_ibf_resize_bins(rb_ibf *f, uint64_t new_n_bins, rb_stream stream);

/* ---- build: seqan::insertKmer(filter, fragment, bin), src/IBF/IBFBuild.cpp:189-190 ---- */
/* For every fragment i: every k-mer of bases[frag_begin[i] .. frag_end[i]) sets bit
 * (row(h_j(kmer)), frag_bin[i]) for all hash functions j.  Fragments shorter than k
 * insert nothing; a bin outside this handle's bin range is skipped (bin shards) and a
 * bin >= n_bins is reported as RB_ERR_INSERT_SEQUENCE. */
RB_API int rb_ibf_insert_batch(rb_ibf *f, const char *bases, uint64_t n_bases,
                               const uint64_t *frag_begin, const uint64_t *frag_end,
                               const uint64_t *frag_bin, uint64_t n_frags, rb_stream stream);
/* device-pointer variant; max_frag_len is a launch-shape hint (0 = unknown).
 * Device base buffers (here and in rb_ibf_count_batch_dev) are fetched as aligned 16-byte blocks: they must be readable from
 * the 16-byte boundary at or below their first byte to the one at or above their last byte -- true of every cudaMalloc /
 * torch allocation and of any sub-range of one; the bytes outside the fragments / reads are never interpreted. */
RB_API int rb_ibf_insert_batch_dev(rb_ibf *f, const uint8_t *d_bases, const uint64_t *d_frag_begin,
                                   const uint64_t *d_frag_end, const uint64_t *d_frag_bin,
                                   uint64_t n_frags, uint64_t max_frag_len, rb_stream stream);

/* ---- classify: seqan::count x2 + threshold + select/max_matches -------------------------
 * Replaces, per read, Read::count_matches / find_matches (src/IBF/IBFClassify.cpp:81-171):
 *   counts_fwd = seqan::count(filter, read), counts_rev = seqan::count(filter, revcomp(read)),
 *   thr = lut[len], hit = select_matches(...), max_count = max_matches(...).
 * n_lut threshold tables (1 to 4; each uint16[65536]) are evaluated in the same pass so the
 * error_rate-0.02 retry of check_unblock / classify_deplete_target needs no second launch.
 *
 * Outputs (any may be NULL): counts_* [n_reads][n_bins_local] dense per-bin counts;
 * max_count/hit/argmax_bin [n_lut][n_reads]; argmax_bin = lowest global bin attaining
 * max_count among bins passing the threshold, 0xFFFFFFFF if none;
 * read_flag [n_reads]: 0 ok, 1 = shorter than k (ShortReadException), 2 = longer than 65535
 * bases (not representable in the reference's uint16 readlen; not classified), 3 = longer than the
 * max_read_len given to rb_ibf_count_batch_dev (not classified; never set by rb_ibf_count_batch). */
RB_API int rb_ibf_count_batch(const rb_ibf *f, const char *bases, const uint64_t *read_off,
                              uint64_t n_reads, const uint16_t *thr_lut, uint32_t n_lut,
                              uint16_t *counts_fwd, uint16_t *counts_rev, uint16_t *max_count,
                              uint8_t *hit, uint32_t *argmax_bin, uint8_t *read_flag,
                              rb_stream stream);

/* Host side of rb_ibf_count_batch: number of host threads that pack reads into bit planes for the
 * PCIe transfer (env RB_HOST_THREADS, default min(cores, 16); RB_HOST_PACK=0 ships ASCII instead) and
 * the packer's instruction set (0 scalar, 2 AVX2, 5 AVX-512 BW+VBMI; env RB_HOST_PACK_ISA caps it). */
RB_API int rb_host_pack_info(int *threads, int *isa);
/* Measurement aid: the packer's host threads stream-read n_bytes of `buf` `reps` times (its access pattern without its
 * work or its stores); GB/s of host memory the packer can see.  bench.py reports it next to the end-to-end rate. */
RB_API int rb_microbench_host_read(const void *buf, uint64_t n_bytes, uint32_t reps, double *gb_per_s);
/* Bytes rb_ibf_count_batch has moved over PCIe since the library was loaded (all threads): host->device
 * copies (bit planes or ASCII bases, offsets, thresholds) and device->host results (copies or mapped stores). */
RB_API int rb_transfer_bytes(uint64_t *h2d, uint64_t *d2h);
/* How rb_ibf_count_batch ships large batches (>= 8 pieces of ~8 MB) of this filter: the library times its 2nd packed and
 * its 2nd ASCII call and keeps the faster way (ASCII only if >= 5 % faster); smaller batches are packed.  choice: -1 not
 * decided yet, 0 packed bit planes, 1 ASCII; the two measured times in ns per base (0 = not measured).  RB_HOST_PACK=1 / 0
 * pins the choice and skips the measurement. */
RB_API int rb_ibf_transfer_policy(const rb_ibf *f, int *choice, double *ns_per_base_packed, double *ns_per_base_ascii);

/* Packed per-read summary key used by the device API and the bin-sharded combine:
 *   bit 48      hit (some bin passes the threshold)
 *   bits 47..32 max_count
 *   bits 31..0  ~argmax_bin   (so that a 64-bit MAX picks the lowest bin on ties)
 * key == 0 means no bin passed; keys are < 2^49, so signed and unsigned 64-bit MAX agree.  The
 * elementwise MAX of the keys of all bin shards is the key of the whole filter, which is what the
 * NCCL combine of the bin-sharded mode reduces. */
#define RB_KEY_HIT(key) ((uint8_t)(((uint64_t)(key) >> 48) & 1u))
#define RB_KEY_MAX_COUNT(key) ((uint16_t)(((uint64_t)(key) >> 32) & 0xFFFFu))
#define RB_KEY_ARGMAX_BIN(key) ((key) ? ~(uint32_t)((uint64_t)(key) & 0xFFFFFFFFu) : 0xFFFFFFFFu)

/* Device-pointer variant.  d_keys [n_lut][n_reads] (required) receives the packed summaries;
 * d_counts_* and d_read_flag may be NULL.  max_read_len (>= every read length, 0 = assume
 * 65535) only selects kernel variants (narrower counters for short reads).  A read LONGER than a non-zero max_read_len
 * promised is never miscounted: the kernels that rely on the promise give it read_flag 3 and key 0 (not classified);
 * rb_ibf_count_batch computes the maximum itself, so flag 3 cannot occur there.  Work is enqueued on `stream`; no host sync. */
RB_API int rb_ibf_count_batch_dev(const rb_ibf *f, const uint8_t *d_bases, const uint64_t *d_read_off,
                                  uint64_t n_reads, uint32_t max_read_len, const uint16_t *d_thr_lut,
                                  uint32_t n_lut, uint64_t *d_keys, uint16_t *d_counts_fwd,
                                  uint16_t *d_counts_rev, uint8_t *d_read_flag, rb_stream stream);
/* Unpack keys on the device into max_count / hit / argmax_bin arrays (any may be NULL). */
RB_API int rb_keys_decode_dev(const uint64_t *d_keys, uint64_t n, uint16_t *d_max_count, uint8_t *d_hit,
                              uint32_t *d_argmax_bin, int device, rb_stream stream);

/* ---- bin-sharded filters (BASELINE config #5: a database larger than one GPU's HBM) --------------------------------
 * Every shard holds a column slice of every row (rb_ibf_load_shard / rb_ibf_create_shard), every shard classifies ALL
 * reads against its bins, and the per-read keys of the shards combine by elementwise MAX (RB_KEY_* above).
 *
 * One host process driving several devices (what a C++ host like ReadBouncer does):
 * shards[i] lives on its own device; the batch is copied to every device, every device counts, and every count kernel
 * folds its key straight into ONE key array on shards[0]'s device with a system-scope 64-bit atomicMax over NVLink
 * peer memory -- the combine is the kernels' own epilogue, there is no separate collective and no host hop; shard 0
 * decodes.  Outputs as rb_ibf_count_batch (max_count / hit / argmax_bin [n_lut][n_reads], global bin ids; read_flag
 * [n_reads]); dense counts are per shard and not offered here.  Synchronous. */
RB_API int rb_ibf_count_batch_sharded(const rb_ibf *const *shards, uint32_t n_shards, const char *bases,
                                      const uint64_t *read_off, uint64_t n_reads, const uint16_t *thr_lut, uint32_t n_lut,
                                      uint16_t *max_count, uint8_t *hit, uint32_t *argmax_bin, uint8_t *read_flag);
/* One process per GPU (MPI / torchrun style): after rb_ibf_count_batch_dev on every rank, ONE in-place
 * ncclAllReduce(ncclUint64, ncclMax) of the n keys on `stream` (8 bytes per read and threshold table; dense counts are
 * never exchanged), then rb_keys_decode_dev.  nccl_comm is the caller's ncclComm_t; libnccl.so.2 is looked up at run
 * time (the copy already loaded in the process, if any), so the library itself does not link NCCL. */
RB_API int rb_keys_combine_nccl(void *nccl_comm, uint64_t *d_keys, uint64_t n, rb_stream stream);

/* Direct k-mer table for narrow filters (row <= 2 words, k <= 16): the AND of the h probed rows is a
 * pure function of the k-mer, so it is tabulated once for all ACGT k-mers and both strands.  An entry
 * covers a window of `span` consecutive k-mers (k+span-1 bases).  span 1: 4^k entries of 16*col_words
 * bytes (2.1 GB for k=13, 100 bins), read by one lane each.  span 2..4 (rows <= 2 words): entries of
 * 2 (span 2) or 4 slots of 16*col_words bytes, read by adjacent lanes in one request; when the window
 * length is odd only the windows whose middle base is A or C are stored (a window and its reverse
 * complement hold the same masks), e.g. k=13, span 3: 4^15/2 entries of 128 bytes = 64 GiB for 100 bins.
 * Count calls then cost one memory request per `span` k-mer positions instead of 2*h row probes per
 * position; windows containing N take the hashed path, so results are bit-identical.  Built
 * automatically by the first count call with >= 1024 reads, choosing the widest span that fits
 * min(60 % of the free HBM, 80 GiB) (env RB_KMER_TABLE=0 disables, RB_KMER_TABLE_MAX_GB changes the
 * cap, RB_KMER_TABLE_SPAN the widest span tried); dropped by rb_ibf_insert_batch*.  This call
 * (re)builds it now under the given byte budget (0 = automatic); UINT64_MAX disables the table for
 * this handle.  With the automatic budget an explicit call may take up to 85 % of the free HBM when only the one-k-mer
 * table is possible (k = 16).
 * MEDIUM filters (rows of 3..32 words = 129..2048 bins, k <= 16) get the same one-k-mer table with the row padded to
 * 4 / 8 / 16 / 32 words: entries of 64..512 bytes, each loaded by 4..32 adjacent lanes in one instruction and counted in
 * bit-sliced registers (ibf_ctable.cu; 8.6 GB for 303 bins, 17 GB for 1010 bins at k = 13; rows of 17..32 words only when
 * their postings lists would average more than 4 units, RB_CTABLE_WIDE=1/0 forces that choice).  RB_CTABLE=0 turns this layout
 * off (rows of 3-4 words then use the unpadded lane-per-entry table, wider ones the postings below).
 * Wide filters (rows > 32 words, or > 4 words when that table does not fit; <= 65280 local bins, k <= 15) get a POSTINGS table instead: the AND of
 * the probed rows is ~1 % dense by the reference's own sizing, so the list of set bins of every
 * k-mer (2 bytes each, ~50 GB for a human-genome filter at k=13) replaces streaming 2*h rows of
 * thousands of bytes per position; same policy, budget and env switches.  Two layouts: pointer + lists
 * of 16-byte units read straight into registers (default; lists averaging up to 16 units are walked by groups of 2 / 4 / 8
 * lanes, longer ones by the whole warp -- RB_POSTINGS_SUB=0/2/4/8 forces one), or a fixed,
 * 128-byte-aligned slot per k-mer (size chosen from the sampled list lengths; the few longer lists go to an
 * overflow area): slots of one or two lines are loaded by groups of 8 / 16 lanes and are chosen AUTOMATICALLY for lists
 * averaging 2.6..14 units (2 000..8 000 bins); larger slots (RB_POSTINGS_LAYOUT=slots only) are fetched by one bulk copy
 * into a shared-memory ring.  RB_POSTINGS_LAYOUT=lists / slots forces a layout.
 *
 * SUPPORTED ENVELOPE of the table paths (outside it results are the same, from the hashed / streaming kernels):
 *   rows <= 2 words (<= 128 bins): window tables for k + span - 1 <= 16, i.e. span 3 up to k = 14, span 2 up to k = 15,
 *                                   span 1 up to k = 16; k >= 17: hashed probes (count_tile_kernel)
 *                                   (k = 16, 65-128 bins: 137 GB, built on an explicit request / by the joint planner only)
 *   rows of 3-32 words:            padded one-k-mer table while 4^k * 16 * {4,8,16,32} bytes fit (k = 13: 4.3 .. 34 GB;
 *                                   k = 15 up to 8 words); else the unpadded table (3-4 words) / postings (5+ words)
 *   rows > 32 words:               postings up to k = 15 while 4^k slots fit the HBM budget (k = 13 for a human-sized
 *                                   filter); else count_stream_kernel
 * bench.py's `secondary` reports k = 13, 15 and 17 on the BASELINE config #2 shape so the steps are visible. */
RB_API int rb_ibf_enable_kmer_table(rb_ibf *f, uint64_t max_table_bytes, rb_stream stream);

/* The same for ALL filters a caller classifies against (the reference holds every target and depletion filter at
 * once: classify.hpp:142, adaptive_sampling.hpp:555): one joint plan per device under total_bytes (0 = 85 % of the
 * free HBM, counting what these handles' old tables held as free).  Every filter first gets its smallest table
 * (span 1 / postings), then the filter with the narrowest span is widened while the sum fits; a filter whose table
 * does not fit gets none and is marked so that no count call tries to build one later.  Call it once after loading /
 * building the filters (rb_drivers.hpp and rb_live.hpp do): no classify call then stalls on a table build. */
RB_API int rb_ibf_enable_kmer_tables(rb_ibf *const *filters, uint32_t n_filters, uint64_t total_bytes, rb_stream stream);

/* Threshold table WITHOUT the range check of rb_threshold_lut: whatever the reference's FP64 arithmetic gives for this
 * rate.  check_unblock / classify_deplete_target retry at error_rate - 0.02 without validating it
 * (adaptive_sampling.hpp:55-59, classify.hpp:75-80); for rates <= 0 the interval bound is NaN, which the uint16 cast
 * turns into 0 on x86-64, i.e. threshold = number of k-mers (every k-mer must match). */
RB_API int rb_threshold_lut_raw(double error_rate, double significance, uint32_t kmer_size, uint16_t *lut65536);

/* Kernel selection override for tests/benchmarks: 0 auto, 1 warp-per-read tile kernel,
 * 2 CTA-per-read streaming kernel, 3 direct k-mer table kernel (fails if not applicable; bit-sliced
 * register counters for rows <= 2 words), 4 k-mer table kernel with shared-memory counters,
 * 5 k-mer table, window tables always through the warp-per-read kernel (auto uses the
 * group-per-read kernel when every read of the launch has <= 127*span positions). */
RB_API int rb_set_count_kernel(int which);
/* Build kernel selection for tests/benchmarks: 0 auto, 1 one 64-bit RED.OR per (k-mer, hash) straight
 * into the interleaved matrix, 2 column build (a bin's bit column is filled in shared memory, then
 * 32x32 bit tiles are transposed into the matrix) whenever noOfBlocks bits fit in shared memory
 * (<= ~1.7 M rows; the reference's default fragment_size = 100 000 gives 1 236 269).  Auto takes the
 * column build when it fits and the call has at least one fragment and one bin per SM. */
RB_API int rb_set_insert_kernel(int which);
/* Number of kernels this library launched since load (all threads); evidence for gpu_launches. */
RB_API uint64_t rb_kernel_launches(void);

/* L2 fetch granularity of the device (cudaLimitMaxL2FetchGranularity: 32, 64 or 128 bytes).  Random
 * probes of narrow rows touch one 32-byte sector each; with the default granularity the L2 pulls
 * whole 128-byte lines from HBM and DRAM traffic is ~4x the useful bytes.  Device-wide setting. */
RB_API int rb_set_l2_fetch_granularity(int device, uint32_t bytes);
RB_API int rb_get_l2_fetch_granularity(int device, uint32_t *bytes);

/* Measurement aid (not on the product path): uniformly random row-aligned loads of row_bytes
 * (8, 16 or 32) over d_buf[n_rows*row_bytes]; n_blocks CTAs of 256 threads, probes_per_thread
 * loads each (8 in flight).  Establishes the random-sector roofline of SURVEY.md section 8d. */
RB_API int rb_microbench_gather(const void *d_buf, uint64_t n_rows, uint32_t row_bytes,
                                uint64_t probes_per_thread, uint32_t n_blocks, uint64_t *d_sink,
                                rb_stream stream);
/* Same, but row_bytes / lane_bytes adjacent lanes read one row (16..256 bytes) with ONE load
 * instruction of lane_bytes (16 or 32) per lane: the ceiling for warp-cooperative entry loads. */
RB_API int rb_microbench_gather_coop(const void *d_buf, uint64_t n_rows, uint32_t row_bytes, uint32_t lane_bytes,
                                     uint64_t probes_per_group, uint32_t n_blocks, uint64_t *d_sink,
                                     rb_stream stream);

/* Measurement aid: what one rb_ibf_count_batch_dev launch over this batch has to fetch from HBM, derived from the table
 * geometry of the handle as it is now (k-mer window table, postings, or hashed row probes when no table is built):
 *   table_bytes     bytes of table data at 128-byte line granularity (HBM delivers whole lines; nothing assumed in L2)
 *   table_requests  table accesses (entries, lists + list bounds, or row probes)
 *   io_bytes        read bases + offsets in, keys out
 * bench.py divides their sum by the kernel's measured time for roofline.frac; profiles/ holds the ncu dram__bytes check.
 * Synchronises the stream. */
RB_API int rb_ibf_count_traffic_dev(const rb_ibf *f, const uint8_t *d_bases, const uint64_t *d_read_off, uint64_t n_reads,
                                    uint32_t n_lut, uint64_t *table_bytes, uint64_t *table_requests, uint64_t *io_bytes,
                                    rb_stream stream);

/* Test / benchmark aid (not on the product path): n synthetic ACGT bases, positions [start, start + n) of the stream
 * `seed`, written to d_out (device).  base(i) = "ACGT"[(mix64(seed * 0xD1342543DE82EF95 + (i >> 5)) >> 2 (i & 31)) & 3]
 * with the splitmix64 finaliser, so a host regenerates any window without holding the sequence: multi-Gb references
 * (BASELINE configs #3-#5) are generated where they are inserted. */
RB_API int rb_synth_bases_dev(uint8_t *d_out, uint64_t n, uint64_t seed, uint64_t start, rb_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* RB_IBF_H_ */
