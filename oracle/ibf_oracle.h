/*
 * ibf_oracle.h -- CPU oracle for the ReadBouncer IBF classify/build hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (readbouncer_b200/,
 * include/) may include, link or call this.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, and only as the
 * checker or the reported CPU baseline.
 *
 * It is a plain-C restatement of
 *   - the SeqAn-2 `BinningDirectory<InterleavedBloomFilter, BDConfig<Dna5,Normal,
 *     Uncompressed>>` engine that /root/reference/src/IBF wraps
 *     (reference: src/IBF/IBF.hpp:92-94; the engine itself is an un-vendored
 *     dependency: github.com/JensUweUlrich/seqan branch `SeqAn` + sdsl-lite
 *     v2.1.1, src/seqan/CMakeLists.txt.in:20-37), and
 *   - the first-party logic of src/IBF/IBFClassify.cpp, src/IBF/IBFBuild.cpp,
 *     src/IBF/IBF.hpp:268-338, src/main/classify.hpp:58-124 and
 *     src/main/adaptive_sampling.hpp:35-113.
 *
 * Parity is PINNED for k=13/15, h=3, binWidth=1 by the reference's own golden
 * fixtures (tests/golden/, made by tests/golden/make_golden.py from
 * src/test/libIBFTests/data and src/test/classifyTests/data): byte-identical
 * rebuild of all three .ibf files and the known answers 23 / 282 / 182 /
 * (5,30) / -7 / 79121216.  Corners no fixture reaches ("parity unpinned"):
 * binWidth > 1, k > 15, N/IUPAC/U in reads, text shorter than k, bin >= noOfBins.
 */
#ifndef IBF_ORACLE_H_
#define IBF_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_ibf {
    uint64_t n_bins;     /* noOfBins                                   */
    uint64_t n_hash;     /* noOfHashFunc                               */
    uint64_t k;          /* kmerSize                                   */
    uint64_t n_bits;     /* payload bits (without the 256 metadata)    */
    uint64_t bin_width;  /* ceil(n_bins / 64) 64-bit words per row     */
    uint64_t block_bits; /* 64 * bin_width                             */
    uint64_t n_blocks;   /* n_bits / block_bits  (rows)                */
    uint64_t n_words;    /* ceil((n_bits + 256) / 64)                  */
    uint64_t *words;     /* sdsl bit_vector payload incl. metadata tail */
} orc_ibf;

enum {
    ORC_OK = 0,
    ORC_ERR_NULL_FILTER = 1,   /* NullFilterException    IBFExceptions.hpp:178 */
    ORC_ERR_SHORT_READ = 2,    /* ShortReadException     IBFExceptions.hpp:96  */
    ORC_ERR_COUNT_KMER = 3,    /* CountKmerException     IBFExceptions.hpp:123 */
    ORC_ERR_PARSE_IBF = 4,     /* ParseIBFFileException  IBFExceptions.hpp:344 */
    ORC_ERR_MISSING_IBF = 5,   /* MissingIBFFileException IBFExceptions.hpp:317 */
    ORC_ERR_STORE = 6,         /* StoreFilterException   IBFExceptions.hpp:234 */
    ORC_ERR_INSERT = 7,        /* InsertSequenceException IBFExceptions.hpp:206 */
    ORC_ERR_CONFIG = 8,        /* InvalidConfigException IBFExceptions.hpp:150 */
    ORC_ERR_ALLOC = 9
};

/* ---- engine (SeqAn binning_directory restated) ------- */
uint8_t orc_dna5(char c);
orc_ibf *orc_ibf_create(uint64_t n_bins, uint64_t n_hash, uint64_t k, uint64_t n_bits);
orc_ibf *orc_ibf_load(const char *path, int *status);
int orc_ibf_store(const orc_ibf *f, const char *path);
int orc_ibf_resize_bins(orc_ibf *f, uint64_t new_n_bins);
void orc_ibf_free(orc_ibf *f);
uint64_t *orc_ibf_words(orc_ibf *f);
void orc_ibf_info(const orc_ibf *f, uint64_t *n_bins, uint64_t *n_hash, uint64_t *k,
                  uint64_t *n_bits, uint64_t *n_words);
uint64_t orc_kmer_hash(const char *text, uint64_t k);
uint64_t orc_hash_row(const orc_ibf *f, uint64_t kmer_value, unsigned i);
void orc_insert(orc_ibf *f, const char *text, uint64_t len, uint64_t bin);
void orc_count(const orc_ibf *f, const char *text, uint64_t len, int revcomp, uint16_t *counts);

/* ---- src/IBF first-party logic ------------------------------------------ */
uint64_t orc_filter_size_bits(uint64_t fragment_length, uint64_t k, uint64_t n_hash,
                              double max_fp, uint64_t n_bins);
void orc_calculate_ci(double r, uint8_t k, uint32_t readlen, double confidence,
                      uint16_t *low, uint16_t *high);
uint16_t orc_threshold(double r, uint64_t k, uint64_t readlen, double confidence);
void orc_threshold_lut(double r, uint64_t k, double confidence, uint16_t *lut65536);
uint64_t orc_cut_out_nnns(const char *seq, uint64_t len, char *out);
uint64_t orc_bins_for_sequence(uint64_t cut_len, uint64_t fragment_length);
uint64_t orc_fragment_schedule(uint64_t seqlen, uint64_t fragment_length, uint64_t k,
                               uint64_t *begin, uint64_t *end, uint64_t cap);
int orc_select_matches(const uint16_t *fwd, const uint16_t *rev, uint64_t n_bins, uint16_t thr);
uint64_t orc_max_matches(const uint16_t *fwd, const uint16_t *rev, uint64_t n_bins, uint16_t thr);
uint64_t orc_count_matches(const orc_ibf *f, const char *read, uint64_t len, double error_rate,
                           double significance);
int orc_classify_any(const orc_ibf *const *filters, uint64_t n_filters, const char *read,
                     uint64_t len, double error_rate, double significance, int *status);
int orc_classify_best(const orc_ibf *const *filters, uint64_t n_filters, const char *read,
                      uint64_t len, double error_rate, double significance, int *status);
int orc_classify_pair(const orc_ibf *const *filt1, uint64_t n1, const orc_ibf *const *filt2,
                      uint64_t n2, const char *read, uint64_t len, double error_rate,
                      double significance, uint64_t *first, uint64_t *second);
int orc_check_unblock(const orc_ibf *const *deplete, uint64_t n_dep, const orc_ibf *const *target,
                      uint64_t n_tgt, const char *read, uint64_t len, double error_rate,
                      double significance, int *status);

/* ---- batch drivers (CPU baseline; threads over reads / fragments) ------- */
/* Same outputs as the C-ABI rb_ibf_count_batch; any output pointer may be NULL. */
int orc_count_batch(const orc_ibf *f, const char *bases, const uint64_t *read_off,
                    uint64_t n_reads, const uint16_t *thr_lut, uint16_t *counts_fwd,
                    uint16_t *counts_rev, uint16_t *max_count, uint8_t *hit,
                    uint32_t *argmax_bin, uint8_t *short_read, int n_threads);
int orc_insert_batch(orc_ibf *f, const char *bases, const uint64_t *frag_begin,
                     const uint64_t *frag_end, const uint64_t *frag_bin, uint64_t n_frags,
                     int n_threads);

#ifdef __cplusplus
}
#endif
#endif
