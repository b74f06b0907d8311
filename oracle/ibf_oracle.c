/*
 * ibf_oracle.c -- CPU oracle (TEST INFRASTRUCTURE, see ibf_oracle.h).
 *
 * Every function cites the reference location it restates (paths relative to
 * /root/reference).  The SeqAn engine itself is not in the reference tree
 * (src/seqan/CMakeLists.txt.in:28-37 clones JensUweUlrich/seqan@SeqAn at
 * configure time); its published algorithm is restated from SURVEY.md
 * Appendix A and pinned by the golden fixtures in tests/golden/.
 */
#define _GNU_SOURCE
#include "ibf_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* SeqAn IBF hashing constants. */
#define ORC_SEED 0x90b45d39fb6da1faULL
#define ORC_SHIFT 27

/* ------------------------------------------------------------------------ */
/* Alphabet: seqan::Dna5 conversion from char, as applied by the casts at
 * src/main/classify.hpp:274 `(seqan::Dna5String) fragment` and
 * src/IBF/IBFBuild.cpp:88 `((seqan::Dna5String) newseq)`.
 * A/a 0, C/c 1, G/g 2, T/t 3, U/u 3 (SeqAn-2 translate table), else 4 (N). */
uint8_t orc_dna5(char c)
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': case 'U': case 'u': return 3;
    default: return 4;
    }
}

/* complement on Dna5 ranks (seqan::ModComplementDna, src/IBF/IBF.hpp:96-97): N stays N */
static inline uint8_t comp5(uint8_t d) { return d < 4 ? (uint8_t)(3 - d) : 4; }

static void derive(orc_ibf *f)
{
    f->bin_width = (f->n_bins + 63) / 64;
    f->block_bits = 64 * f->bin_width;
    f->n_blocks = f->block_bits ? f->n_bits / f->block_bits : 0;
    f->n_words = (f->n_bits + 256 + 63) / 64;
}

static void write_meta(orc_ibf *f)
{
    /* metadata tail: bits n_bits .. n_bits+255 = [bins, hashes, k, k] (Appendix A.4).
     * n_bits is a multiple of 64 for every filter the reference can build
     * (BinSizeBits * optBins with optBins % 64 == 0, src/IBF/IBFBuild.cpp:407-412). */
    uint64_t w = f->n_bits / 64;
    f->words[w + 0] = f->n_bins;
    f->words[w + 1] = f->n_hash;
    f->words[w + 2] = f->k;
    f->words[w + 3] = f->k;
}

/* TIbf(bins, hashes, k, bits) -- src/IBF/IBFBuild.cpp:465 */
orc_ibf *orc_ibf_create(uint64_t n_bins, uint64_t n_hash, uint64_t k, uint64_t n_bits)
{
    if (n_bins == 0 || n_hash == 0 || n_hash > 16 || k == 0 || k > 32 || (n_bits % 64) != 0)
        return NULL;
    orc_ibf *f = (orc_ibf *)calloc(1, sizeof(orc_ibf));
    if (!f) return NULL;
    f->n_bins = n_bins; f->n_hash = n_hash; f->k = k; f->n_bits = n_bits;
    derive(f);
    if (f->n_blocks == 0) { free(f); return NULL; }
    f->words = (uint64_t *)calloc(f->n_words, sizeof(uint64_t));
    if (!f->words) { free(f); return NULL; }
    write_meta(f);
    return f;
}

/* seqan::retrieve -- src/IBF/IBFBuild.cpp:343,360; sniffing at src/config/configReader.cpp:210-224.
 * File = sdsl bit_vector::serialize: u64 bit length, then ceil(len/64) u64 words. */
orc_ibf *orc_ibf_load(const char *path, int *status)
{
    int st = ORC_OK;
    orc_ibf *f = NULL;
    FILE *fp = fopen(path, "rb");
    if (!fp) { st = ORC_ERR_MISSING_IBF; goto out; }
    uint64_t bit_len = 0;
    if (fseek(fp, 0, SEEK_END) != 0) { st = ORC_ERR_PARSE_IBF; goto out; }
    long fsize = ftell(fp);
    rewind(fp);
    if (fsize < 8 + 32 || fread(&bit_len, 8, 1, fp) != 1) { st = ORC_ERR_PARSE_IBF; goto out; }
    uint64_t n_words = (bit_len + 63) / 64;
    if (bit_len < 256 + 64 || (bit_len % 64) != 0 || (uint64_t)fsize != 8 + 8 * n_words) {
        st = ORC_ERR_PARSE_IBF; goto out;
    }
    f = (orc_ibf *)calloc(1, sizeof(orc_ibf));
    if (!f) { st = ORC_ERR_ALLOC; goto out; }
    f->words = (uint64_t *)malloc(n_words * 8);
    if (!f->words) { st = ORC_ERR_ALLOC; goto out; }
    if (fread(f->words, 8, n_words, fp) != n_words) { st = ORC_ERR_PARSE_IBF; goto out; }
    f->n_bits = bit_len - 256;
    uint64_t w = f->n_bits / 64;
    f->n_bins = f->words[w]; f->n_hash = f->words[w + 1]; f->k = f->words[w + 2];
    if (f->n_bins == 0 || f->n_hash == 0 || f->n_hash > 16 || f->k == 0 || f->k > 32 ||
        f->words[w + 3] != f->k) { st = ORC_ERR_PARSE_IBF; goto out; }
    derive(f);
    if (f->n_blocks == 0) { st = ORC_ERR_PARSE_IBF; goto out; }
out:
    if (fp) fclose(fp);
    if (st != ORC_OK && f) { free(f->words); free(f); f = NULL; }
    if (status) *status = st;
    return f;
}

/* seqan::store -- src/IBF/IBFBuild.cpp:307,505 */
int orc_ibf_store(const orc_ibf *f, const char *path)
{
    if (!f) return ORC_ERR_NULL_FILTER;
    FILE *fp = fopen(path, "wb");
    if (!fp) return ORC_ERR_STORE;
    uint64_t bit_len = f->n_bits + 256;
    int ok = fwrite(&bit_len, 8, 1, fp) == 1 && fwrite(f->words, 8, f->n_words, fp) == f->n_words;
    ok = (fclose(fp) == 0) && ok;
    return ok ? ORC_OK : ORC_ERR_STORE;
}

/* filter.resizeBins(n) -- src/IBF/IBFBuild.cpp:274.  SeqAn keeps noOfBlocks and widens every row when
 * ceil(n/64) grows; bit (row, bin) keeps its coordinates.  UNPINNED: no reference fixture exercises it. */
int orc_ibf_resize_bins(orc_ibf *f, uint64_t new_n_bins)
{
    if (!f) return ORC_ERR_NULL_FILTER;
    if (new_n_bins < f->n_bins) return ORC_ERR_CONFIG;
    uint64_t new_width = (new_n_bins + 63) / 64;
    if (new_width != f->bin_width) {
        uint64_t n_bits = f->n_blocks * new_width * 64;
        uint64_t n_words = (n_bits + 256 + 63) / 64;
        uint64_t *w = (uint64_t *)calloc(n_words, sizeof(uint64_t));
        if (!w) return ORC_ERR_ALLOC;
        for (uint64_t r = 0; r < f->n_blocks; ++r)
            memcpy(w + r * new_width, f->words + r * f->bin_width, f->bin_width * 8);
        free(f->words);
        f->words = w;
        f->n_bits = n_bits;
    }
    f->n_bins = new_n_bins;
    uint64_t rows = f->n_blocks;
    derive(f);
    f->n_blocks = rows;
    write_meta(f);
    return ORC_OK;
}

void orc_ibf_free(orc_ibf *f)
{
    if (f) { free(f->words); free(f); }
}

uint64_t *orc_ibf_words(orc_ibf *f) { return f ? f->words : NULL; }

/* seqan::getNumberOfBins / getKmerSize -- src/IBF/IBFBuild.cpp:380-381 */
void orc_ibf_info(const orc_ibf *f, uint64_t *n_bins, uint64_t *n_hash, uint64_t *k,
                  uint64_t *n_bits, uint64_t *n_words)
{
    if (n_bins) *n_bins = f->n_bins;
    if (n_hash) *n_hash = f->n_hash;
    if (k) *k = f->k;
    if (n_bits) *n_bits = f->n_bits;
    if (n_words) *n_words = f->n_words;
}

/* k-mer value: base-5 polynomial over Dna5 ranks, mod 2^64 (Appendix A.3) */
uint64_t orc_kmer_hash(const char *text, uint64_t k)
{
    uint64_t h = 0;
    for (uint64_t j = 0; j < k; ++j) h = h * 5 + orc_dna5(text[j]);
    return h;
}

/* row selected by hash function i (Appendix A.3):
 * v = (i ^ (k*seed)) * H;  v ^= v >> 27;  row = v % noOfBlocks */
uint64_t orc_hash_row(const orc_ibf *f, uint64_t kmer_value, unsigned i)
{
    uint64_t pre = (uint64_t)i ^ (f->k * ORC_SEED);
    uint64_t v = pre * kmer_value;
    v ^= v >> ORC_SHIFT;
    return v % f->n_blocks;
}

/* Rolling iteration over the k-mers of one strand.  revcomp != 0 walks the
 * reverse-complement string (TSeqRevComp, src/IBF/IBF.hpp:96-97) left to right. */
typedef struct {
    const char *text; uint64_t len, k; int rc;
    uint64_t pos, h, top; /* top = 5^(k-1) */
} kmer_iter;

static inline uint8_t strand_digit(const kmer_iter *it, uint64_t p)
{
    if (!it->rc) return orc_dna5(it->text[p]);
    return comp5(orc_dna5(it->text[it->len - 1 - p]));
}

static int iter_init(kmer_iter *it, const char *text, uint64_t len, uint64_t k, int rc)
{
    it->text = text; it->len = len; it->k = k; it->rc = rc; it->pos = 0; it->h = 0; it->top = 1;
    if (len < k) return 0;
    for (uint64_t j = 1; j < k; ++j) it->top *= 5;
    for (uint64_t j = 0; j < k; ++j) it->h = it->h * 5 + strand_digit(it, j);
    return 1;
}

static inline int iter_next(kmer_iter *it)
{
    if (it->pos + it->k >= it->len) return 0;
    it->h = (it->h - strand_digit(it, it->pos) * it->top) * 5 + strand_digit(it, it->pos + it->k);
    it->pos++;
    return 1;
}

/* seqan::insertKmer(filter, fragment, bin) -- src/IBF/IBFBuild.cpp:189-190 (Appendix A.5).
 * Text shorter than k inserts nothing (unpinned; reachable through quirk Q3). */
static void insert_impl(orc_ibf *f, const char *text, uint64_t len, uint64_t bin, int atomic)
{
    kmer_iter it;
    if (bin >= f->n_bins) return; /* out-of-range bin after the Q3 shift: fenced, unpinned */
    if (!iter_init(&it, text, len, f->k, 0)) return;
    uint64_t pre[16];
    for (unsigned i = 0; i < f->n_hash; ++i) pre[i] = (uint64_t)i ^ (f->k * ORC_SEED);
    const uint64_t wsel = bin >> 6, bit = 1ULL << (bin & 63);
    do {
        for (unsigned i = 0; i < f->n_hash; ++i) {
            uint64_t v = pre[i] * it.h;
            v ^= v >> ORC_SHIFT;
            uint64_t row = v % f->n_blocks;
            uint64_t *w = &f->words[row * f->bin_width + wsel];
            if (atomic) __atomic_fetch_or(w, bit, __ATOMIC_RELAXED);
            else *w |= bit;
        }
    } while (iter_next(&it));
}

void orc_insert(orc_ibf *f, const char *text, uint64_t len, uint64_t bin)
{
    insert_impl(f, text, len, bin, 0);
}

/* seqan::count(filter, text) -- call sites src/IBF/IBFClassify.cpp:97-98,149-150
 * (Appendix A.6): per k-mer AND the h selected rows, +1 on every surviving bin. */
void orc_count(const orc_ibf *f, const char *text, uint64_t len, int revcomp, uint16_t *counts)
{
    kmer_iter it;
    memset(counts, 0, f->n_bins * sizeof(uint16_t));
    if (!iter_init(&it, text, len, f->k, revcomp)) return;
    uint64_t pre[16];
    const uint64_t *rows[16];
    for (unsigned i = 0; i < f->n_hash; ++i) pre[i] = (uint64_t)i ^ (f->k * ORC_SEED);
    do {
        for (unsigned i = 0; i < f->n_hash; ++i) {
            uint64_t v = pre[i] * it.h;
            v ^= v >> ORC_SHIFT;
            rows[i] = &f->words[(v % f->n_blocks) * f->bin_width];
        }
        for (uint64_t w = 0; w < f->bin_width; ++w) {
            uint64_t t = rows[0][w];
            for (unsigned i = 1; i < f->n_hash; ++i) t &= rows[i][w];
            while (t) {
                uint64_t b = 64 * w + (uint64_t)__builtin_ctzll(t);
                if (b < f->n_bins) counts[b]++;
                t &= t - 1;
            }
        }
    } while (iter_next(&it));
}

/* ------------------------------------------------------------------------ */
/* IBF::calculate_filter_size_bits -- src/IBF/IBFBuild.cpp:404-413 */
uint64_t orc_filter_size_bits(uint64_t fragment_length, uint64_t k, uint64_t n_hash,
                              double max_fp, uint64_t n_bins)
{
    uint64_t max_kmer_count = fragment_length - k + 1;
    uint64_t optimal_bins = (uint64_t)(floor(((double)n_bins / 64.0) + 1) * 64);
    uint64_t bin_size_bits = (uint64_t)ceil(
        -1 / (pow(1 - pow(max_fp, 1.0 / (double)n_hash),
                  1.0 / ((double)(n_hash * max_kmer_count))) - 1));
    return bin_size_bits * optimal_bins;
}

/* RationalApproximation -- src/IBF/IBF.hpp:268-277 (Abramowitz-Stegun 26.2.23) */
static double rational_approximation(double t)
{
    const double c[] = {2.515517, 0.802853, 0.010328};
    const double d[] = {1.432788, 0.189269, 0.001308};
    return t - ((c[2] * t + c[1]) * t + c[0]) / (((d[2] * t + d[1]) * t + d[0]) * t + 1.0);
}

/* NormalCDFInverse -- src/IBF/IBF.hpp:284-308 */
static double normal_cdf_inverse(double p)
{
    if (p < 0.5) return -rational_approximation(sqrt(-2.0 * log(p)));
    return rational_approximation(sqrt(-2.0 * log(1.0 - p)));
}

static inline uint16_t to_u16(double x)
{
    /* the reference's (uint16_t) cast of a double; made well-defined via int64 */
    return (uint16_t)(int64_t)x;
}

/* calculateCI -- src/IBF/IBF.hpp:320-338 */
void orc_calculate_ci(double r, uint8_t kmer_size, uint32_t readlen, double confidence,
                      uint16_t *low, uint16_t *high)
{
    double k = (double)kmer_size;
    double q = 1.0 - pow(1.0 - r, k);
    double L = ((double)readlen - k + 1.0);
    double varN = L * (1.0 - q) * (q * (2.0 * k + (2.0 / r) - 1.0) - 2.0 * k)
                  + k * (k - 1.0) * pow((1.0 - q), 2.0)
                  + (2.0 * (1.0 - q) / (pow(r, 2.0))) * ((1.0 + (k - 1.0) * (1.0 - q)) * r - q);
    double alpha = 1 - confidence;
    double z = normal_cdf_inverse(1.0 - alpha / 2.0);
    if (low) *low = to_u16(floor(L * q - z * sqrt(varN)));
    if (high) *high = to_u16(ceil(L * q + z * sqrt(varN)));
}

/* threshold -- src/IBF/IBFClassify.cpp:105-109,156-159: uint16 readlen, int16
 * threshold, implicitly converted to uint16_t at the select/max_matches call
 * (src/IBF/IBF.hpp:178-183), so a negative threshold wraps to >= 32768. */
uint16_t orc_threshold(double r, uint64_t k, uint64_t readlen, double confidence)
{
    uint16_t hi;
    orc_calculate_ci(r, (uint8_t)k, (uint32_t)readlen, confidence, NULL, &hi);
    uint16_t readlen16 = (uint16_t)readlen;
    int16_t thr = (int16_t)((int)readlen16 - (int)k + 1 - (int)hi);
    return (uint16_t)thr;
}

void orc_threshold_lut(double r, uint64_t k, double confidence, uint16_t *lut65536)
{
    for (uint64_t len = 0; len < 65536; ++len) lut65536[len] = orc_threshold(r, k, len, confidence);
}

/* IBF::cutOutNNNs + the concatenation at src/IBF/IBFBuild.cpp:81-88,112-132.
 * Removes every run of 'N'; when the last piece is not followed by an 'N' its
 * final base is dropped (substr(start, seqlen - start - 1), quirk Q1).
 * `out` needs room for len bytes; returns the new length. */
uint64_t orc_cut_out_nnns(const char *seq, uint64_t len, char *out)
{
    uint64_t n = 0, end = 0;
    for (;;) {
        uint64_t start = end;
        while (start < len && seq[start] == 'N') start++;   /* find_first_not_of("N", end) */
        if (start >= len) break;
        end = start;
        while (end < len && seq[end] != 'N') end++;          /* find("N", start) */
        if (end >= len) {                                    /* npos: end > seqlen */
            uint64_t cnt = len - start - 1;
            memcpy(out + n, seq + start, cnt);
            n += cnt;
            break;
        }
        memcpy(out + n, seq + start, end - start);
        n += end - start;
    }
    return n;
}

/* bins reserved per sequence -- src/IBF/IBFBuild.cpp:90 */
uint64_t orc_bins_for_sequence(uint64_t cut_len, uint64_t fragment_length)
{
    return cut_len / fragment_length + 1;
}

/* fragment schedule -- src/IBF/IBFBuild.cpp:165-202 (overlap_length 1500 only
 * clamps fragment 0 to start 0, quirk Q2).  Returns the number of fragments,
 * i.e. the number of bin ids consumed (quirk Q3: may exceed len/F + 1). */
uint64_t orc_fragment_schedule(uint64_t seqlen_u, uint64_t fragment_length, uint64_t k,
                               uint64_t *begin, uint64_t *end, uint64_t cap)
{
    int64_t seqlen = (int64_t)seqlen_u, frag_idx = 0, fragstart = 0;
    uint64_t n = 0;
    while (fragstart < seqlen - 1) {
        uint64_t fragend = (uint64_t)(frag_idx + 1) * fragment_length;
        if (fragend > seqlen_u) fragend = seqlen_u;
        if (n < cap) {
            if (begin) begin[n] = (uint64_t)fragstart;
            if (end) end[n] = fragend;
        }
        n++;
        frag_idx++;
        fragstart = frag_idx * (int64_t)fragment_length - (int64_t)k + 1;
    }
    return n;
}

/* Read::select_matches -- src/IBF/IBFClassify.cpp:16-38 */
int orc_select_matches(const uint16_t *fwd, const uint16_t *rev, uint64_t n_bins, uint16_t thr)
{
    for (uint64_t b = 0; b < n_bins; ++b)
        if (fwd[b] >= thr || rev[b] >= thr) return 1;
    return 0;
}

/* Read::max_matches -- src/IBF/IBFClassify.cpp:48-71 */
uint64_t orc_max_matches(const uint16_t *fwd, const uint16_t *rev, uint64_t n_bins, uint16_t thr)
{
    uint64_t m = 0;
    for (uint64_t b = 0; b < n_bins; ++b)
        if (fwd[b] >= thr || rev[b] >= thr) {
            if (fwd[b] > m) m = fwd[b];
            if (rev[b] > m) m = rev[b];
        }
    return m;
}

/* Read::count_matches -- src/IBF/IBFClassify.cpp:138-171 */
uint64_t orc_count_matches(const orc_ibf *f, const char *read, uint64_t len, double error_rate,
                           double significance)
{
    uint16_t *fwd = (uint16_t *)malloc(2 * f->n_bins * sizeof(uint16_t));
    uint16_t *rev = fwd + f->n_bins;
    orc_count(f, read, len, 0, fwd);
    orc_count(f, read, len, 1, rev);
    uint16_t thr = orc_threshold(error_rate, f->k, len, significance);
    uint64_t m = orc_max_matches(fwd, rev, f->n_bins, thr);
    free(fwd);
    return m;
}

/* Read::classify(std::vector<TIbf>&) + find_matches -- src/IBF/IBFClassify.cpp:81-128,181-226 */
int orc_classify_any(const orc_ibf *const *filters, uint64_t n_filters, const char *read,
                     uint64_t len, double error_rate, double significance, int *status)
{
    if (status) *status = ORC_OK;
    if (n_filters == 0) { if (status) *status = ORC_ERR_NULL_FILTER; return 0; }
    if (len < filters[0]->k) { if (status) *status = ORC_ERR_SHORT_READ; return 0; }
    int found = 0;
    for (uint64_t i = 0; i < n_filters && !found; ++i) {
        const orc_ibf *f = filters[i];
        uint16_t *fwd = (uint16_t *)malloc(2 * f->n_bins * sizeof(uint16_t));
        uint16_t *rev = fwd + f->n_bins;
        orc_count(f, read, len, 0, fwd);
        orc_count(f, read, len, 1, rev);
        uint16_t thr = orc_threshold(error_rate, f->k, len, significance);
        found = orc_select_matches(fwd, rev, f->n_bins, thr);
        free(fwd);
    }
    return found;
}

/* Read::classify(std::vector<IBFMeta>&) -- src/IBF/IBFClassify.cpp:239-297:
 * index of the filter with the strictly greatest count_matches, -1 if all are 0 */
int orc_classify_best(const orc_ibf *const *filters, uint64_t n_filters, const char *read,
                      uint64_t len, double error_rate, double significance, int *status)
{
    if (status) *status = ORC_OK;
    if (n_filters == 0) { if (status) *status = ORC_ERR_NULL_FILTER; return -1; }
    if (len < filters[0]->k) { if (status) *status = ORC_ERR_SHORT_READ; return -1; }
    uint64_t best = 0;
    int best_index = -1;
    for (uint64_t i = 0; i < n_filters; ++i) {
        uint64_t c = orc_count_matches(filters[i], read, len, error_rate, significance);
        if (c > best) { best = c; best_index = (int)i; }
    }
    return best_index;
}

/* Read::classify(filt1, filt2) -- src/IBF/IBFClassify.cpp:299-365:
 * filters whose k exceeds the read length are skipped silently (:318,:340) */
int orc_classify_pair(const orc_ibf *const *filt1, uint64_t n1, const orc_ibf *const *filt2,
                      uint64_t n2, const char *read, uint64_t len, double error_rate,
                      double significance, uint64_t *first, uint64_t *second)
{
    *first = 0; *second = 0;
    if (n1 == 0 || n2 == 0) return ORC_ERR_NULL_FILTER;
    for (uint64_t i = 0; i < n1; ++i)
        if (len >= filt1[i]->k) {
            uint64_t c = orc_count_matches(filt1[i], read, len, error_rate, significance);
            if (c > *first) *first = c;
        }
    for (uint64_t i = 0; i < n2; ++i)
        if (len >= filt2[i]->k) {
            uint64_t c = orc_count_matches(filt2[i], read, len, error_rate, significance);
            if (c > *second) *second = c;
        }
    return ORC_OK;
}

/* check_unblock -- src/main/adaptive_sampling.hpp:35-113
 * returns 0 keep sequencing / 1 unblock / 2 stop_further_data */
int orc_check_unblock(const orc_ibf *const *deplete, uint64_t n_dep, const orc_ibf *const *target,
                      uint64_t n_tgt, const char *read, uint64_t len, double error_rate,
                      double significance, int *status)
{
    int st = ORC_OK, decision = 0;
    if (n_dep > 0 && n_tgt > 0) {
        uint64_t a, b;
        st = orc_classify_pair(deplete, n_dep, target, n_tgt, read, len, error_rate, significance, &a, &b);
        if (a > 0) {
            if (b > 0) {
                /* conf.error_rate -= 0.02; ... += 0.02  (adaptive_sampling.hpp:55-59) */
                double e = error_rate;
                e -= 0.02;
                st = orc_classify_pair(deplete, n_dep, target, n_tgt, read, len, e, significance, &a, &b);
                decision = (a > 0 && b == 0) ? 1 : 0;
            } else decision = 1;
        } else decision = (b > 0) ? 2 : 0;
    } else if (n_dep > 0) {
        decision = orc_classify_best(deplete, n_dep, read, len, error_rate, significance, &st) > -1 ? 1 : 0;
    } else {
        int best = orc_classify_best(target, n_tgt, read, len, error_rate, significance, &st);
        decision = best < 0 ? 1 : 2;
    }
    if (status) *status = st;
    return st == ORC_OK ? decision : 0;
}

/* ------------------------------------------------------------------------ */
/* Batch drivers: the CPU baseline.  One worker per thread pulls blocks of
 * reads (or fragments) from a shared counter; per read it does exactly what
 * count_matches does (two orc_count + threshold + scan over bins). */
typedef struct {
    const orc_ibf *f; const char *bases; const uint64_t *off; uint64_t n;
    const uint16_t *lut; uint16_t *cf, *cr, *mx; uint8_t *hit; uint32_t *amax; uint8_t *sr;
    uint64_t next; pthread_mutex_t mu;
} count_job;

static void count_one(const count_job *j, uint64_t r, uint16_t *tmp)
{
    const orc_ibf *f = j->f;
    const char *read = j->bases + j->off[r];
    uint64_t len = j->off[r + 1] - j->off[r];
    uint16_t *fwd = j->cf ? j->cf + r * f->n_bins : tmp;
    uint16_t *rev = j->cr ? j->cr + r * f->n_bins : tmp + f->n_bins;
    /* 1 = shorter than k (ShortReadException); 2 = longer than the uint16 read
     * length the reference can represent (quirk Q10) -- fenced, not classified */
    int is_short = len < f->k ? 1 : (len > 65535 ? 2 : 0);
    if (j->sr) j->sr[r] = (uint8_t)is_short;
    if (is_short) {
        memset(fwd, 0, f->n_bins * 2); memset(rev, 0, f->n_bins * 2);
        if (j->mx) j->mx[r] = 0;
        if (j->hit) j->hit[r] = 0;
        if (j->amax) j->amax[r] = 0xFFFFFFFFu;
        return;
    }
    orc_count(f, read, len, 0, fwd);
    orc_count(f, read, len, 1, rev);
    uint16_t thr = j->lut[len];
    uint16_t m = 0; uint32_t am = 0xFFFFFFFFu; int hit = 0;
    for (uint64_t b = 0; b < f->n_bins; ++b)
        if (fwd[b] >= thr || rev[b] >= thr) {
            uint16_t c = fwd[b] > rev[b] ? fwd[b] : rev[b];
            if (!hit || c > m) { m = c; am = (uint32_t)b; }
            hit = 1;
        }
    if (j->mx) j->mx[r] = m;
    if (j->hit) j->hit[r] = (uint8_t)hit;
    if (j->amax) j->amax[r] = am;
}

static void *count_worker(void *arg)
{
    count_job *j = (count_job *)arg;
    uint16_t *tmp = (uint16_t *)malloc(2 * j->f->n_bins * sizeof(uint16_t));
    for (;;) {
        pthread_mutex_lock(&j->mu);
        uint64_t b = j->next; j->next += 64;
        pthread_mutex_unlock(&j->mu);
        if (b >= j->n) break;
        uint64_t e = b + 64 < j->n ? b + 64 : j->n;
        for (uint64_t r = b; r < e; ++r) count_one(j, r, tmp);
    }
    free(tmp);
    return NULL;
}

int orc_count_batch(const orc_ibf *f, const char *bases, const uint64_t *read_off,
                    uint64_t n_reads, const uint16_t *thr_lut, uint16_t *counts_fwd,
                    uint16_t *counts_rev, uint16_t *max_count, uint8_t *hit,
                    uint32_t *argmax_bin, uint8_t *short_read, int n_threads)
{
    if (!f) return ORC_ERR_NULL_FILTER;
    count_job j = {f, bases, read_off, n_reads, thr_lut, counts_fwd, counts_rev, max_count, hit,
                   argmax_bin, short_read, 0, PTHREAD_MUTEX_INITIALIZER};
    if (n_threads <= 1) { count_worker(&j); return ORC_OK; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], NULL, count_worker, &j);
    for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    free(th);
    return ORC_OK;
}

typedef struct {
    orc_ibf *f; const char *bases; const uint64_t *fb, *fe, *bin; uint64_t n;
    uint64_t next; pthread_mutex_t mu; int atomic;
} insert_job;

static void *insert_worker(void *arg)
{
    insert_job *j = (insert_job *)arg;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        uint64_t i = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (i >= j->n) break;
        insert_impl(j->f, j->bases + j->fb[i], j->fe[i] - j->fb[i], j->bin[i], j->atomic);
    }
    return NULL;
}

/* insertKmer over a list of fragments (add_sequences_to_filter, src/IBF/IBFBuild.cpp:143-215).
 * With n_threads > 1 words are updated with atomic OR, so the result does not
 * depend on thread interleaving (the reference's own multi-thread build races
 * on binid, quirk Q4; bins are explicit here). */
int orc_insert_batch(orc_ibf *f, const char *bases, const uint64_t *frag_begin,
                     const uint64_t *frag_end, const uint64_t *frag_bin, uint64_t n_frags,
                     int n_threads)
{
    if (!f) return ORC_ERR_NULL_FILTER;
    insert_job j = {f, bases, frag_begin, frag_end, frag_bin, n_frags, 0,
                    PTHREAD_MUTEX_INITIALIZER, n_threads > 1};
    if (n_threads <= 1) { insert_worker(&j); return ORC_OK; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], NULL, insert_worker, &j);
    for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    free(th);
    return ORC_OK;
}
