/* A plain C99 client of include/rb_ibf.h: what a cgo / JNI / N-API binding sees.  Calls only the host-side helpers, so it
 * runs without a GPU; the known answers are the reference's own (createfilter.hpp:148, read.hpp:156-164,
 * IBFBuild.cpp:112-132 as exercised by the libIBFTests fixtures). */
#include "rb_ibf.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CHECK(cond)                                                              \
    do {                                                                         \
        if (!(cond)) { fprintf(stderr, "FAILED line %d: %s\n", __LINE__, #cond); return 1; } \
    } while (0)

int main(void)
{
    uint16_t lo = 0, hi = 0;
    uint16_t *lut = (uint16_t *)malloc(65536 * sizeof(uint16_t));
    uint64_t begin[8], end[8];
    char out[32];
    rb_ibf_info_t info;
    int status = 0;

    CHECK(lut != NULL);
    /* IBF::calculate_filter_size_bits: 100 000-base fragments, k = 13, 3 hash functions, 1 % -> 79 121 216 bits for 2 bins */
    CHECK(rb_ibf_size_bits(100000, 13, 3, 0.01, 2) == 79121216ull);
    /* calculateCI(0.1, 13, 35, 0.95) == (5, 30); threshold -7 wraps to 65 529; a 354-base read gets 36 */
    CHECK(rb_calculate_ci(0.1, 13, 35, 0.95, &lo, &hi) == RB_OK && lo == 5 && hi == 30);
    CHECK(rb_threshold_lut(0.1, 0.95, 13, lut) == RB_OK && lut[35] == 65529 && lut[354] == 36 && lut[250] == 18);
    CHECK(rb_threshold_lut(0.0, 0.95, 13, lut) == RB_ERR_INVALID_CONFIG);          /* checked variant rejects, raw does not */
    CHECK(rb_threshold_lut_raw(0.0, 0.95, 13, lut) == RB_OK);
    /* cutOutNNNs: runs of N removed, and the last base dropped when the sequence does not end in N (quirk Q1) */
    CHECK(rb_cut_out_nnns("ACGTNNNACGT", 11, out) == 7 && memcmp(out, "ACGTACG", 7) == 0);
    CHECK(rb_cut_out_nnns("ACGTNN", 6, out) == 4 && memcmp(out, "ACGT", 4) == 0);
    /* fragment schedule: fragment 0 = [0, F), fragment j = [j F - k + 1, (j + 1) F) */
    CHECK(rb_fragment_schedule(250000, 100000, 13, begin, end, 8) == 3);
    CHECK(begin[0] == 0 && end[0] == 100000 && begin[1] == 99988 && end[1] == 200000 && begin[2] == 199988 && end[2] == 250000);
    /* status codes are the reference's exception classes; nothing falls back to the CPU */
    CHECK(strlen(rb_status_string(RB_ERR_SHORT_READ)) > 0 && strlen(rb_status_string(RB_ERR_NO_DEVICE)) > 0);
    CHECK(rb_ibf_info(NULL, &info) == RB_ERR_NULL_FILTER);
    CHECK(RB_KEY_HIT(0) == 0 && RB_KEY_ARGMAX_BIN(0) == 0xFFFFFFFFu);
    CHECK(RB_KEY_MAX_COUNT(((uint64_t)1 << 48) | ((uint64_t)282 << 32) | 0xFFFFFFFEu) == 282);
    CHECK(RB_KEY_ARGMAX_BIN(((uint64_t)1 << 48) | ((uint64_t)282 << 32) | 0xFFFFFFFEu) == 1);
    if (rb_device_count() <= 0) {                                                   /* no GPU: loud failure, no fallback */
        rb_ibf *f = rb_ibf_create(2, 3, 13, 79121216ull, 0, &status);
        CHECK(f == NULL && status == RB_ERR_NO_DEVICE && strlen(rb_last_error()) > 0);
    }
    free(lut);
    puts("c client ok");
    return 0;
}
