// rb_ibf_api.cu -- the C ABI declared in include/rb_ibf.h.
//
// Host side of the drop-in boundary: filter life cycle (sdsl/SeqAn file format, HBM
// residency, optional bin sharding and L2 persistence), the FP64 scalar helpers that must
// agree bit for bit with the reference (thresholds, sizing), and the batch entry points
// that enqueue the kernels of ibf_count.cu / ibf_insert.cu.  There is no CPU fallback:
// every compute entry point fails with RB_ERR_NO_DEVICE when no CUDA device is usable.
#include "../../include/rb_ibf.h"
#include "ibf_kernels.cuh"
#include "host_pack.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <dlfcn.h>
#include <sys/stat.h>
#include <vector>

namespace {

thread_local std::string g_last_error;
std::atomic<uint64_t> g_launches{0};
std::atomic<uint64_t> g_h2d_bytes{0}, g_d2h_bytes{0};     // moved by rb_ibf_count_batch (copies and mapped result stores)
std::atomic<int> g_count_kernel{0};

int fail(int status, const std::string &msg)
{
    g_last_error = msg;
    return status;
}

#define RB_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(e__ == cudaErrorMemoryAllocation ? RB_ERR_ALLOC : RB_ERR_CUDA,             \
                        std::string(#expr) + ": " + cudaGetErrorString(e__));                      \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; return; }
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

}  // namespace

struct rb_ibf {
    int device = 0, shard = 0, n_shards = 1, sm_count = 148;
    uint64_t n_bins = 0, n_hash = 0, k = 0, n_bits = 0, bin_width = 0, n_blocks = 0;
    uint64_t col_begin = 0, col_words = 0, n_bins_local = 0;
    uint64_t n_local_words = 0;   // words allocated in HBM
    uint64_t *d_words = nullptr;
    unsigned int *d_err = nullptr;
    bool l2_persist = false;
    rb::HashParams hp{};
    // direct k-mer table (ibf_table.cu): built lazily by the first count call, dropped by inserts
    mutable std::mutex table_mu;
    mutable uint64_t *d_table = nullptr;
    mutable uint64_t table_entries = 0;
    mutable int table_span = 0;             // k-mers per entry: 1 = one k-mer per lane (ibf_table.cu), 2..4 = window
                                            // entries loaded by adjacent lanes (ibf_wtable.cu)
    mutable uint64_t table_bytes = 0;
    // wide filters: k-mer postings table (ibf_postings.cu) instead of the dense window table
    mutable double post_mean_units = 0;     // postings lists: mean list length in 16-byte units (picks the lookup kernel)
    mutable int table_kind = 0;             // 0 none, 1 dense k-mer / window table, 2 postings, 4 k-mer table loaded by lane groups
    mutable uint32_t *d_post_ptr = nullptr;
    mutable uint16_t *d_post_ids = nullptr;
    // ... or the same lists as one fixed slot per k-mer + an overflow area (table_kind 3: automatic for mid-length lists, else RB_POSTINGS_LAYOUT=slots)
    mutable uint8_t *d_slots = nullptr;
    mutable uint16_t *d_slot_ovf = nullptr;
    mutable uint32_t slot_bytes = 0;
    mutable bool table_tried = false;
    // working memory of the column build (ibf_insert.cu), kept between insert calls; released when the k-mer table is built
    mutable rb::ScratchBuf build_scratch;
    mutable uint64_t table_budget = 0;      // 0 = automatic
    // streams, events and staging buffers of rb_ibf_count_batch, kept between calls (one set per concurrent caller)
    mutable std::mutex ctx_mu;
    mutable std::vector<struct CallCtx *> ctx_free;
    // rb_ibf_count_batch, large batches: bit planes packed by the host threads or the ASCII bases as they are -- whichever
    // this host moves faster (it depends on the cores and the memory bandwidth each GPU's process is left with), decided
    // by timing the 2nd call of each kind.  RB_HOST_PACK=1 / 0 pins the choice.
    mutable std::mutex pol_mu;
    mutable int pol_calls = 0;                       // large packable calls seen while undecided
    mutable double pol_ns_per_base[2] = {0.0, 0.0};  // [0] packed, [1] ASCII
    mutable int pol_choice = -1;                     // -1 undecided, 0 packed, 1 ASCII
};

namespace {

// ---- host FP64 helpers ---------------------------------------------------------------------
// RationalApproximation / NormalCDFInverse, src/IBF/IBF.hpp:268-308
double rational_approx(double t)
{
    const double c0 = 2.515517, c1 = 0.802853, c2 = 0.010328;
    const double d0 = 1.432788, d1 = 0.189269, d2 = 0.001308;
    return t - ((c2 * t + c1) * t + c0) / (((d2 * t + d1) * t + d0) * t + 1.0);
}

double normal_cdf_inverse(double p)
{
    return p < 0.5 ? -rational_approx(std::sqrt(-2.0 * std::log(p)))
                   : rational_approx(std::sqrt(-2.0 * std::log(1.0 - p)));
}

// double -> uint16 as x86-64 does it for the reference (cvttsd2si, low 16 bits); NaN / out of range give 0 there
uint16_t cast_u16(double x) { return (x >= -9.2e18 && x <= 9.2e18) ? (uint16_t)(int64_t)x : (uint16_t)0; }

// calculateCI, src/IBF/IBF.hpp:320-338 (kmer_size arrives as uint8_t there)
void calculate_ci(double r, uint32_t kmer_size, uint32_t readlen, double confidence, uint16_t *lo, uint16_t *hi)
{
    const double k = (double)(uint8_t)kmer_size;
    const double q = 1.0 - std::pow(1.0 - r, k);
    const double L = ((double)readlen - k + 1.0);
    const double varN = L * (1.0 - q) * (q * (2.0 * k + (2.0 / r) - 1.0) - 2.0 * k)
                        + k * (k - 1.0) * std::pow((1.0 - q), 2.0)
                        + (2.0 * (1.0 - q) / (std::pow(r, 2.0))) * ((1.0 + (k - 1.0) * (1.0 - q)) * r - q);
    const double alpha = 1 - confidence;
    const double z = normal_cdf_inverse(1.0 - alpha / 2.0);
    if (lo) *lo = cast_u16(std::floor(L * q - z * std::sqrt(varN)));
    if (hi) *hi = cast_u16(std::ceil(L * q + z * std::sqrt(varN)));
}

int check_device(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        return fail(RB_ERR_NO_DEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(RB_ERR_INVALID_ARG, "device index out of range");
    return RB_OK;
}

int derive_geometry(rb_ibf *f, int shard, int n_shards)
{
    if (f->n_bins == 0 || f->n_hash == 0 || f->n_hash > (uint64_t)rb::kMaxHash || f->k == 0 || f->k > 32)
        return fail(RB_ERR_INVALID_CONFIG, "bins/hash functions/k-mer size out of range");
    if (f->n_bits == 0 || (f->n_bits % 64) != 0) return fail(RB_ERR_INVALID_CONFIG, "filter size must be a positive multiple of 64 bits");
    f->bin_width = (f->n_bins + 63) / 64;
    f->n_blocks = f->n_bits / (64 * f->bin_width);
    if (f->n_blocks == 0) return fail(RB_ERR_INVALID_CONFIG, "filter smaller than one row");
    if (n_shards < 1 || shard < 0 || shard >= n_shards) return fail(RB_ERR_INVALID_ARG, "bad shard index");
    f->shard = shard;
    f->n_shards = n_shards;
    f->col_begin = f->bin_width * (uint64_t)shard / (uint64_t)n_shards;
    const uint64_t col_end = f->bin_width * (uint64_t)(shard + 1) / (uint64_t)n_shards;
    f->col_words = col_end - f->col_begin;
    if (f->col_words == 0) return fail(RB_ERR_INVALID_ARG, "more shards than row words");
    const uint64_t bin_end = std::min<uint64_t>(f->n_bins, 64 * col_end);
    f->n_bins_local = bin_end - 64 * f->col_begin;
    f->n_local_words = n_shards == 1 ? f->n_bits / 64 : f->n_blocks * f->col_words;
    f->hp = rb::make_hash_params(f->n_blocks, (uint32_t)f->k, (uint32_t)f->n_hash);
    return RB_OK;
}

int alloc_device(rb_ibf *f, bool zero)
{
    RB_CUDA(cudaSetDevice(f->device));
    cudaDeviceProp prop{};
    RB_CUDA(cudaGetDeviceProperties(&prop, f->device));
    f->sm_count = prop.multiProcessorCount;
    RB_CUDA(cudaMalloc(&f->d_words, f->n_local_words * 8));
    // [0] insert error (bin >= n_bins), [1] count kernel gave up waiting for a bulk copy, [2..3] counters of the slot-table build
    RB_CUDA(cudaMalloc(&f->d_err, 4 * sizeof(unsigned int)));
    RB_CUDA(cudaMemset(f->d_err, 0, 4 * sizeof(unsigned int)));
    if (zero) RB_CUDA(cudaMemset(f->d_words, 0, f->n_local_words * 8));
    // host-buffer calls take their staging buffers from the stream-ordered pool; keep freed blocks
    // cached instead of returning them to the driver at every synchronisation
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, f->device) == cudaSuccess) {
        uint64_t keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
    // a filter that fits the persisting-L2 carve-out is pinned there for count launches
    const size_t bytes = f->n_local_words * 8;
    if (prop.persistingL2CacheMaxSize > 0 && bytes <= (size_t)prop.persistingL2CacheMaxSize &&
        bytes <= (size_t)prop.accessPolicyMaxWindowSize) {
        // device-wide limit shared by all filters of the device: only ever raised (the largest small filter decides)
        size_t cur = 0;
        if (cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize) != cudaSuccess) { cudaGetLastError(); cur = 0; }
        if (cur >= bytes || cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, bytes) == cudaSuccess) f->l2_persist = true;
        else cudaGetLastError();
    }
    return RB_OK;
}

void free_call_contexts(rb_ibf *f);

void destroy(rb_ibf *f)
{
    if (!f) return;
    free_call_contexts(f);
    if (f->d_words || f->d_err || f->d_table || f->d_post_ptr || f->d_post_ids || f->d_slots || f->d_slot_ovf || f->build_scratch.p) {
        DeviceGuard g(f->device);
        if (f->build_scratch.p) cudaFree(f->build_scratch.p);
        if (f->d_words) cudaFree(f->d_words);
        if (f->d_err) cudaFree(f->d_err);
        if (f->d_table) cudaFree(f->d_table);
        if (f->d_post_ptr) cudaFree(f->d_post_ptr);
        if (f->d_post_ids) cudaFree(f->d_post_ids);
        if (f->d_slots) cudaFree(f->d_slots);
        if (f->d_slot_ovf) cudaFree(f->d_slot_ovf);
    }
    delete f;
}

// copy `rows` rows of the host matrix (row pitch = bin_width words) into the local slice
int upload_rows(rb_ibf *f, const uint64_t *host_rows, uint64_t row0, uint64_t rows, cudaStream_t st)
{
    RB_CUDA(cudaMemcpy2DAsync(f->d_words + row0 * f->col_words, f->col_words * 8, host_rows + f->col_begin,
                              f->bin_width * 8, f->col_words * 8, rows, cudaMemcpyHostToDevice, st));
    return RB_OK;
}

struct PinnedPair {
    uint64_t *buf[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaStream_t st = nullptr;
    size_t words = 0;
    int init(size_t w)
    {
        words = w;
        RB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            RB_CUDA(cudaMallocHost(&buf[i], w * 8));
            RB_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
        }
        return RB_OK;
    }
    ~PinnedPair()
    {
        for (int i = 0; i < 2; ++i) {
            if (buf[i]) cudaFreeHost(buf[i]);
            if (ev[i]) cudaEventDestroy(ev[i]);
        }
        if (st) cudaStreamDestroy(st);
    }
};

constexpr size_t kStageBytes = 64u << 20;

int sticky_insert_error(const rb_ibf *f)
{
    unsigned int flag = 0;
    RB_CUDA(cudaMemcpy(&flag, f->d_err, sizeof(flag), cudaMemcpyDeviceToHost));
    if (flag) {
        cudaMemset(f->d_err, 0, sizeof(flag));
        return fail(RB_ERR_INSERT_SEQUENCE, "a fragment named a bin >= number of bins (not inserted)");
    }
    return RB_OK;
}

rb::FilterView view_of(const rb_ibf *f)
{
    rb::FilterView v{};
    v.words = f->d_words;
    v.stride = f->col_words;
    v.bin_begin = 64 * f->col_begin;
    v.n_bins_local = f->n_bins_local;
    v.hp = f->hp;
    return v;
}

// ---- direct k-mer table policy -----------------------------------------------------------------
constexpr uint64_t kTableMinReads = 1024;   // batches smaller than this never trigger the (GB-sized) table build

// span = consecutive k-mers per entry.  span 1: 4^k entries of 2*col_words words, one lane per entry
// (ibf_table.cu, rows <= 4 words).  span 2..4: (k+span-1)-base windows, canonical when that length is odd,
// `lanes` slots of 2*col_words words per entry, loaded by adjacent lanes (ibf_wtable.cu, rows <= 2 words).
// rows of 3..16 words: one k-mer per entry, the row padded to 4, 8 or 16 words, loaded by that many lanes (ibf_ctable.cu).
// RB_CTABLE=0 turns the layout off (measurements: rows of 3-4 words fall back to ibf_table.cu, wider ones to postings).
// Rows of 17..32 words (512-byte entries, four lines per position) beat the short-list postings kernel only when the lists are
// long: measured at 2 020 bins of 100 kb, 3.3 units per list, 32.7 M chunks/s (34 GB table) against 34.6 M (3.8 GB of lists),
// while 8-unit lists run at ~20 M.  So that width takes the table only when a sample of the lists averages more than 4 units
// (RB_CTABLE_WIDE=1 / 0 forces the answer: tests, measurements).
uint64_t ctable_bytes(const rb_ibf *f, cudaStream_t st = nullptr)
{
    const char *e = std::getenv("RB_CTABLE");
    const bool on = !(e && e[0] == '0');
    const int lanes = rb::ctable_lanes(f->col_words);
    if (!on || !lanes || f->k > 16) return 0;
    if (lanes == 32) {
        const char *w = std::getenv("RB_CTABLE_WIDE");
        bool take = w && w[0] == '1';
        if (!w && rb::postings_applicable(view_of(f))) {
            const uint64_t n_kmers = 1ull << (2 * f->k);
            const uint32_t n_sample = (uint32_t)std::min<uint64_t>(65536, n_kmers);
            uint32_t *d_tmp = nullptr;
            double mean_units = 0;
            if (cudaMalloc(&d_tmp, (size_t)n_sample * 4) == cudaSuccess) {
                if (rb::postings_sample_units(view_of(f), d_tmp, n_sample, &mean_units, f->sm_count, st) >= 0) take = mean_units > 4.0;
                cudaFree(d_tmp);
            }
            cudaGetLastError();
        } else if (!w) take = true;                                      // no lists possible (k = 16): the table or nothing
        if (!take) return 0;
    }
    return (1ull << (2 * f->k)) * (uint64_t)lanes * 16;
}

uint64_t table_bytes_needed(const rb_ibf *f, int span)
{
    if (f->col_words == 0 || f->k + span - 1 > 16) return 0;
    if (span == 1) return f->col_words > 4 ? 0 : (1ull << (2 * f->k)) * 2 * f->col_words * 8;
    int lanes = 0;
    uint64_t n_entries = 0;
    if (!rb::wtable_geometry(f->col_words, (uint32_t)f->k, span, &lanes, nullptr, &n_entries)) return 0;
    return n_entries * (uint64_t)lanes * 2 * f->col_words * 8;
}

// Mid-length lists -- 2.6 to 14 units on average, the 2 000..8 000-bin filters -- are faster in slots of one or two lines
// read by groups of 8 / 16 lanes (count_slots_sub_kernel) than as pointer + lists: no dependent pointer fetch.  Measured
// 38.6 / 30.2 / 21.1 M chunks/s against 34.8 / 25.4 / 16.7 M at 2 020 / 4 040 / 8 080 bins; shorter lists run faster through
// the 2-lane list kernel, longer ones need slots of four lines and more, where the lists win
// (profiles/r2_ah_slots_sub_sweep.jsonl).  Returns the slot size (128 or 256) or 0; RB_POSTINGS_LAYOUT overrides.
uint32_t auto_slot_bytes(const std::vector<uint32_t> &lengths, uint32_t k, uint64_t budget, uint64_t *ovf_units, uint64_t *total)
{
    if (lengths.empty()) return 0;
    double units = 0;
    for (uint32_t n : lengths) units += (n + 7) / 8;
    const double mean_units = units / (double)lengths.size();
    if (mean_units < 2.6 || mean_units > 14.0) return 0;
    const double n_kmers = (double)(1ull << (2 * k));
    uint32_t best = 0;
    double best_cost = 0;
    for (uint32_t sb = 128; sb <= 256; sb += 128) {
        const uint32_t cap = (sb - 8) / 2;
        double over_units = 0, over_lines = 0;
        for (uint32_t n : lengths)
            if (n > cap) { over_units += (n + 7) / 8; over_lines += ((n + 7) / 8 * 16 + 127) / 128 + 1; }
        const double cost = sb + 128.0 * over_lines / (double)lengths.size();
        const uint64_t ou = (uint64_t)(2.0 * over_units / (double)lengths.size() * n_kmers) + (1u << 20);
        const uint64_t bytes = (uint64_t)(n_kmers * sb) + ou * 16;
        if (bytes > budget || ou > 0xFFFFFFF0ull) continue;
        if (!best || cost < best_cost) { best = sb; best_cost = cost; *ovf_units = ou; *total = bytes; }
    }
    return best;
}

// Expected size of the postings table of a wide filter (0: not applicable), from a sample of 65 536 k-mers.
uint64_t postings_estimate_bytes(const rb_ibf *f, cudaStream_t st)
{
    const rb::FilterView fv = view_of(f);
    const uint64_t n_kmers = 1ull << (2 * f->k);
    const uint32_t n_sample = (uint32_t)std::min<uint64_t>(65536, n_kmers);
    uint32_t *d_tmp = nullptr;
    const char *lay = std::getenv("RB_POSTINGS_LAYOUT");
    if (lay && lay[0] == 's' && rb::slots_applicable(fv)) {
        if (cudaMalloc(&d_tmp, (size_t)n_sample * 4) != cudaSuccess) { cudaGetLastError(); return 0; }
        std::vector<uint32_t> lengths;
        const int r0 = rb::slots_sample_lengths(fv, d_tmp, n_sample, &lengths, f->sm_count, st);
        cudaFree(d_tmp);
        uint32_t sb = 0;
        uint64_t ou = 0, total = 0;
        if (r0 < 0 || !rb::slots_choose(lengths, (uint32_t)f->k, ~0ull, &sb, &ou, &total)) { cudaGetLastError(); return 0; }
        return total;
    }
    if (!lay && rb::slots_applicable(fv)) {                                // automatic: slots of one or two lines for mid-length lists
        if (cudaMalloc(&d_tmp, (size_t)n_sample * 4) != cudaSuccess) { cudaGetLastError(); return 0; }
        std::vector<uint32_t> lengths;
        const int r0 = rb::slots_sample_lengths(fv, d_tmp, n_sample, &lengths, f->sm_count, st);
        cudaFree(d_tmp);
        uint64_t ou = 0, total = 0;
        if (r0 >= 0 && auto_slot_bytes(lengths, (uint32_t)f->k, ~0ull, &ou, &total)) return total;
        cudaGetLastError();
    }
    if (!rb::postings_applicable(fv)) return 0;
    if (cudaMalloc(&d_tmp, (size_t)n_sample * 4) != cudaSuccess) { cudaGetLastError(); return 0; }
    double mean_units = 0;
    const int r = rb::postings_sample_units(fv, d_tmp, n_sample, &mean_units, f->sm_count, st);
    cudaFree(d_tmp);
    if (r < 0) { cudaGetLastError(); return 0; }
    return (uint64_t)(mean_units * (double)n_kmers * 16.0) + (n_kmers + 1) * 4;
}

// Builds the table on `st` if the policy allows it; returns the device pointer or null.
// force: ignore the batch-size heuristic (tests, explicit rb_ibf_enable_kmer_table).
const uint64_t *ensure_table(const rb_ibf *f, cudaStream_t st, bool force, uint64_t n_reads = ~0ull)
{
    std::lock_guard<std::mutex> lock(f->table_mu);
    if (f->d_table) return f->d_table;
    if (f->table_kind == 2) return reinterpret_cast<const uint64_t *>(f->d_post_ptr);
    if (f->table_kind == 3) return reinterpret_cast<const uint64_t *>(f->d_slots);
    if (f->table_tried && !force) return nullptr;
    if (!force && n_reads < kTableMinReads) return nullptr;     // not "tried": a later large batch builds it
    f->table_tried = true;
    const char *env = std::getenv("RB_KMER_TABLE");
    if (!force && env && env[0] == '0') return nullptr;
    if (f->build_scratch.p) {             // the build is over: its working memory goes to the table
        cudaFree(f->build_scratch.p);
        f->build_scratch = rb::ScratchBuf{};
    }
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    // Sized for 180 GB of HBM: the widest window whose table fits 60 % of the free memory (at most 80 GiB)
    // wins, because the classify kernel is bound by REQUESTS (~43 G/s for anything up to 128 contiguous
    // bytes, profiles/r1_d_gather_sweep3_coop.jsonl) and a span-S entry answers S positions per request.
    uint64_t cap = 80ull << 30;
    if (const char *gb = std::getenv("RB_KMER_TABLE_MAX_GB")) cap = (uint64_t)std::max(0, std::atoi(gb)) << 30;
    const uint64_t budget = f->table_budget ? f->table_budget : std::min<uint64_t>((uint64_t)(free_b * 0.6), cap);
    if (const uint64_t need = ctable_bytes(f, st); need && need <= budget && need <= free_b) {
        uint64_t *t = nullptr;
        if (cudaMalloc(&t, need) == cudaSuccess) {
            const uint64_t n_kmers = 1ull << (2 * f->k);
            const int n = rb::launch_ctable_build(view_of(f), t, n_kmers, f->sm_count, st);
            if (n < 0 || cudaStreamSynchronize(st) != cudaSuccess) { cudaFree(t); cudaGetLastError(); return nullptr; }
            g_launches += (uint64_t)n;
            f->d_table = t;
            f->table_kind = 4;
            f->table_entries = n_kmers;
            f->table_span = 1;
            f->table_bytes = need;
            return t;
        }
        cudaGetLastError();               // no memory after all: the smaller layouts below
    }
    if (f->col_words > 4) {
        const rb::FilterView fv = view_of(f);
        // wide rows: postings.  Default layout: pointer + variable-length lists, loaded straight into registers.
        // RB_POSTINGS_LAYOUT=slots: one fixed slot per k-mer fetched by bulk copies into shared-memory rings -- fewer DRAM bytes
        // and no pointer chase, but the counting is bound by the shared-memory pipe (ATOMS wavefronts), and staging the lists
        // through shared memory adds a third to its load: measured 6.8 ms against 6.3 ms per 65 536 chunks on BASELINE
        // config #3 (profiles/r2_c_slots_v2_cfg3_ncu_full.json), so it stays opt-in.
        // Without RB_POSTINGS_LAYOUT: slots of one or two lines when the sampled lists are of middle length (auto_slot_bytes),
        // else the lists; whatever goes wrong with automatic slots falls back to the lists.
        const char *lay = std::getenv("RB_POSTINGS_LAYOUT");
        const bool want_slots = lay && lay[0] == 's';
        const bool auto_slots = !lay;
        if ((want_slots || auto_slots) && rb::slots_applicable(fv)) do {
            const uint64_t n_kmers = 1ull << (2 * f->k);
            const uint32_t n_sample = (uint32_t)std::min<uint64_t>(65536, n_kmers);
            uint32_t *d_tmp = nullptr;
            if (cudaMalloc(&d_tmp, (size_t)n_sample * 4) != cudaSuccess) { cudaGetLastError(); if (auto_slots) break; return nullptr; }
            std::vector<uint32_t> lengths;
            const int r0 = rb::slots_sample_lengths(fv, d_tmp, n_sample, &lengths, f->sm_count, st);
            cudaFree(d_tmp);
            if (r0 < 0) { cudaGetLastError(); if (auto_slots) break; return nullptr; }
            uint32_t slot_bytes = 0;
            uint64_t ovf_units = 0, total = 0;
            if (auto_slots) {
                slot_bytes = auto_slot_bytes(lengths, (uint32_t)f->k, budget, &ovf_units, &total);
                if (!slot_bytes) break;                                   // short or long lists, or no room: pointer + lists
            } else if (const char *sb = std::getenv("RB_SLOT_BYTES")) {   // tests / tuning: force the slot size
                std::vector<uint32_t> none;
                slot_bytes = (uint32_t)std::max(128, std::min(4096, std::atoi(sb) / 128 * 128));
                double over_units = 0;
                for (uint32_t n : lengths) if (n > (slot_bytes - 8) / 2) over_units += (n + 7) / 8;
                ovf_units = (uint64_t)(2.0 * over_units / (double)lengths.size() * (double)n_kmers) + (1u << 20);
                total = n_kmers * slot_bytes + ovf_units * 16;
                if (total > budget || ovf_units > 0xFFFFFFF0ull) return nullptr;
            } else if (!rb::slots_choose(lengths, (uint32_t)f->k, budget, &slot_bytes, &ovf_units, &total)) return nullptr;
            bool give_up_slots = false;
            for (int attempt = 0; attempt < 2 && !give_up_slots; ++attempt) {
                uint8_t *d_slots = nullptr;
                uint16_t *d_ovf = nullptr;
                if (cudaMalloc(&d_slots, n_kmers * slot_bytes) != cudaSuccess) { cudaGetLastError(); if (auto_slots) { give_up_slots = true; break; } return nullptr; }
                if (cudaMalloc(&d_ovf, ovf_units * 16) != cudaSuccess) { cudaFree(d_slots); cudaGetLastError(); if (auto_slots) { give_up_slots = true; break; } return nullptr; }
                const int r1 = rb::slots_fill(fv, d_slots, slot_bytes, d_ovf, ovf_units, f->d_err + 2, f->sm_count, st);
                if (r1 == 1) {
                    g_launches += 2;
                    f->d_slots = d_slots; f->d_slot_ovf = d_ovf; f->slot_bytes = slot_bytes;
                    f->table_kind = 3; f->table_span = 1; f->table_entries = n_kmers;
                    f->table_bytes = n_kmers * slot_bytes + ovf_units * 16;
                    return reinterpret_cast<const uint64_t *>(d_slots);
                }
                cudaFree(d_slots); cudaFree(d_ovf); cudaGetLastError();
                if (r1 != -3) { if (auto_slots) break; return nullptr; }
                ovf_units *= 4;                                            // the sample underestimated the long lists: once more
                if (n_kmers * slot_bytes + ovf_units * 16 > budget || ovf_units > 0xFFFFFFF0ull) { if (auto_slots) break; return nullptr; }
            }
            if (!auto_slots) return nullptr;
        } while (false);
        // pointer + lists.  Size known only after counting: sample first, then count everything, then fill.
        if (!rb::postings_applicable(fv)) return nullptr;
        const uint64_t n_kmers = 1ull << (2 * f->k);
        const uint64_t ptr_bytes = (n_kmers + 1) * 4;
        if (ptr_bytes > budget) return nullptr;
        uint32_t *d_ptr = nullptr;
        if (cudaMalloc(&d_ptr, ptr_bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        auto give_up = [&]() -> const uint64_t * { cudaFree(d_ptr); cudaGetLastError(); return nullptr; };
        const uint32_t n_sample = (uint32_t)std::min<uint64_t>(65536, n_kmers);
        double mean_units = 0;
        if (rb::postings_sample_units(fv, d_ptr, n_sample, &mean_units, f->sm_count, st) < 0) return give_up();
        const double est_units = mean_units * (double)n_kmers;
        if (est_units * 16.0 * 0.95 + (double)ptr_bytes > (double)budget || est_units > 4.0e9) return give_up();
        uint64_t total_units = 0;
        int n1 = rb::postings_build_ptr(fv, d_ptr, &total_units, f->sm_count, st);
        if (n1 < 0) return give_up();
        const uint64_t ids_bytes = std::max<uint64_t>(16, total_units * 16);
        if (ids_bytes + ptr_bytes > budget) return give_up();
        uint16_t *d_ids = nullptr;
        if (cudaMalloc(&d_ids, ids_bytes) != cudaSuccess) return give_up();
        int n2 = rb::postings_fill(fv, d_ptr, d_ids, f->sm_count, st);
        if (n2 < 0 || cudaStreamSynchronize(st) != cudaSuccess) { cudaFree(d_ids); return give_up(); }
        g_launches += (uint64_t)(1 + n1 + n2);
        f->d_post_ptr = d_ptr;
        f->d_post_ids = d_ids;
        f->post_mean_units = (double)total_units / (double)n_kmers;
        f->table_kind = 2;
        f->table_span = 1;
        f->table_entries = n_kmers;
        f->table_bytes = ids_bytes + ptr_bytes;
        return reinterpret_cast<const uint64_t *>(d_ptr);
    }
    int max_span = f->col_words <= 2 ? 4 : 1;
    if (const char *sp_env = std::getenv("RB_KMER_TABLE_SPAN")) max_span = std::min(4, std::max(1, std::atoi(sp_env)));
    int span = 0;
    uint64_t need = 0;
    for (int sp = max_span; sp >= 1 && !span; --sp) {   // the widest window that fits the budget
        const uint64_t b = table_bytes_needed(f, sp);
        if (b && b <= budget && b <= free_b) { span = sp; need = b; }
    }
    // An explicit request with the automatic budget: when only the one-k-mer table is possible and it misses the 60 % / 80 GiB
    // rule (k = 16: 137 GB for 65-128 bins) it still gets up to 85 % of the free memory -- the alternative is six hashed probes
    // per k-mer and strand, 4-5 times slower.  A lazy build inside a count call never takes that much.
    if (!span && force && f->table_budget == 0) {
        const uint64_t b = table_bytes_needed(f, 1);
        if (b && b <= (uint64_t)(free_b * 0.85)) { span = 1; need = b; }
    }
    if (!span) return nullptr;
    uint64_t *t = nullptr;
    if (cudaMalloc(&t, need) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    const uint64_t entries = 1ull << (2 * f->k);
    int n = span == 1 ? rb::launch_table_build(view_of(f), t, entries, 1, f->sm_count, st)
                      : rb::launch_wtable_build(view_of(f), t, span, f->sm_count, st);
    // other host threads may use the table from their own streams right away: finish the build first
    if (n < 0 || cudaStreamSynchronize(st) != cudaSuccess) { cudaFree(t); cudaGetLastError(); return nullptr; }
    g_launches += (uint64_t)n;
    f->d_table = t;
    f->table_kind = 1;
    f->table_entries = entries;
    f->table_span = span;
    f->table_bytes = need;
    return t;
}

void drop_table(rb_ibf *f)
{
    std::lock_guard<std::mutex> lock(f->table_mu);
    if (f->d_table || f->d_post_ptr || f->d_post_ids || f->d_slots || f->d_slot_ovf) {
        cudaDeviceSynchronize();
        if (f->d_table) cudaFree(f->d_table);
        if (f->d_post_ptr) cudaFree(f->d_post_ptr);
        if (f->d_post_ids) cudaFree(f->d_post_ids);
        if (f->d_slots) cudaFree(f->d_slots);
        if (f->d_slot_ovf) cudaFree(f->d_slot_ovf);
    }
    f->d_table = nullptr;
    f->d_post_ptr = nullptr;
    f->d_post_ids = nullptr;
    f->d_slots = nullptr;
    f->d_slot_ovf = nullptr;
    f->slot_bytes = 0;
    f->table_kind = 0;
    f->table_entries = 0;
    f->table_span = 0;
    f->table_bytes = 0;
    f->table_tried = false;
}

// RAII for the L2 access-policy window around count launches
struct L2Window {
    cudaStream_t st;
    bool active = false;
    L2Window(const rb_ibf *f, cudaStream_t s) : st(s)
    {
        if (!f->l2_persist) return;
        cudaStreamAttrValue attr{};
        attr.accessPolicyWindow.base_ptr = f->d_words;
        attr.accessPolicyWindow.num_bytes = f->n_local_words * 8;
        attr.accessPolicyWindow.hitRatio = 1.0f;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr) == cudaSuccess) active = true;
        else cudaGetLastError();
    }
    ~L2Window()
    {
        if (!active) return;
        cudaStreamAttrValue attr{};
        attr.accessPolicyWindow.num_bytes = 0;
        cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);
    }
};


// ---- host-buffer classify call: pieces, slots, packed transfer ---------------------------------------------
// The batch is cut into pieces of ~8 MB of bases that flow through kSlots streams: while piece i is
// classified, piece i+1 is copied in and piece i-1's results leave.  Two ways in:
//   packed  every read of the piece fits the group-per-read window-table kernel: the host threads turn the ASCII
//           bases into the three bit planes that kernel works on (host_pack.hpp) in pinned staging memory, and only
//           those 3 bits per base cross PCIe;
//   ascii   anything else: the piece's bases are copied as they are.
// Two ways out: result arrays in pinned (or registered) host memory are written by the decode kernel directly
// (mapped, coalesced stores); otherwise results gather in a device buffer and are copied at the end.
constexpr int kSlots = 4;
constexpr uint64_t kPieceBases = 8ull << 20, kPieceReads = 1ull << 18;   // env RB_PIECE_MB overrides the former
constexpr uint64_t kPackTask = 128ull << 10;            // bases per packing task (a multiple of 32)
constexpr uint64_t kStageBases = 512ull << 20;          // bases per round of the packed pipeline (192 MB of planes)
// share of the pieces shipped as ASCII next to the packed ones (pinned input, >= 8 pieces).  0: on the measured host the
// packer and the copy engine compete for the same host-memory bandwidth, every share > 0 was slower
// (profiles/r1_k_e2e_ascii_share_sweep.jsonl); RB_ASCII_SHARE turns it on for hosts where they do not.
constexpr double kAsciiShare = 0.0;
// packed or ASCII for batches of at least kPolicyMinPieces pieces: ASCII has to be this much faster to be chosen
constexpr size_t kPolicyMinPieces = 8;
constexpr double kPolicyMargin = 0.95;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes)
    {
        if (bytes <= cap) return RB_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        const size_t want = bytes + bytes / 4 + 256;
        RB_CUDA(cudaMalloc(&p, want));
        cap = want;
        return RB_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct HostBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes)
    {
        if (bytes <= cap) return RB_OK;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        const size_t want = bytes + bytes / 4 + 256;
        RB_CUDA(cudaMallocHost(&p, want));
        cap = want;
        return RB_OK;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

}  // namespace

struct CallCtx {
    bool ok = false;
    cudaStream_t st[kSlots] = {};
    cudaEvent_t done[kSlots] = {};
    cudaEvent_t ev_start = nullptr;
    DevBuf d_off[kSlots], d_keys[kSlots], d_flag[kSlots], d_in[kSlots], d_cf[kSlots], d_cr[kSlots];
    DevBuf d_lut, d_res;
    DevBuf d_zero;                           // all-zero "not ACGT" plane for the pieces that have no such base
    HostBuf h_planes;
    std::vector<uint16_t> lut_host;          // what d_lut holds
    int init()
    {
        for (int s = 0; s < kSlots; ++s) {
            RB_CUDA(cudaStreamCreateWithFlags(&st[s], cudaStreamNonBlocking));
            RB_CUDA(cudaEventCreateWithFlags(&done[s], cudaEventDisableTiming));
        }
        RB_CUDA(cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming));
        ok = true;
        return RB_OK;
    }
    ~CallCtx()
    {
        for (int s = 0; s < kSlots; ++s) {
            if (st[s]) { cudaStreamSynchronize(st[s]); cudaStreamDestroy(st[s]); }
            if (done[s]) cudaEventDestroy(done[s]);
            d_off[s].release(); d_keys[s].release(); d_flag[s].release(); d_in[s].release(); d_cf[s].release(); d_cr[s].release();
        }
        if (ev_start) cudaEventDestroy(ev_start);
        d_lut.release(); d_res.release(); d_zero.release(); h_planes.release();
        cudaGetLastError();
    }
};

namespace {

CallCtx *acquire_ctx(const rb_ibf *f)
{
    {
        std::lock_guard<std::mutex> lock(f->ctx_mu);
        if (!f->ctx_free.empty()) {
            CallCtx *c = f->ctx_free.back();
            f->ctx_free.pop_back();
            return c;
        }
    }
    CallCtx *c = new CallCtx();
    if (c->init() != RB_OK) { delete c; cudaGetLastError(); return nullptr; }
    return c;
}

void release_ctx(const rb_ibf *f, CallCtx *c)
{
    std::lock_guard<std::mutex> lock(f->ctx_mu);
    f->ctx_free.push_back(c);
}

void free_call_contexts(rb_ibf *f)
{
    std::lock_guard<std::mutex> lock(f->ctx_mu);
    if (f->ctx_free.empty()) return;
    DeviceGuard g(f->device);
    for (CallCtx *c : f->ctx_free) delete c;
    f->ctx_free.clear();
}

// device-visible address of a host result array if it is pinned / registered, else null
template <typename T>
T *mapped_ptr(T *host)
{
    if (!host) return nullptr;
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
    return static_cast<T *>(at.devicePointer);
}

int run_count_batch(const rb_ibf *f, CallCtx *ctx, const uint8_t *bases, const uint64_t *read_off, uint64_t n_reads,
                    const uint16_t *thr_lut, uint32_t n_lut, uint16_t *counts_fwd, uint16_t *counts_rev,
                    uint16_t *max_count, uint8_t *hit, uint32_t *argmax_bin, uint8_t *read_flag, cudaStream_t user)
{
    // RB_TRACE=1: host-clock milestones of the call on stderr (where does the end-to-end time go)
    static const int trace_level = [] { const char *e = std::getenv("RB_TRACE"); return e ? std::atoi(e) : 0; }();
    const bool trace = trace_level >= 1;
    // RB_TRACE=2: additionally a device timeline per piece (CUDA events after the copies, the count and the decode kernel)
    struct PieceEv { cudaEvent_t h2d, cnt, dec; double t_submit; };
    std::vector<PieceEv> pev;
    cudaEvent_t ev_t0 = nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto ms_since = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
    double t_setup = 0, t_first_submit = -1, t_packed = 0, t_submitted = 0;
    const bool dense = counts_fwd || counts_rev;
    const uint64_t nbl = f->n_bins_local;
    const uint64_t piece_reads_cap = dense ? std::max<uint64_t>(1, std::min<uint64_t>(kPieceReads, (64ull << 20) / (2 * nbl)))
                                           : kPieceReads;
    // ---- pieces: ~kPieceBases bases or piece_reads_cap reads, whichever comes first (binary search: offsets are sorted;
    //      every piece is checked for that when it is submitted) -------------------------------------------------------
    auto env_mb = [](const char *name, uint64_t dflt) {       // test / tuning overrides, read at every call
        const char *e = std::getenv(name);
        const long v = e ? std::atol(e) : 0;
        return v > 0 ? (uint64_t)v << 20 : dflt;
    };
    const uint64_t piece_bases = env_mb("RB_PIECE_MB", kPieceBases), stage_bases = env_mb("RB_STAGE_MB", kStageBases);
    std::vector<uint64_t> cut{0};
    while (cut.back() < n_reads) {
        const uint64_t r0 = cut.back();
        const uint64_t rmax = std::min<uint64_t>(n_reads, r0 + piece_reads_cap);
        const uint64_t *lo = read_off + r0 + 1, *hi = read_off + rmax;
        const uint64_t *it = std::lower_bound(lo, hi, read_off[r0] + piece_bases);     // first read END >= r0 + piece bases
        cut.push_back(std::min<uint64_t>(rmax, (uint64_t)(it - read_off)));
        if (cut.back() <= r0) cut.back() = r0 + 1;
    }
    const size_t n_pieces = cut.size() - 1;

    // ---- which way in ----------------------------------------------------------------------------------------------------
    const int which = g_count_kernel.load();
    const uint64_t *table = nullptr;
    if (which == 0 || which >= 3) table = ensure_table(f, user, which >= 3, n_reads);
    if (which >= 3 && !table) return fail(RB_ERR_INVALID_ARG, "k-mer table not applicable to this filter (row > 4 words, k > 16 or no memory)");
    bool packed_ok = table && f->table_span >= 2 && !dense && (which == 0 || which == 3);
    int pol_measure = -1;                            // >= 0: this call's time per base decides the policy (0 packed, 1 ASCII)
    {
        const char *e = std::getenv("RB_HOST_PACK");
        if (e && e[0] == '0') packed_ok = false;
        else if (packed_ok && !(e && e[0] == '1') && n_pieces >= kPolicyMinPieces) {
            std::lock_guard<std::mutex> lock(f->pol_mu);
            if (f->pol_choice >= 0) packed_ok = f->pol_choice == 0;
            else {
                const int c = f->pol_calls++;        // calls 0, 1 packed; 2, 3 ASCII; the first of each kind warms its buffers
                const int mode = (c & 2) ? 1 : 0;
                if (c & 1) pol_measure = mode;
                packed_ok = mode == 0;
            }
        }
    }

    // ---- which way out ---------------------------------------------------------------------------------------------------
    uint16_t *m_max = mapped_ptr(max_count);
    uint8_t *m_hit = mapped_ptr(hit), *m_flag = mapped_ptr(read_flag);
    uint32_t *m_amax = mapped_ptr(argmax_bin);
    const bool zero_copy = (!max_count || m_max) && (!hit || m_hit) && (!argmax_bin || m_amax) && (!read_flag || m_flag);
    const uint64_t nk_all = (uint64_t)n_lut * n_reads;
    if (!zero_copy) {        // device staging in the caller's layout: argmax | max | hit | flag
        int s2 = ctx->d_res.reserve(nk_all * 7 + n_reads + 64);
        if (s2 != RB_OK) return s2;
        uint8_t *b = static_cast<uint8_t *>(ctx->d_res.p);
        m_amax = argmax_bin ? reinterpret_cast<uint32_t *>(b) : nullptr;
        m_max = max_count ? reinterpret_cast<uint16_t *>(b + nk_all * 4) : nullptr;
        m_hit = hit ? b + nk_all * 6 : nullptr;
        m_flag = read_flag ? b + nk_all * 7 : nullptr;
    }

    auto body = [&]() -> int {
        // thresholds: uploaded only when they changed since the last call of this context
        const size_t lut_elems = (size_t)n_lut * rb::kLutSize;
        if (ctx->lut_host.size() != lut_elems || std::memcmp(ctx->lut_host.data(), thr_lut, lut_elems * 2) != 0) {
            int s2 = ctx->d_lut.reserve(lut_elems * 2);
            if (s2 != RB_OK) return s2;
            ctx->lut_host.assign(thr_lut, thr_lut + lut_elems);
            RB_CUDA(cudaMemcpyAsync(ctx->d_lut.p, ctx->lut_host.data(), lut_elems * 2, cudaMemcpyHostToDevice, user));
            g_h2d_bytes += lut_elems * 2;
        }
        if (trace_level >= 2) { cudaEventCreate(&ev_t0); cudaEventRecord(ev_t0, user); }
        RB_CUDA(cudaEventRecord(ctx->ev_start, user));          // slot streams start after the caller's stream
        for (int s = 0; s < kSlots; ++s) RB_CUDA(cudaStreamWaitEvent(ctx->st[s], ctx->ev_start, 0));

        // one piece: offsets in, classify (packed planes already on their way, or ASCII bases copied here), results out
        t_setup = ms_since();
        auto submit = [&](size_t p, const uint32_t *d_lo, const uint32_t *d_hi, const uint32_t *d_bad, uint64_t base0) -> int {
            if (t_first_submit < 0) t_first_submit = ms_since();
            const int s = (int)(p % kSlots);
            cudaStream_t sst = ctx->st[s];
            const uint64_t r0 = cut[p], r1 = cut[p + 1], n = r1 - r0;
            const uint64_t piece_max = rb::max_read_length(read_off + r0, n);
            if (piece_max >> 63) return fail(RB_ERR_INVALID_ARG, "read offsets must be non-decreasing");
            const uint64_t b0 = read_off[r0], nb = read_off[r1] - b0;
            const uint32_t max_len32 = (uint32_t)std::min<uint64_t>(piece_max, 0xFFFFFFFFu);
            const bool packed = d_lo && rb::wgroup_applicable((uint32_t)f->k, f->table_span, max_len32);
            int s2;
            if ((s2 = ctx->d_off[s].reserve((n + 1) * 8)) != RB_OK) return s2;
            if ((s2 = ctx->d_keys[s].reserve((uint64_t)n_lut * n * 8)) != RB_OK) return s2;
            if ((s2 = ctx->d_flag[s].reserve(n)) != RB_OK) return s2;
            RB_CUDA(cudaMemcpyAsync(ctx->d_off[s].p, read_off + r0, (n + 1) * 8, cudaMemcpyHostToDevice, sst));
            g_h2d_bytes += (n + 1) * 8;
            PieceEv pe{};
            if (trace_level >= 2) {
                cudaEventCreate(&pe.h2d); cudaEventCreate(&pe.cnt); cudaEventCreate(&pe.dec);
                pe.t_submit = ms_since();
            }
            rb::CountArgs a{};
            a.fv = view_of(f);
            a.read_off = static_cast<const uint64_t *>(ctx->d_off[s].p);
            a.n_reads = n; a.lut = static_cast<const uint16_t *>(ctx->d_lut.p); a.n_lut = n_lut;
            a.keys = static_cast<uint64_t *>(ctx->d_keys[s].p);
            a.read_flag = static_cast<uint8_t *>(ctx->d_flag[s].p);
            if (packed) {
                a.pk_lo = d_lo; a.pk_hi = d_hi; a.pk_bad = d_bad; a.pk_base0 = base0;
                if (trace_level >= 2) cudaEventRecord(pe.h2d, sst);
                int nl = rb::launch_count_wtable(a, table, f->table_span, max_len32, f->sm_count, sst);
                if (nl < 0) return fail(RB_ERR_COUNT_KMER, std::string("count launch failed: ") + cudaGetErrorString(cudaGetLastError()));
                g_launches += (uint64_t)nl;
            } else {
                if ((s2 = ctx->d_in[s].reserve(nb ? nb : 1)) != RB_OK) return s2;
                RB_CUDA(cudaMemcpyAsync(ctx->d_in[s].p, bases + b0, nb, cudaMemcpyHostToDevice, sst));
                g_h2d_bytes += nb;
                uint16_t *d_cf = nullptr, *d_cr = nullptr;
                if (counts_fwd) { if ((s2 = ctx->d_cf[s].reserve(n * nbl * 2)) != RB_OK) return s2; d_cf = static_cast<uint16_t *>(ctx->d_cf[s].p); }
                if (counts_rev) { if ((s2 = ctx->d_cr[s].reserve(n * nbl * 2)) != RB_OK) return s2; d_cr = static_cast<uint16_t *>(ctx->d_cr[s].p); }
                // offsets stay absolute: bias the base pointer instead of rewriting them
                const uint8_t *biased = reinterpret_cast<const uint8_t *>(reinterpret_cast<uintptr_t>(ctx->d_in[s].p) - (uintptr_t)b0);
                s2 = rb_ibf_count_batch_dev(f, biased, a.read_off, n, max_len32, a.lut, n_lut, a.keys, d_cf, d_cr, a.read_flag, sst);
                if (s2 != RB_OK) return s2;
                if (counts_fwd) RB_CUDA(cudaMemcpyAsync(counts_fwd + r0 * nbl, d_cf, n * nbl * 2, cudaMemcpyDeviceToHost, sst));
                if (counts_rev) RB_CUDA(cudaMemcpyAsync(counts_rev + r0 * nbl, d_cr, n * nbl * 2, cudaMemcpyDeviceToHost, sst));
            }
            if (trace_level >= 2) cudaEventRecord(pe.cnt, sst);
            int nl = rb::launch_keys_decode_piece(a.keys, a.read_flag, n, n_lut, n_reads, r0, m_max, m_hit, m_amax, m_flag, sst);
            if (nl < 0) return fail(RB_ERR_CUDA, "decode launch failed");
            g_launches += (uint64_t)nl;
            if (trace_level >= 2) { cudaEventRecord(pe.dec, sst); pev.push_back(pe); }
            return RB_OK;
        };

        if (!packed_ok) {
            for (size_t p = 0; p < n_pieces; ++p) { int s2 = submit(p, nullptr, nullptr, nullptr, 0); if (s2 != RB_OK) return s2; }
        } else {
            // rounds of at most kStageBases bases; inside a round the host threads pack 128 K-base tasks in order and the
            // calling thread submits every piece as soon as its tasks are done.  Every piece has its own region of the
            // pinned staging buffer -- [low plane][high plane][bad plane], bit i = base base0 + i -- so it crosses PCIe
            // as ONE copy.
            //
            // Mixed transfer: the packer (host cores) and the PCIe link are separate resources, so when the caller's bases
            // are in pinned memory (the copy engine reads them without any host work) a share of the pieces is shipped as
            // ASCII while the host threads pack the others; the ASCII pieces are queued first, so the link is busy from
            // the start of the round.  Off by default (kAsciiShare); RB_ASCII_SHARE sets the share.
            struct Region { uint64_t base0, nb, nw; size_t piece, h_word, task0, n_tasks; };
            double ascii_share = 0.0;
            {
                cudaPointerAttributes at{};
                const bool pinned_in = cudaPointerGetAttributes(&at, bases) == cudaSuccess && at.type == cudaMemoryTypeHost;
                cudaGetLastError();
                if (pinned_in && n_pieces >= 8) {
                    ascii_share = kAsciiShare;
                    if (const char *e = std::getenv("RB_ASCII_SHARE")) ascii_share = std::min(1.0, std::max(0.0, std::atof(e)));
                }
            }
            double ascii_acc = 0.0;
            size_t p_begin = 0;
            while (p_begin < n_pieces) {
                size_t p_end = p_begin + 1;
                const uint64_t R0 = read_off[cut[p_begin]];
                while (p_end < n_pieces && read_off[cut[p_end + 1]] - R0 <= stage_bases) ++p_end;
                std::vector<Region> reg;
                reg.reserve(p_end - p_begin);
                std::vector<uint32_t> task_piece;
                size_t h_words = 0;
                for (size_t p = p_begin; p < p_end; ++p) {
                    ascii_acc += ascii_share;
                    if (ascii_acc >= 1.0) {                                   // this piece goes as it is, right now
                        ascii_acc -= 1.0;
                        int s3 = submit(p, nullptr, nullptr, nullptr, 0);
                        if (s3 != RB_OK) return s3;
                        continue;
                    }
                    Region r{};
                    const uint64_t b0 = read_off[cut[p]], b1 = read_off[cut[p + 1]];
                    r.piece = p;
                    r.base0 = b0 - ((b0 - R0) & 31u);                        // word aligned, never before the round's first base
                    r.nb = b1 - r.base0;
                    r.nw = ((r.nb + 31) / 32 + 2 + 15) / 16 * 16;             // 64-byte multiples: streaming stores, aligned planes
                    r.h_word = h_words;
                    r.task0 = task_piece.size();
                    r.n_tasks = (size_t)std::max<uint64_t>(1, (r.nb + kPackTask - 1) / kPackTask);
                    task_piece.insert(task_piece.end(), r.n_tasks, (uint32_t)reg.size());
                    h_words += 3 * r.nw;
                    reg.push_back(r);
                }
                int s2;
                if ((s2 = ctx->h_planes.reserve(h_words * 4 + 64)) != RB_OK) return s2;
                uint32_t *const h_base = static_cast<uint32_t *>(ctx->h_planes.p);
                const size_t n_tasks = task_piece.size();
                size_t next_reg = 0;
                int err = RB_OK;
                // The "not ACGT" plane of a task is written only if the task has such a base (task_bad); a piece none of whose
                // tasks has one ships two planes and classifies against a zero plane that lives on the device.
                std::vector<uint8_t> task_bad(n_tasks, 0);
                uint64_t max_nw = 0;
                for (const Region &r : reg) max_nw = std::max(max_nw, r.nw);
                if (ctx->d_zero.cap < max_nw * 4) {
                    if ((s2 = ctx->d_zero.reserve(max_nw * 4)) != RB_OK) return s2;
                    RB_CUDA(cudaMemsetAsync(ctx->d_zero.p, 0, ctx->d_zero.cap, user));     // the slot streams start after `user`
                    RB_CUDA(cudaEventRecord(ctx->ev_start, user));
                    for (int s = 0; s < kSlots; ++s) RB_CUDA(cudaStreamWaitEvent(ctx->st[s], ctx->ev_start, 0));
                }
                auto pack_task = [&](size_t t) {
                    const Region &r = reg[task_piece[t]];
                    const uint64_t o = (uint64_t)(t - r.task0) * kPackTask;
                    if (o >= r.nb) return;
                    const uint64_t m = std::min<uint64_t>(kPackTask, r.nb - o);
                    uint32_t *h = h_base + r.h_word + o / 32;
                    task_bad[t] = rb::pack_bases_lazy(bases + r.base0 + o, m, h, h + r.nw, h + 2 * r.nw, true) ? 1 : 0;
                };
                auto poll = [&](size_t tasks_done) {
                    while (err == RB_OK && next_reg < reg.size()) {
                        const Region &r = reg[next_reg];
                        if (tasks_done < r.task0 + r.n_tasks) break;
                        const int s = (int)(r.piece % kSlots);
                        if ((err = ctx->d_in[s].reserve(r.nw * 12)) != RB_OK) break;
                        uint32_t *d = static_cast<uint32_t *>(ctx->d_in[s].p);
                        bool any_bad = false;
                        for (size_t t = r.task0; t < r.task0 + r.n_tasks; ++t) any_bad = any_bad || task_bad[t];
                        if (any_bad) {                     // rare: the clean tasks of this piece left their share of the plane unwritten
                            for (size_t t = r.task0; t < r.task0 + r.n_tasks; ++t) {
                                const uint64_t o = (uint64_t)(t - r.task0) * kPackTask;
                                if (task_bad[t] || o >= r.nb) continue;
                                const uint64_t m = std::min<uint64_t>(kPackTask, r.nb - o);
                                std::memset(h_base + r.h_word + 2 * r.nw + o / 32, 0, (m + 31) / 32 * 4);
                            }
                        }
                        const size_t copy_bytes = r.nw * (any_bad ? 12 : 8);
                        cudaError_t ce = cudaMemcpyAsync(d, h_base + r.h_word, copy_bytes, cudaMemcpyHostToDevice, ctx->st[s]);
                        if (ce != cudaSuccess) { err = fail(RB_ERR_CUDA, std::string("plane copy: ") + cudaGetErrorString(ce)); break; }
                        g_h2d_bytes += copy_bytes;
                        err = submit(r.piece, d, d + r.nw, any_bad ? d + 2 * r.nw : static_cast<const uint32_t *>(ctx->d_zero.p), r.base0);
                        ++next_reg;
                    }
                };
                if (n_tasks) rb::parallel_tasks(n_tasks, pack_task, poll);
                t_packed = ms_since();
                poll(n_tasks);
                if (err != RB_OK) return err;
                if (p_end < n_pieces) {            // the staging buffer is reused by the next round
                    for (int s = 0; s < kSlots; ++s) RB_CUDA(cudaStreamSynchronize(ctx->st[s]));
                }
                p_begin = p_end;
            }
        }
        t_submitted = ms_since();
        for (int s = 0; s < kSlots; ++s) {
            RB_CUDA(cudaEventRecord(ctx->done[s], ctx->st[s]));
            RB_CUDA(cudaStreamWaitEvent(user, ctx->done[s], 0));
        }
        if (!zero_copy) {
            if (argmax_bin) RB_CUDA(cudaMemcpyAsync(argmax_bin, m_amax, nk_all * 4, cudaMemcpyDeviceToHost, user));
            if (max_count) RB_CUDA(cudaMemcpyAsync(max_count, m_max, nk_all * 2, cudaMemcpyDeviceToHost, user));
            if (hit) RB_CUDA(cudaMemcpyAsync(hit, m_hit, nk_all, cudaMemcpyDeviceToHost, user));
            if (read_flag) RB_CUDA(cudaMemcpyAsync(read_flag, m_flag, n_reads, cudaMemcpyDeviceToHost, user));
        }
        g_d2h_bytes += (argmax_bin ? nk_all * 4 : 0) + (max_count ? nk_all * 2 : 0) + (hit ? nk_all : 0) + (read_flag ? n_reads : 0) +
                       ((counts_fwd ? 1 : 0) + (counts_rev ? 1 : 0)) * n_reads * nbl * 2;
        cudaError_t e = cudaStreamSynchronize(user);
        if (e != cudaSuccess) return fail(RB_ERR_COUNT_KMER, std::string("Error counting kmers in IBF bins: ") + cudaGetErrorString(e));
        if (trace_level >= 2 && ev_t0) {
            for (size_t i = 0; i < pev.size(); ++i) {
                float a1 = 0, a2 = 0, a3 = 0;
                cudaEventElapsedTime(&a1, ev_t0, pev[i].h2d); cudaEventElapsedTime(&a2, ev_t0, pev[i].cnt); cudaEventElapsedTime(&a3, ev_t0, pev[i].dec);
                std::fprintf(stderr, "[rb piece %2zu] submit %.3f | h2d done %.3f count done %.3f decode done %.3f ms\n", i, pev[i].t_submit, a1, a2, a3);
                cudaEventDestroy(pev[i].h2d); cudaEventDestroy(pev[i].cnt); cudaEventDestroy(pev[i].dec);
            }
            cudaEventDestroy(ev_t0);
        }
        if (trace)
            std::fprintf(stderr, "[rb trace] reads %llu pieces %zu packed %d zero_copy %d | setup %.3f first_submit %.3f packed %.3f "
                                 "submitted %.3f done %.3f ms\n", (unsigned long long)n_reads, n_pieces, (int)packed_ok, (int)zero_copy,
                         t_setup, t_first_submit, t_packed, t_submitted, ms_since());
        return RB_OK;
    };
    int st = body();
    if (st == RB_OK && pol_measure >= 0 && read_off[n_reads] > read_off[0]) {
        std::lock_guard<std::mutex> lock(f->pol_mu);
        f->pol_ns_per_base[pol_measure] = 1e6 * ms_since() / (double)(read_off[n_reads] - read_off[0]);
        if (f->pol_choice < 0 && f->pol_ns_per_base[0] > 0 && f->pol_ns_per_base[1] > 0)
            f->pol_choice = f->pol_ns_per_base[1] < kPolicyMargin * f->pol_ns_per_base[0] ? 1 : 0;
    }
    if (st != RB_OK) {                       // drain whatever was enqueued before the buffers are reused
        const std::string keep = g_last_error;
        for (int s = 0; s < kSlots; ++s) cudaStreamSynchronize(ctx->st[s]);
        cudaStreamSynchronize(user);
        cudaGetLastError();
        g_last_error = keep;
    }
    return st;
}

}  // namespace

// =============================================================================================
extern "C" {

const char *rb_status_string(int status)
{
    switch (status) {
    case RB_OK: return "ok";
    case RB_ERR_NULL_FILTER: return "NullFilterException";
    case RB_ERR_SHORT_READ: return "ShortReadException";
    case RB_ERR_COUNT_KMER: return "CountKmerException";
    case RB_ERR_PARSE_IBF_FILE: return "ParseIBFFileException";
    case RB_ERR_MISSING_IBF_FILE: return "MissingIBFFileException";
    case RB_ERR_STORE_FILTER: return "StoreFilterException";
    case RB_ERR_INSERT_SEQUENCE: return "InsertSequenceException";
    case RB_ERR_INVALID_CONFIG: return "InvalidConfigException";
    case RB_ERR_ALLOC: return "out of memory";
    case RB_ERR_CUDA: return "CUDA error";
    case RB_ERR_NO_DEVICE: return "no CUDA device";
    case RB_ERR_INVALID_ARG: return "invalid argument";
    default: return "unknown status";
    }
}

const char *rb_last_error(void) { return g_last_error.c_str(); }

int rb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

uint64_t rb_kernel_launches(void) { return g_launches.load(); }

int rb_set_l2_fetch_granularity(int device, uint32_t bytes)
{
    if (bytes != 32 && bytes != 64 && bytes != 128) return fail(RB_ERR_INVALID_ARG, "granularity must be 32, 64 or 128");
    int st = check_device(device);
    if (st != RB_OK) return st;
    DeviceGuard g(device);
    RB_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, bytes));
    return RB_OK;
}

int rb_get_l2_fetch_granularity(int device, uint32_t *bytes)
{
    int st = check_device(device);
    if (st != RB_OK) return st;
    DeviceGuard g(device);
    size_t v = 0;
    RB_CUDA(cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity));
    if (bytes) *bytes = (uint32_t)v;
    return RB_OK;
}

int rb_set_count_kernel(int which)
{
    if (which < 0 || which > 5) return fail(RB_ERR_INVALID_ARG, "kernel selector must be 0..5");
    g_count_kernel.store(which);
    rb::set_wtable_variant(which == 5 ? 1 : 0);
    return RB_OK;
}

int rb_set_insert_kernel(int which)
{
    if (which < 0 || which > 2) return fail(RB_ERR_INVALID_ARG, "insert kernel selector must be 0..2");
    rb::set_insert_variant(which);
    return RB_OK;
}

// ---- scalar helpers ----------------------------------------------------------------------------
uint64_t rb_ibf_size_bits(uint64_t fragment_length, uint32_t kmer_size, uint32_t n_hash, double max_fp,
                          uint64_t n_bins)
{
    // IBF::calculate_filter_size_bits, src/IBF/IBFBuild.cpp:404-413
    const uint64_t max_kmer_count = fragment_length - kmer_size + 1;
    const uint64_t optimal_bins = (uint64_t)(std::floor(((double)n_bins / 64.0) + 1) * 64);
    const uint64_t bin_size_bits = (uint64_t)std::ceil(
        -1 / (std::pow(1 - std::pow(max_fp, 1.0 / (double)n_hash), 1.0 / ((double)(n_hash * max_kmer_count))) - 1));
    return bin_size_bits * optimal_bins;
}

int rb_calculate_ci(double error_rate, uint32_t kmer_size, uint32_t readlen, double significance, uint16_t *low,
                    uint16_t *high)
{
    if (!(error_rate > 0.0 && error_rate < 1.0) || !(significance > 0.0 && significance < 1.0))
        return fail(RB_ERR_INVALID_CONFIG, "error rate and significance must be in (0, 1)");
    calculate_ci(error_rate, kmer_size, readlen, significance, low, high);
    return RB_OK;
}

int rb_threshold_lut(double error_rate, double significance, uint32_t kmer_size, uint16_t *lut65536)
{
    if (!lut65536) return fail(RB_ERR_INVALID_ARG, "null lut");
    if (!(error_rate > 0.0 && error_rate < 1.0) || !(significance > 0.0 && significance < 1.0))
        return fail(RB_ERR_INVALID_CONFIG, "error rate and significance must be in (0, 1)");
    return rb_threshold_lut_raw(error_rate, significance, kmer_size, lut65536);
}

int rb_threshold_lut_raw(double error_rate, double significance, uint32_t kmer_size, uint16_t *lut65536)
{
    if (!lut65536) return fail(RB_ERR_INVALID_ARG, "null lut");
    for (uint32_t len = 0; len < 65536; ++len) {
        // src/IBF/IBFClassify.cpp:154-159: uint16 readlen, int16 threshold, used as uint16
        uint16_t hi = 0;
        calculate_ci(error_rate, kmer_size, len, significance, nullptr, &hi);
        const int16_t thr = (int16_t)((int)(uint16_t)len - (int)kmer_size + 1 - (int)hi);
        lut65536[len] = (uint16_t)thr;
    }
    return RB_OK;
}

uint64_t rb_cut_out_nnns(const char *seq, uint64_t len, char *out)
{
    // IBF::cutOutNNNs, src/IBF/IBFBuild.cpp:112-132: pieces between runs of 'N' are kept; the
    // last piece loses its final base when the sequence does not end in 'N'.
    uint64_t n = 0, pos = 0;
    while (pos < len) {
        while (pos < len && seq[pos] == 'N') ++pos;
        if (pos >= len) break;
        uint64_t stop = pos;
        while (stop < len && seq[stop] != 'N') ++stop;
        uint64_t take = stop < len ? stop - pos : len - pos - 1;
        std::memcpy(out + n, seq + pos, take);
        n += take;
        pos = stop;
    }
    return n;
}

uint64_t rb_fragment_schedule(uint64_t seqlen, uint64_t fragment_length, uint32_t kmer_size, uint64_t *begin,
                              uint64_t *end, uint64_t cap)
{
    // src/IBF/IBFBuild.cpp:165-202
    uint64_t n = 0;
    int64_t start = 0;
    for (int64_t idx = 0; start < (int64_t)seqlen - 1; ++idx) {
        uint64_t stop = std::min<uint64_t>((uint64_t)(idx + 1) * fragment_length, seqlen);
        if (n < cap) {
            if (begin) begin[n] = (uint64_t)start;
            if (end) end[n] = stop;
        }
        ++n;
        start = (idx + 1) * (int64_t)fragment_length - (int64_t)kmer_size + 1;
    }
    return n;
}

// ---- life cycle --------------------------------------------------------------------------------
rb_ibf *rb_ibf_create(uint64_t n_bins, uint32_t n_hash, uint32_t kmer_size, uint64_t n_bits, int device, int *status)
{
    int st = check_device(device);
    rb_ibf *f = nullptr;
    if (st == RB_OK) {
        f = new rb_ibf();
        f->device = device;
        f->n_bins = n_bins; f->n_hash = n_hash; f->k = kmer_size; f->n_bits = n_bits;
        DeviceGuard g(device);
        st = derive_geometry(f, 0, 1);
        if (st == RB_OK) st = alloc_device(f, true);
        if (st != RB_OK) { destroy(f); f = nullptr; }
    }
    if (status) *status = st;
    return f;
}

rb_ibf *rb_ibf_create_shard(uint64_t n_bins, uint32_t n_hash, uint32_t kmer_size, uint64_t n_bits, int device, int shard,
                            int n_shards, int *status)
{
    int st = check_device(device);
    rb_ibf *f = nullptr;
    if (st == RB_OK) {
        f = new rb_ibf();
        f->device = device;
        f->n_bins = n_bins; f->n_hash = n_hash; f->k = kmer_size; f->n_bits = n_bits;
        DeviceGuard g(device);
        st = derive_geometry(f, shard, n_shards);
        if (st == RB_OK) st = alloc_device(f, true);
        if (st != RB_OK) { destroy(f); f = nullptr; }
    }
    if (status) *status = st;
    return f;
}

rb_ibf *rb_ibf_from_words(const uint64_t *words, uint64_t n_bins, uint32_t n_hash, uint32_t kmer_size, uint64_t n_bits,
                          int device, int shard, int n_shards, int *status)
{
    int st = words ? check_device(device) : fail(RB_ERR_INVALID_ARG, "null words");
    rb_ibf *f = nullptr;
    if (st == RB_OK) {
        f = new rb_ibf();
        f->device = device;
        f->n_bins = n_bins; f->n_hash = n_hash; f->k = kmer_size; f->n_bits = n_bits;
        DeviceGuard g(device);
        st = derive_geometry(f, shard, n_shards);
        if (st == RB_OK) st = alloc_device(f, false);
        if (st == RB_OK) {
            auto body = [&]() -> int {
                if (n_shards == 1) RB_CUDA(cudaMemcpy(f->d_words, words, f->n_local_words * 8, cudaMemcpyHostToDevice));
                else {
                    int s2 = upload_rows(f, words, 0, f->n_blocks, nullptr);
                    if (s2 != RB_OK) return s2;
                    RB_CUDA(cudaStreamSynchronize(nullptr));
                }
                return RB_OK;
            };
            st = body();
        }
        if (st != RB_OK) { destroy(f); f = nullptr; }
    }
    if (status) *status = st;
    return f;
}

rb_ibf *rb_ibf_load_shard(const char *path, int device, int shard, int n_shards, int *status)
{
    rb_ibf *f = nullptr;
    FILE *fp = nullptr;
    auto body = [&]() -> int {
        if (!path || !*path) return fail(RB_ERR_MISSING_IBF_FILE, "either update_filter_file or input_filter_file has to be specified");
        struct stat sb;
        if (stat(path, &sb) != 0 || !S_ISREG(sb.st_mode)) return fail(RB_ERR_MISSING_IBF_FILE, std::string("cannot open IBF file ") + path);
        fp = std::fopen(path, "rb");
        if (!fp) return fail(RB_ERR_MISSING_IBF_FILE, std::string("cannot open IBF file ") + path);
        // sdsl bit_vector::serialize: u64 bit length, then ceil(len/64) words
        uint64_t bit_len = 0;
        const uint64_t fsize = (uint64_t)sb.st_size;
        if (fsize < 8 + 32 + 8 || std::fread(&bit_len, 8, 1, fp) != 1) return fail(RB_ERR_PARSE_IBF_FILE, "file too small for an IBF");
        if (bit_len < 256 + 64 || (bit_len % 64) != 0 || fsize != 8 + bit_len / 8)
            return fail(RB_ERR_PARSE_IBF_FILE, "sdsl bit_vector header does not match the file size");
        const uint64_t n_bits = bit_len - 256;
        uint64_t tail[4];
        if (fseeko(fp, (off_t)(8 + n_bits / 8), SEEK_SET) != 0 || std::fread(tail, 8, 4, fp) != 4)
            return fail(RB_ERR_PARSE_IBF_FILE, "cannot read the metadata tail");
        if (tail[0] == 0 || tail[1] == 0 || tail[1] > (uint64_t)rb::kMaxHash || tail[2] == 0 || tail[2] > 32 || tail[3] != tail[2])
            return fail(RB_ERR_PARSE_IBF_FILE, "metadata tail is not [bins, hash functions, k, k]");
        int st = check_device(device);
        if (st != RB_OK) return st;
        f = new rb_ibf();
        f->device = device;
        f->n_bins = tail[0]; f->n_hash = tail[1]; f->k = tail[2]; f->n_bits = n_bits;
        st = derive_geometry(f, shard, n_shards);
        if (st == RB_ERR_INVALID_CONFIG) return fail(RB_ERR_PARSE_IBF_FILE, std::string(g_last_error));
        if (st != RB_OK) return st;
        st = alloc_device(f, false);
        if (st != RB_OK) return st;
        if (fseeko(fp, 8, SEEK_SET) != 0) return fail(RB_ERR_PARSE_IBF_FILE, "seek failed");
        // stream the payload through two pinned staging buffers so disk reads overlap the H2D copies
        const uint64_t row_words = f->bin_width;
        uint64_t rows_per_stage = std::max<uint64_t>(1, kStageBytes / 8 / row_words);
        PinnedPair pp;
        st = pp.init((size_t)(n_shards == 1 ? kStageBytes / 8 : rows_per_stage * row_words));
        if (st != RB_OK) return st;
        int cur = 0;
        if (n_shards == 1) {
            const uint64_t total = n_bits / 64;
            for (uint64_t w0 = 0; w0 < total; w0 += pp.words, cur ^= 1) {
                const uint64_t nw = std::min<uint64_t>(pp.words, total - w0);
                RB_CUDA(cudaEventSynchronize(pp.ev[cur]));
                if (std::fread(pp.buf[cur], 8, nw, fp) != nw) return fail(RB_ERR_PARSE_IBF_FILE, "short read of the bit matrix");
                RB_CUDA(cudaMemcpyAsync(f->d_words + w0, pp.buf[cur], nw * 8, cudaMemcpyHostToDevice, pp.st));
                RB_CUDA(cudaEventRecord(pp.ev[cur], pp.st));
            }
        } else {
            for (uint64_t r0 = 0; r0 < f->n_blocks; r0 += rows_per_stage, cur ^= 1) {
                const uint64_t nr = std::min<uint64_t>(rows_per_stage, f->n_blocks - r0);
                RB_CUDA(cudaEventSynchronize(pp.ev[cur]));
                if (std::fread(pp.buf[cur], 8, nr * row_words, fp) != nr * row_words) return fail(RB_ERR_PARSE_IBF_FILE, "short read of the bit matrix");
                st = upload_rows(f, pp.buf[cur], r0, nr, pp.st);
                if (st != RB_OK) return st;
                RB_CUDA(cudaEventRecord(pp.ev[cur], pp.st));
            }
        }
        RB_CUDA(cudaStreamSynchronize(pp.st));
        return RB_OK;
    };
    int prev = -1;
    cudaGetDevice(&prev);
    int st = body();
    if (prev >= 0) cudaSetDevice(prev);
    cudaGetLastError();
    if (fp) std::fclose(fp);
    if (st != RB_OK) { destroy(f); f = nullptr; }
    if (status) *status = st;
    return f;
}

rb_ibf *rb_ibf_load(const char *path, int device, int *status) { return rb_ibf_load_shard(path, device, 0, 1, status); }

int rb_ibf_download(const rb_ibf *f, uint64_t *words, uint64_t n_words)
{
    if (!f) return fail(RB_ERR_NULL_FILTER, "null filter");
    if (!words || n_words != f->n_local_words) return fail(RB_ERR_INVALID_ARG, "word count does not match the local matrix");
    DeviceGuard g(f->device);
    RB_CUDA(cudaDeviceSynchronize());
    int st = sticky_insert_error(f);
    if (st != RB_OK) return st;
    RB_CUDA(cudaMemcpy(words, f->d_words, n_words * 8, cudaMemcpyDeviceToHost));
    return RB_OK;
}

int rb_ibf_store(const rb_ibf *f, const char *path)
{
    if (!f) return fail(RB_ERR_NULL_FILTER, "null filter");
    if (f->n_shards != 1) return fail(RB_ERR_STORE_FILTER, "a bin shard cannot be stored as a whole filter");
    if (!path || !*path) return fail(RB_ERR_STORE_FILTER, "no output filter file given");
    DeviceGuard g(f->device);
    RB_CUDA(cudaDeviceSynchronize());
    int st = sticky_insert_error(f);
    if (st != RB_OK) return st;
    FILE *fp = std::fopen(path, "wb");
    if (!fp) return fail(RB_ERR_STORE_FILTER, std::string("could not store IBF to ") + path);
    auto body = [&]() -> int {
        const uint64_t bit_len = f->n_bits + 256;
        if (std::fwrite(&bit_len, 8, 1, fp) != 1) return fail(RB_ERR_STORE_FILTER, "write failed");
        PinnedPair pp;
        int s2 = pp.init(kStageBytes / 8);
        if (s2 != RB_OK) return s2;
        const uint64_t total = f->n_bits / 64;
        // D2H of stage i+1 overlaps the fwrite of stage i
        uint64_t issued = 0, written = 0;
        int cur = 0;
        uint64_t pending[2] = {0, 0};
        auto issue = [&](int b) -> int {
            const uint64_t nw = std::min<uint64_t>(pp.words, total - issued);
            RB_CUDA(cudaMemcpyAsync(pp.buf[b], f->d_words + issued, nw * 8, cudaMemcpyDeviceToHost, pp.st));
            RB_CUDA(cudaEventRecord(pp.ev[b], pp.st));
            pending[b] = nw;
            issued += nw;
            return RB_OK;
        };
        if (total) { s2 = issue(0); if (s2 != RB_OK) return s2; }
        while (written < total) {
            if (issued < total) { s2 = issue(cur ^ 1); if (s2 != RB_OK) return s2; }
            RB_CUDA(cudaEventSynchronize(pp.ev[cur]));
            if (std::fwrite(pp.buf[cur], 8, pending[cur], fp) != pending[cur]) return fail(RB_ERR_STORE_FILTER, "write failed");
            written += pending[cur];
            cur ^= 1;
        }
        const uint64_t tail[4] = {f->n_bins, f->n_hash, f->k, f->k};
        if (std::fwrite(tail, 8, 4, fp) != 4) return fail(RB_ERR_STORE_FILTER, "write failed");
        return RB_OK;
    };
    st = body();
    if (std::fclose(fp) != 0 && st == RB_OK) st = fail(RB_ERR_STORE_FILTER, "close failed");
    return st;
}

void rb_ibf_free(rb_ibf *f) { destroy(f); }

int rb_ibf_info(const rb_ibf *f, rb_ibf_info_t *out)
{
    if (!f) return fail(RB_ERR_NULL_FILTER, "null filter");
    if (!out) return fail(RB_ERR_INVALID_ARG, "null out");
    out->n_bins = f->n_bins; out->n_hash = f->n_hash; out->kmer_size = f->k; out->n_bits = f->n_bits;
    out->bin_width = f->bin_width; out->n_blocks = f->n_blocks; out->col_begin = f->col_begin;
    out->col_words = f->col_words; out->bin_begin = 64 * f->col_begin; out->n_bins_local = f->n_bins_local;
    out->device_bytes = f->n_local_words * 8; out->device = f->device; out->shard = f->shard; out->n_shards = f->n_shards;
    {
        std::lock_guard<std::mutex> lock(f->table_mu);
        out->kmer_table_bytes = f->table_kind ? f->table_bytes : 0;
        out->kmer_table_span = f->table_kind ? f->table_span : 0;
        out->kmer_table_kind = f->table_kind;
    }
    return RB_OK;
}

int rb_ibf_enable_kmer_table(rb_ibf *f, uint64_t max_table_bytes, rb_stream stream)
{
    if (!f) return fail(RB_ERR_NULL_FILTER, "null filter");
    DeviceGuard g(f->device);
    if (max_table_bytes == UINT64_MAX) { drop_table(f); f->table_tried = true; return RB_OK; }   // disable
    drop_table(f);
    f->table_budget = max_table_bytes;
    if (!ensure_table(f, (cudaStream_t)stream, true))
        return fail(RB_ERR_INVALID_ARG, "k-mer table not applicable (k > 16, more than 65520 bins for postings) or over the memory budget");
    return RB_OK;
}

// Joint table plan for several filters (the reference classifies every read against ALL target and depletion filters:
// classify.hpp:142, adaptive_sampling.hpp:555), so that the first filter does not take the HBM the others need.
int rb_ibf_enable_kmer_tables(rb_ibf *const *filters, uint32_t n_filters, uint64_t total_bytes, rb_stream stream)
{
    if (!filters && n_filters) return fail(RB_ERR_INVALID_ARG, "null filter list");
    for (uint32_t i = 0; i < n_filters; ++i)
        if (!filters[i]) return fail(RB_ERR_NULL_FILTER, "null filter in the list");
    std::vector<int> devices;
    for (uint32_t i = 0; i < n_filters; ++i)
        if (std::find(devices.begin(), devices.end(), filters[i]->device) == devices.end()) devices.push_back(filters[i]->device);
    for (int dev : devices) {
        DeviceGuard g(dev);
        std::vector<rb_ibf *> fs;
        for (uint32_t i = 0; i < n_filters; ++i)
            if (filters[i]->device == dev && std::find(fs.begin(), fs.end(), filters[i]) == fs.end()) fs.push_back(filters[i]);
        for (rb_ibf *f : fs) {                                    // their old tables and build scratch count as free
            drop_table(f);
            if (f->build_scratch.p) { cudaFree(f->build_scratch.p); f->build_scratch = rb::ScratchBuf{}; }
        }
        size_t free_b = 0, total_b = 0;
        RB_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const uint64_t budget = total_bytes ? total_bytes : (uint64_t)(free_b * 0.85);
        struct Opt { uint64_t bytes; int span; };
        std::vector<std::vector<Opt>> opts(fs.size());
        for (size_t i = 0; i < fs.size(); ++i) {
            rb_ibf *f = fs[i];
            if (f->col_words > 4) {
                const uint64_t est = postings_estimate_bytes(f, (cudaStream_t)stream);
                if (est) opts[i].push_back({est + est / 16, 0});  // the estimate comes from a sample: leave a margin
                // rows of 5..16 words: the group-loaded k-mer table is the widening step after the postings lists
                if (const uint64_t ct = ctable_bytes(f, (cudaStream_t)stream); ct && (opts[i].empty() || ct > opts[i][0].bytes)) opts[i].push_back({ct, 1});
            } else {
                for (int sp = 1; sp <= (f->col_words <= 2 ? 4 : 1); ++sp) {
                    const uint64_t ct = sp == 1 ? ctable_bytes(f) : 0;           // rows of 3-4 words: padded, group-loaded entries
                    const uint64_t b = ct ? ct : table_bytes_needed(f, sp);
                    if (b) opts[i].push_back({b, sp});
                }
            }
        }
        std::vector<int> chosen(fs.size(), -1);
        uint64_t used = 0;
        for (size_t i = 0; i < fs.size(); ++i)                    // everybody gets the smallest table first ...
            if (!opts[i].empty() && used + opts[i][0].bytes <= budget) { chosen[i] = 0; used += opts[i][0].bytes; }
        for (;;) {                                                // ... then the narrowest span is widened while it fits
            int best = -1;
            for (size_t i = 0; i < fs.size(); ++i) {
                if (chosen[i] < 0 || chosen[i] + 1 >= (int)opts[i].size()) continue;
                const uint64_t delta = opts[i][chosen[i] + 1].bytes - opts[i][chosen[i]].bytes;
                if (used + delta > budget) continue;
                if (best < 0 || opts[i][chosen[i]].span < opts[best][chosen[best]].span ||
                    (opts[i][chosen[i]].span == opts[best][chosen[best]].span &&
                     delta < opts[best][chosen[best] + 1].bytes - opts[best][chosen[best]].bytes))
                    best = (int)i;
            }
            if (best < 0) break;
            used += opts[best][chosen[best] + 1].bytes - opts[best][chosen[best]].bytes;
            chosen[best] += 1;
        }
        for (size_t i = 0; i < fs.size(); ++i) {
            rb_ibf *f = fs[i];
            if (chosen[i] < 0) { f->table_tried = true; continue; }      // no table: hashed probes / streaming, never a lazy build
            f->table_budget = opts[i][chosen[i]].bytes;
            if (!ensure_table(f, (cudaStream_t)stream, true)) f->table_tried = true;
        }
    }
    return RB_OK;
}

uint64_t *rb_ibf_device_words(const rb_ibf *f) { return f ? f->d_words : nullptr; }

const uint64_t *rb_ibf_device_kmer_table(const rb_ibf *f)
{
    if (!f) return nullptr;
    std::lock_guard<std::mutex> lock(f->table_mu);
    return f->table_kind == 3 ? reinterpret_cast<const uint64_t *>(f->d_slots)
           : f->table_kind == 2 ? reinterpret_cast<const uint64_t *>(f->d_post_ids) : f->d_table;
}

// filter.resizeBins(n), src/IBF/IBFBuild.cpp:274 (update_filter): the number of rows stays, rows get wider
// when the bin count crosses a multiple of 64; existing bits keep their (row, bin) coordinates.
int rb_ibf_resize_bins(rb_ibf *f, uint64_t new_n_bins, rb_stream stream)
{
    if (!f) return fail(RB_ERR_NULL_FILTER, "null filter");
    if (f->n_shards != 1) return fail(RB_ERR_INVALID_ARG, "cannot resize a bin shard");
    if (new_n_bins < f->n_bins) return fail(RB_ERR_INVALID_CONFIG, "resizeBins cannot shrink the filter");
    if (new_n_bins == f->n_bins) return RB_OK;
    DeviceGuard g(f->device);
    cudaStream_t st = (cudaStream_t)stream;
    drop_table(f);
    const uint64_t new_width = (new_n_bins + 63) / 64;
    if (new_width != f->bin_width) {
        uint64_t *d_new = nullptr;
        const uint64_t new_words = f->n_blocks * new_width;
        RB_CUDA(cudaMalloc(&d_new, new_words * 8));
        cudaError_t e = cudaMemsetAsync(d_new, 0, new_words * 8, st);
        if (e == cudaSuccess)
            e = cudaMemcpy2DAsync(d_new, new_width * 8, f->d_words, f->bin_width * 8, f->bin_width * 8, f->n_blocks,
                                  cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { cudaFree(d_new); return fail(RB_ERR_CUDA, std::string("resizeBins: ") + cudaGetErrorString(e)); }
        cudaFree(f->d_words);
        f->d_words = d_new;
        f->bin_width = new_width;
        f->col_words = new_width;
        f->n_bits = f->n_blocks * new_width * 64;
        f->n_local_words = new_words;
    }
    f->n_bins = new_n_bins;
    f->n_bins_local = new_n_bins;
    return RB_OK;
}

// ---- build ---------------------------------------------------------------------------------------
int rb_ibf_insert_batch_dev(rb_ibf *f, const uint8_t *d_bases, const uint64_t *d_frag_begin, const uint64_t *d_frag_end,
                            const uint64_t *d_frag_bin, uint64_t n_frags, uint64_t max_frag_len, rb_stream stream)
{
    if (!f) return fail(RB_ERR_NULL_FILTER, "null filter");
    if (n_frags == 0) return RB_OK;
    if (!d_bases || !d_frag_begin || !d_frag_end || !d_frag_bin) return fail(RB_ERR_INVALID_ARG, "null device pointer");
    DeviceGuard g(f->device);
    drop_table(f);                      // the table tabulates the old matrix
    rb::InsertArgs a{};
    a.words = f->d_words; a.stride = f->col_words;
    a.bin_begin = 64 * f->col_begin; a.bin_end = a.bin_begin + f->n_bins_local; a.n_bins = f->n_bins;
    a.hp = f->hp; a.bases = d_bases; a.frag_begin = d_frag_begin; a.frag_end = d_frag_end; a.frag_bin = d_frag_bin;
    a.n_frags = n_frags; a.error_flag = f->d_err;
    int n = rb::launch_insert(a, max_frag_len, f->sm_count, &f->build_scratch, (cudaStream_t)stream);
    if (n < 0) return fail(RB_ERR_CUDA, std::string("insert launch failed: ") + cudaGetErrorString(cudaGetLastError()));
    g_launches += (uint64_t)n;
    return RB_OK;
}

int rb_ibf_insert_batch(rb_ibf *f, const char *bases, uint64_t n_bases, const uint64_t *frag_begin,
                        const uint64_t *frag_end, const uint64_t *frag_bin, uint64_t n_frags, rb_stream stream)
{
    if (!f) return fail(RB_ERR_NULL_FILTER, "null filter");
    if (n_frags == 0) return RB_OK;
    if (!bases || !frag_begin || !frag_end || !frag_bin) return fail(RB_ERR_INVALID_ARG, "null pointer");
    uint64_t max_len = 0;
    for (uint64_t i = 0; i < n_frags; ++i) {
        if (frag_end[i] < frag_begin[i] || frag_end[i] > n_bases) return fail(RB_ERR_INSERT_SEQUENCE, "fragment outside the sequence buffer");
        max_len = std::max(max_len, frag_end[i] - frag_begin[i]);
    }
    DeviceGuard g(f->device);
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t *d_bases = nullptr;
    uint64_t *d_frag = nullptr;
    auto body = [&]() -> int {
        // the kernels fetch bases as aligned 16-byte blocks: the copy may be over-read up to the next 16-byte boundary
        RB_CUDA(cudaMallocAsync(&d_bases, n_bases + 32, st));
        RB_CUDA(cudaMallocAsync(&d_frag, 3 * n_frags * 8, st));
        RB_CUDA(cudaMemcpyAsync(d_bases, bases, n_bases, cudaMemcpyHostToDevice, st));
        RB_CUDA(cudaMemcpyAsync(d_frag, frag_begin, n_frags * 8, cudaMemcpyHostToDevice, st));
        RB_CUDA(cudaMemcpyAsync(d_frag + n_frags, frag_end, n_frags * 8, cudaMemcpyHostToDevice, st));
        RB_CUDA(cudaMemcpyAsync(d_frag + 2 * n_frags, frag_bin, n_frags * 8, cudaMemcpyHostToDevice, st));
        int s2 = rb_ibf_insert_batch_dev(f, d_bases, d_frag, d_frag + n_frags, d_frag + 2 * n_frags, n_frags, max_len, stream);
        if (s2 != RB_OK) return s2;
        RB_CUDA(cudaStreamSynchronize(st));
        return sticky_insert_error(f);
    };
    int s = body();
    if (d_bases) cudaFreeAsync(d_bases, st);
    if (d_frag) cudaFreeAsync(d_frag, st);
    cudaStreamSynchronize(st);
    return s;
}

// ---- classify --------------------------------------------------------------------------------------
// keys_shared: d_keys is one array shared by the count kernels of all bin shards (see CountArgs::keys_shared)
static int count_dev_impl(const rb_ibf *f, const uint8_t *d_bases, const uint64_t *d_read_off, uint64_t n_reads,
                          uint32_t max_read_len, const uint16_t *d_thr_lut, uint32_t n_lut, uint64_t *d_keys,
                          uint16_t *d_counts_fwd, uint16_t *d_counts_rev, uint8_t *d_read_flag, rb_stream stream, int keys_shared)
{
    if (!f) return fail(RB_ERR_NULL_FILTER, "No IBF provided to classify the read!");
    if (n_reads == 0) return RB_OK;
    if (!d_bases || !d_read_off || !d_thr_lut || !d_keys) return fail(RB_ERR_INVALID_ARG, "null device pointer");
    if (n_lut == 0 || n_lut > (uint32_t)rb::kMaxLut) return fail(RB_ERR_INVALID_ARG, "n_lut must be 1..4");
    if (n_reads > 0x7FFFFFFFull) return fail(RB_ERR_INVALID_ARG, "more than 2^31-1 reads in one batch");
    DeviceGuard g(f->device);
    rb::CountArgs a{};
    a.fv = view_of(f);
    a.bases = d_bases; a.read_off = d_read_off; a.n_reads = n_reads; a.lut = d_thr_lut; a.n_lut = n_lut;
    a.keys = d_keys; a.counts_fwd = d_counts_fwd; a.counts_rev = d_counts_rev; a.read_flag = d_read_flag;
    a.keys_shared = keys_shared;
    const int which = g_count_kernel.load();
    const uint64_t *table = nullptr;
    if (which == 0 || which >= 3) table = ensure_table(f, (cudaStream_t)stream, which >= 3, n_reads);   // 3..5: table kernels
    if (which >= 3 && !table) return fail(RB_ERR_INVALID_ARG, "k-mer table not applicable to this filter (row > 4 words, k > 16 or no memory)");
    if (keys_shared && table && f->table_kind == 1) table = nullptr;      // the dense-table kernels store their keys: hashed probes fold them
    if (table && f->table_kind == 4) {
        int n = rb::launch_count_ctable(a, table, max_read_len, which == 4 ? 1 : 0, f->sm_count, (cudaStream_t)stream);
        if (n < 0) return fail(RB_ERR_COUNT_KMER, std::string("count launch failed: ") + cudaGetErrorString(cudaGetLastError()));
        g_launches += (uint64_t)n;
        return RB_OK;
    }
    if (table && f->table_kind == 3) {
        int n = rb::launch_count_slots(a, f->d_slots, f->slot_bytes, f->d_slot_ovf, max_read_len, f->sm_count, f->d_err + 1, (cudaStream_t)stream);
        if (n == -1) return fail(RB_ERR_COUNT_KMER, std::string("count launch failed: ") + cudaGetErrorString(cudaGetLastError()));
        if (n >= 0) { g_launches += (uint64_t)n; return RB_OK; }
        table = nullptr;                          // -2: counters + rings of this launch's reads do not fit shared memory
    }
    if (table && f->table_kind == 2) {
        int n = rb::launch_count_postings(a, f->d_post_ptr, f->d_post_ids, max_read_len, f->post_mean_units, f->sm_count, (cudaStream_t)stream);
        if (n == -1) return fail(RB_ERR_COUNT_KMER, std::string("count launch failed: ") + cudaGetErrorString(cudaGetLastError()));
        if (n >= 0) { g_launches += (uint64_t)n; return RB_OK; }
        table = nullptr;                          // -2: this launch's reads need wider counters than shared memory holds
    }
    if (table) {
        // selector 4 (shared-memory atomic counters) exists for one k-mer per entry only
        int n = f->table_span >= 2
                    ? rb::launch_count_wtable(a, table, f->table_span, max_read_len, f->sm_count, (cudaStream_t)stream)
                    : rb::launch_count_table(a, table, 1, max_read_len, which == 4 ? 1 : 0, f->sm_count, (cudaStream_t)stream);
        if (n < 0) return fail(RB_ERR_COUNT_KMER, std::string("count launch failed: ") + cudaGetErrorString(cudaGetLastError()));
        g_launches += (uint64_t)n;
        return RB_OK;
    }
    L2Window win(f, (cudaStream_t)stream);
    int n = rb::launch_count(a, max_read_len, which, f->sm_count, (cudaStream_t)stream);
    if (n < 0) return fail(RB_ERR_COUNT_KMER, std::string("count launch failed: ") + cudaGetErrorString(cudaGetLastError()));
    g_launches += (uint64_t)n;
    return RB_OK;
}

int rb_ibf_count_batch_dev(const rb_ibf *f, const uint8_t *d_bases, const uint64_t *d_read_off, uint64_t n_reads,
                           uint32_t max_read_len, const uint16_t *d_thr_lut, uint32_t n_lut, uint64_t *d_keys,
                           uint16_t *d_counts_fwd, uint16_t *d_counts_rev, uint8_t *d_read_flag, rb_stream stream)
{
    return count_dev_impl(f, d_bases, d_read_off, n_reads, max_read_len, d_thr_lut, n_lut, d_keys, d_counts_fwd, d_counts_rev,
                          d_read_flag, stream, 0);
}

// ---- bin shards driven by ONE host process: count on every device, keys folded over NVLink into shard 0's array ----------
int rb_ibf_count_batch_sharded(const rb_ibf *const *shards, uint32_t n_shards, const char *bases, const uint64_t *read_off,
                               uint64_t n_reads, const uint16_t *thr_lut, uint32_t n_lut, uint16_t *max_count, uint8_t *hit,
                               uint32_t *argmax_bin, uint8_t *read_flag)
{
    if (!shards || n_shards == 0) return fail(RB_ERR_NULL_FILTER, "No IBF provided to classify the read!");
    if (n_shards > 64) return fail(RB_ERR_INVALID_ARG, "more than 64 shards");
    for (uint32_t i = 0; i < n_shards; ++i) {
        if (!shards[i]) return fail(RB_ERR_NULL_FILTER, "No IBF provided to classify the read!");
        const rb_ibf *a = shards[0], *b = shards[i];
        if (b->n_bins != a->n_bins || b->n_bits != a->n_bits || b->k != a->k || b->n_hash != a->n_hash)
            return fail(RB_ERR_INVALID_ARG, "the handles are not shards of one filter");
    }
    if (n_reads == 0) return RB_OK;
    if (!bases || !read_off || !thr_lut) return fail(RB_ERR_INVALID_ARG, "null pointer");
    if (n_lut == 0 || n_lut > (uint32_t)rb::kMaxLut) return fail(RB_ERR_INVALID_ARG, "n_lut must be 1..4");
    if (n_reads > 0x7FFFFFFFull) return fail(RB_ERR_INVALID_ARG, "more than 2^31-1 reads in one batch");
    const uint64_t max_len = rb::max_read_length(read_off, n_reads);
    if (max_len >> 63) return fail(RB_ERR_INVALID_ARG, "read offsets must be non-decreasing");
    const uint32_t max_len32 = (uint32_t)std::min<uint64_t>(max_len, 0xFFFFFFFFu);
    const uint64_t b0 = read_off[0], nb = read_off[n_reads] - b0, nk = (uint64_t)n_lut * n_reads;
    const rb_ibf *root = shards[0];
    int prev = -1;
    cudaGetDevice(&prev);
    struct Per { cudaStream_t st = nullptr; cudaEvent_t done = nullptr; uint8_t *d_bases = nullptr; uint64_t *d_off = nullptr;
                 uint16_t *d_lut = nullptr; };
    std::vector<Per> per(n_shards);
    uint64_t *d_keys = nullptr;
    uint8_t *d_flag = nullptr, *d_res = nullptr;
    cudaEvent_t zeroed = nullptr;
    auto body = [&]() -> int {
        // shard 0's device owns the key array; every other device gets peer access to it
        RB_CUDA(cudaSetDevice(root->device));
        RB_CUDA(cudaStreamCreateWithFlags(&per[0].st, cudaStreamNonBlocking));
        RB_CUDA(cudaEventCreateWithFlags(&zeroed, cudaEventDisableTiming));
        // plain cudaMalloc: peers reach it through cudaDeviceEnablePeerAccess (stream-ordered pool memory would need its own
        // cudaMemPoolSetAccess grant per peer)
        RB_CUDA(cudaMalloc(&d_keys, nk * 8));
        RB_CUDA(cudaMallocAsync(&d_flag, n_reads, per[0].st));
        RB_CUDA(cudaMallocAsync(&d_res, nk * 7 + 64, per[0].st));
        RB_CUDA(cudaMemsetAsync(d_keys, 0, nk * 8, per[0].st));
        RB_CUDA(cudaEventRecord(zeroed, per[0].st));
        for (uint32_t i = 0; i < n_shards; ++i) {
            const rb_ibf *f = shards[i];
            RB_CUDA(cudaSetDevice(f->device));
            if (f->device != root->device) {
                int can = 0;
                RB_CUDA(cudaDeviceCanAccessPeer(&can, f->device, root->device));
                if (!can) return fail(RB_ERR_INVALID_ARG, "a shard's device has no peer access to shard 0's device");
                cudaError_t e = cudaDeviceEnablePeerAccess(root->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(RB_ERR_CUDA, std::string("peer access: ") + cudaGetErrorString(e));
                cudaGetLastError();
            }
            Per &p = per[i];
            if (i) RB_CUDA(cudaStreamCreateWithFlags(&p.st, cudaStreamNonBlocking));
            RB_CUDA(cudaEventCreateWithFlags(&p.done, cudaEventDisableTiming));
            RB_CUDA(cudaMallocAsync(&p.d_bases, nb + 32, p.st));
            RB_CUDA(cudaMallocAsync(&p.d_off, (n_reads + 1) * 8, p.st));
            RB_CUDA(cudaMallocAsync(&p.d_lut, (size_t)n_lut * rb::kLutSize * 2, p.st));
            RB_CUDA(cudaMemcpyAsync(p.d_bases, bases + b0, nb, cudaMemcpyHostToDevice, p.st));
            RB_CUDA(cudaMemcpyAsync(p.d_off, read_off, (n_reads + 1) * 8, cudaMemcpyHostToDevice, p.st));
            RB_CUDA(cudaMemcpyAsync(p.d_lut, thr_lut, (size_t)n_lut * rb::kLutSize * 2, cudaMemcpyHostToDevice, p.st));
            g_h2d_bytes += nb + (n_reads + 1) * 8 + (size_t)n_lut * rb::kLutSize * 2;
            RB_CUDA(cudaStreamWaitEvent(p.st, zeroed, 0));
            const uint8_t *biased = reinterpret_cast<const uint8_t *>(reinterpret_cast<uintptr_t>(p.d_bases) - (uintptr_t)b0);
            int s2 = count_dev_impl(f, biased, p.d_off, n_reads, max_len32, p.d_lut, n_lut, d_keys, nullptr, nullptr,
                                    i == 0 ? d_flag : nullptr, p.st, 1);
            if (s2 != RB_OK) return s2;
            RB_CUDA(cudaEventRecord(p.done, p.st));
        }
        RB_CUDA(cudaSetDevice(root->device));
        for (uint32_t i = 1; i < n_shards; ++i) RB_CUDA(cudaStreamWaitEvent(per[0].st, per[i].done, 0));
        uint32_t *r_amax = reinterpret_cast<uint32_t *>(d_res);
        uint16_t *r_max = reinterpret_cast<uint16_t *>(d_res + nk * 4);
        uint8_t *r_hit = d_res + nk * 6;
        int nl = rb::launch_keys_decode(d_keys, nk, r_max, r_hit, r_amax, per[0].st);
        if (nl < 0) return fail(RB_ERR_CUDA, "decode launch failed");
        g_launches += (uint64_t)nl;
        if (argmax_bin) RB_CUDA(cudaMemcpyAsync(argmax_bin, r_amax, nk * 4, cudaMemcpyDeviceToHost, per[0].st));
        if (max_count) RB_CUDA(cudaMemcpyAsync(max_count, r_max, nk * 2, cudaMemcpyDeviceToHost, per[0].st));
        if (hit) RB_CUDA(cudaMemcpyAsync(hit, r_hit, nk, cudaMemcpyDeviceToHost, per[0].st));
        if (read_flag) RB_CUDA(cudaMemcpyAsync(read_flag, d_flag, n_reads, cudaMemcpyDeviceToHost, per[0].st));
        g_d2h_bytes += (argmax_bin ? nk * 4 : 0) + (max_count ? nk * 2 : 0) + (hit ? nk : 0) + (read_flag ? n_reads : 0);
        cudaError_t e = cudaStreamSynchronize(per[0].st);
        if (e != cudaSuccess) return fail(RB_ERR_COUNT_KMER, std::string("Error counting kmers in IBF bins: ") + cudaGetErrorString(e));
        return RB_OK;
    };
    int st = body();
    const std::string keep = g_last_error;
    for (uint32_t i = 0; i < n_shards; ++i) {                      // drain and free, also after an error
        Per &p = per[i];
        if (!p.st) continue;
        cudaSetDevice(shards[i]->device);
        cudaStreamSynchronize(p.st);
        if (p.d_bases) cudaFreeAsync(p.d_bases, p.st);
        if (p.d_off) cudaFreeAsync(p.d_off, p.st);
        if (p.d_lut) cudaFreeAsync(p.d_lut, p.st);
        if (i == 0) {
            if (d_flag) cudaFreeAsync(d_flag, p.st);
            if (d_res) cudaFreeAsync(d_res, p.st);
        }
        cudaStreamSynchronize(p.st);
        if (p.done) cudaEventDestroy(p.done);
        cudaStreamDestroy(p.st);
    }
    if (zeroed) cudaEventDestroy(zeroed);
    if (d_keys) { cudaSetDevice(root->device); cudaFree(d_keys); }
    if (prev >= 0) cudaSetDevice(prev);
    cudaGetLastError();
    g_last_error = keep;
    return st;
}

// ---- bin shards, one process per GPU: the combine as one NCCL call (libnccl is looked up at run time) ----------------------
int rb_keys_combine_nccl(void *nccl_comm, uint64_t *d_keys, uint64_t n, rb_stream stream)
{
    if (!nccl_comm || !d_keys) return fail(RB_ERR_INVALID_ARG, "null communicator or keys");
    if (n == 0) return RB_OK;
    // ncclResult_t ncclAllReduce(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t)
    typedef int (*allreduce_fn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
    static allreduce_fn fn = [] {
        // the process usually has NCCL loaded already (its own link, or torch's bundled copy): take that one
        void *sym = dlsym(RTLD_DEFAULT, "ncclAllReduce");
        if (!sym) {
            void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
            if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
            if (h) sym = dlsym(h, "ncclAllReduce");
        }
        return reinterpret_cast<allreduce_fn>(sym);
    }();
    if (!fn) return fail(RB_ERR_INVALID_ARG, "libnccl.so.2 not found (ncclAllReduce)");
    // keys < 2^49: MAX over uint64; nccl.h: ncclUint64 = 5, ncclMax = 2
    const int r = fn(d_keys, d_keys, (size_t)n, 5, 2, nccl_comm, (cudaStream_t)stream);
    if (r != 0) return fail(RB_ERR_CUDA, "ncclAllReduce failed with ncclResult " + std::to_string(r));
    return RB_OK;
}

int rb_keys_decode_dev(const uint64_t *d_keys, uint64_t n, uint16_t *d_max_count, uint8_t *d_hit, uint32_t *d_argmax_bin,
                       int device, rb_stream stream)
{
    if (n == 0) return RB_OK;
    if (!d_keys) return fail(RB_ERR_INVALID_ARG, "null keys");
    DeviceGuard g(device);
    int k = rb::launch_keys_decode(d_keys, n, d_max_count, d_hit, d_argmax_bin, (cudaStream_t)stream);
    if (k < 0) return fail(RB_ERR_CUDA, "decode launch failed");
    g_launches += (uint64_t)k;
    return RB_OK;
}

int rb_ibf_count_batch(const rb_ibf *f, const char *bases, const uint64_t *read_off, uint64_t n_reads,
                       const uint16_t *thr_lut, uint32_t n_lut, uint16_t *counts_fwd, uint16_t *counts_rev,
                       uint16_t *max_count, uint8_t *hit, uint32_t *argmax_bin, uint8_t *read_flag, rb_stream stream)
{
    if (!f) return fail(RB_ERR_NULL_FILTER, "No IBF provided to classify the read!");
    if (n_reads == 0) return RB_OK;
    if (!bases || !read_off || !thr_lut) return fail(RB_ERR_INVALID_ARG, "null pointer");
    if (n_lut == 0 || n_lut > (uint32_t)rb::kMaxLut) return fail(RB_ERR_INVALID_ARG, "n_lut must be 1..4");
    if (n_reads > 0x7FFFFFFFull) return fail(RB_ERR_INVALID_ARG, "more than 2^31-1 reads in one batch");
    if (read_off[n_reads] < read_off[0]) return fail(RB_ERR_INVALID_ARG, "read offsets must be non-decreasing");
    DeviceGuard g(f->device);
    CallCtx *ctx = acquire_ctx(f);
    if (!ctx) return fail(RB_ERR_CUDA, "cannot create streams for the classify call");
    int st = run_count_batch(f, ctx, reinterpret_cast<const uint8_t *>(bases), read_off, n_reads, thr_lut, n_lut, counts_fwd,
                             counts_rev, max_count, hit, argmax_bin, read_flag, (cudaStream_t)stream);
    release_ctx(f, ctx);
    if (st == RB_OK && f->table_kind == 3) {
        // the slot kernel never hangs on a bulk copy that does not land: it gives up and says so here
        unsigned int gave_up = 0;
        RB_CUDA(cudaMemcpy(&gave_up, f->d_err + 1, sizeof(gave_up), cudaMemcpyDeviceToHost));
        if (gave_up) {
            cudaMemset(f->d_err + 1, 0, sizeof(gave_up));
            return fail(RB_ERR_COUNT_KMER, "Error counting kmers in IBF bins: a postings slot copy did not complete");
        }
    }
    return st;
}

int rb_ibf_count_traffic_dev(const rb_ibf *f, const uint8_t *d_bases, const uint64_t *d_read_off, uint64_t n_reads, uint32_t n_lut,
                             uint64_t *table_bytes, uint64_t *table_requests, uint64_t *io_bytes, rb_stream stream)
{
    if (!f) return fail(RB_ERR_NULL_FILTER, "no filter");
    if (!d_bases || !d_read_off) return fail(RB_ERR_INVALID_ARG, "null device pointer");
    DeviceGuard g(f->device);
    cudaStream_t st = (cudaStream_t)stream;
    int kind, span;
    uint32_t entry_bytes;
    const uint32_t *ptr = nullptr;
    {
        std::lock_guard<std::mutex> lock(f->table_mu);
        kind = f->table_kind; span = f->table_span;
        if (kind == 4) { entry_bytes = 16u * (uint32_t)rb::ctable_lanes(f->col_words); kind = 1; }
        else if (kind == 1) entry_bytes = (uint32_t)(span == 1 ? 16 * f->col_words : (span == 2 ? 2 : 4) * 16 * f->col_words);
        else if (kind == 2) { entry_bytes = 16; ptr = f->d_post_ptr; }
        else if (kind == 3) { entry_bytes = f->slot_bytes; ptr = reinterpret_cast<const uint32_t *>(f->d_slots); }
        else entry_bytes = (uint32_t)(8 * f->col_words);
    }
    unsigned long long *d_out = nullptr, h_out[3] = {0, 0, 0};
    RB_CUDA(cudaMalloc(&d_out, sizeof(h_out)));
    int n = rb::launch_traffic(d_bases, d_read_off, n_reads, (uint32_t)f->k, (uint32_t)f->n_hash, kind, span, entry_bytes, ptr, d_out,
                               f->sm_count, st);
    cudaError_t e = n < 0 ? cudaErrorUnknown : cudaMemcpyAsync(h_out, d_out, sizeof(h_out), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_out);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(RB_ERR_CUDA, "traffic kernel failed"); }
    g_launches += 1;
    if (table_bytes) *table_bytes = h_out[0];
    if (table_requests) *table_requests = h_out[1];
    if (io_bytes) *io_bytes = h_out[2] + n_reads * 8 + 8 + (uint64_t)n_lut * n_reads * 8;    // bases + offsets in, keys out
    return RB_OK;
}

int rb_ibf_transfer_policy(const rb_ibf *f, int *choice, double *ns_per_base_packed, double *ns_per_base_ascii)
{
    if (!f) return fail(RB_ERR_NULL_FILTER, "no filter");
    std::lock_guard<std::mutex> lock(f->pol_mu);
    if (choice) *choice = f->pol_choice;
    if (ns_per_base_packed) *ns_per_base_packed = f->pol_ns_per_base[0];
    if (ns_per_base_ascii) *ns_per_base_ascii = f->pol_ns_per_base[1];
    return RB_OK;
}

int rb_transfer_bytes(uint64_t *h2d, uint64_t *d2h)
{
    if (h2d) *h2d = g_h2d_bytes.load();
    if (d2h) *d2h = g_d2h_bytes.load();
    return RB_OK;
}

int rb_microbench_host_read(const void *buf, uint64_t n_bytes, uint32_t reps, double *gb_per_s)
{
    if (!buf || !gb_per_s || n_bytes == 0 || reps == 0) return fail(RB_ERR_INVALID_ARG, "null buffer");
    volatile uint64_t sink = rb::stream_read(static_cast<const uint8_t *>(buf), (size_t)n_bytes);      // warm the pool
    const auto t0 = std::chrono::steady_clock::now();
    for (uint32_t r = 0; r < reps; ++r) sink = sink ^ rb::stream_read(static_cast<const uint8_t *>(buf), (size_t)n_bytes);
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    *gb_per_s = (double)n_bytes * reps / s / 1e9;
    return RB_OK;
}

int rb_host_pack_info(int *threads, int *isa)
{
    if (threads) *threads = rb::host_threads();
    if (isa) *isa = rb::pack_isa();
    return RB_OK;
}

}  // extern "C"
