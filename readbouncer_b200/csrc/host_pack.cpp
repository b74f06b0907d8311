// host_pack.cpp -- see host_pack.hpp.  Compiled by g++ (function-level AVX2, chosen at run time).
#include "host_pack.hpp"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace rb {

namespace {

inline void pack_word_scalar(const uint8_t *p, size_t n, uint32_t &lo, uint32_t &hi, uint32_t &bad)
{
    lo = hi = bad = 0;
    for (size_t i = 0; i < n; ++i) {
        const uint32_t c = p[i], u = c & 0xDFu;
        if (u == 'A' || u == 'C' || u == 'G' || u == 'T' || u == 'U') {
            lo |= ((c >> 1) & 1u) << i;
            hi |= ((c >> 2) & 1u) << i;
        } else {
            bad |= 1u << i;
        }
    }
}

void pack_scalar(const uint8_t *bases, size_t n, uint32_t *lo, uint32_t *hi, uint32_t *bad)
{
    const size_t nw = (n + 31) / 32;
    for (size_t w = 0; w < nw; ++w)
        pack_word_scalar(bases + 32 * w, std::min<size_t>(32, n - 32 * w), lo[w], hi[w], bad[w]);
}

#if defined(__x86_64__)
// Character class by two nibble look-ups (pshufb): low nibble 1/3/7 with high nibble 4/6 = A C G a c g, low nibble
// 4/5 with high nibble 5/7 = T U t u.  5 vector ops instead of 9 compares and ors.
__attribute__((target("avx2"))) inline void pack32_avx2(const uint8_t *p, const __m256i lut_lo, const __m256i lut_hi,
                                                        const __m256i nib, uint32_t &lo, uint32_t &hi, uint32_t &bad)
{
    const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(p));
    const __m256i cls = _mm256_and_si256(_mm256_shuffle_epi8(lut_lo, _mm256_and_si256(x, nib)),
                                         _mm256_shuffle_epi8(lut_hi, _mm256_and_si256(_mm256_srli_epi16(x, 4), nib)));
    const uint32_t notgood = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(cls, _mm256_setzero_si256()));
    // movemask takes bit 7 of every byte: shift code bit 1 (resp. 2) of each byte up to it
    lo = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(x, 6)) & ~notgood;
    hi = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(x, 5)) & ~notgood;
    bad = notgood;
}

__attribute__((target("avx2"))) void pack_avx2(const uint8_t *bases, size_t n, uint32_t *lo, uint32_t *hi, uint32_t *bad)
{
    const size_t full = n / 32;
    // bytes >= 0x80 index the high-nibble table at 8..15 (class 0); pshufb itself zeroes lanes whose index has bit 7
    // set, which only ever happens for the low-nibble look-up of such bytes -- also class 0
    const __m256i lut_lo = _mm256_setr_epi8(0, 1, 0, 1, 2, 2, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0,
                                            0, 1, 0, 1, 2, 2, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i lut_hi = _mm256_setr_epi8(0, 0, 0, 0, 1, 2, 1, 2, 0, 0, 0, 0, 0, 0, 0, 0,
                                            0, 0, 0, 0, 1, 2, 1, 2, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i nib = _mm256_set1_epi8(0x0F);
    size_t w = 0;
    for (; w + 2 <= full; w += 2) {
        pack32_avx2(bases + 32 * w, lut_lo, lut_hi, nib, lo[w], hi[w], bad[w]);
        pack32_avx2(bases + 32 * w + 32, lut_lo, lut_hi, nib, lo[w + 1], hi[w + 1], bad[w + 1]);
    }
    for (; w < full; ++w) pack32_avx2(bases + 32 * w, lut_lo, lut_hi, nib, lo[w], hi[w], bad[w]);
    if (n % 32) pack_word_scalar(bases + 32 * full, n % 32, lo[full], hi[full], bad[full]);
}

// AVX-512 (BW + VBMI): one 64-entry byte look-up (vpermb on the low 6 bits) classifies 64 bases, the mask-register tests
// turn code bits into 64-bit masks directly -- no movemask, no shifts.
//   good(c) = table[c & 63] and bit 6 of c set and bit 7 clear: A C G T U are 0x41 0x43 0x47 0x54 0x55 (+ 0x20 lower case),
//   so the table marks 0x01 0x03 0x07 0x14 0x15 0x21 0x23 0x27 0x34 0x35 and the two high bits rule out the bytes that alias.
// vpermb is one port-5 operation (the two-table vpermt2b took three and a register copy); `good` serves as the write mask of
// the two plane tests, so a 64-base block costs a load, vpermb, vpaddb, vpternlogd, vpmovb2m, 2 vptestmb, knot, 3 mask stores.
// lazy_dirty != null: the "not ACGT" plane is written only from the first block that has such a base on (the words before it
// are zero-filled then); *lazy_dirty tells whether that happened.  A clean call leaves `bad` untouched.
__attribute__((target("avx512f,avx512bw,avx512vbmi"))) void pack_avx512(const uint8_t *bases, size_t n, uint32_t *lo,
                                                                        uint32_t *hi, uint32_t *bad, bool nt, bool *lazy_dirty)
{
    bool dirty = lazy_dirty == nullptr;                  // not lazy: always write
    nt = nt && ((reinterpret_cast<uintptr_t>(lo) | reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(bad)) & 7) == 0;
    alignas(64) uint8_t tab[64];
    for (int c = 0; c < 64; ++c) {
        const int u = 0x40 | (c & 0x1F);                    // the upper-case letter both aliases (0x40 | c, 0x60 | c - 0x20) stand for
        tab[c] = (u == 'A' || u == 'C' || u == 'G' || u == 'T' || u == 'U') ? 0xFF : 0x00;
    }
    const __m512i tb = _mm512_load_si512(tab);
    const __m512i c02 = _mm512_set1_epi8(0x02), c04 = _mm512_set1_epi8(0x04);
    // sign bit of (cls & (x << 1) & ~x) = table hit, bit 6 set, bit 7 clear; the table index is the low 6 bits of the byte
#define RB_CLASSIFY(x) _mm512_movepi8_mask(_mm512_ternarylogic_epi32(_mm512_permutexvar_epi8((x), tb), _mm512_add_epi8((x), (x)), (x), 0x40))
    const size_t full = n / 64;
    size_t w = 0;
    if (nt && ((reinterpret_cast<uintptr_t>(lo) | reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(bad)) & 63) == 0) {
        // 512 bases per step: eight mask words per plane leave as ONE full-line write-combining store, and the input is
        // prefetched 2 KB ahead (a core's demand stream alone does not keep enough lines in flight)
        for (; w + 8 <= full; w += 8) {
            alignas(64) uint64_t L[8], H[8], B[8];
#pragma GCC unroll 8
            for (int j = 0; j < 8; ++j) {
                const uint8_t *p = bases + 64 * (w + j);
                _mm_prefetch(reinterpret_cast<const char *>(p) + 2048, _MM_HINT_T0);
                const __m512i x = _mm512_loadu_si512(p);
                const __mmask64 good = RB_CLASSIFY(x);
                L[j] = _mm512_mask_test_epi8_mask(good, x, c02);
                H[j] = _mm512_mask_test_epi8_mask(good, x, c04);
                B[j] = ~good;
            }
            _mm512_stream_si512(reinterpret_cast<__m512i *>(lo + 2 * w), _mm512_load_si512(L));
            _mm512_stream_si512(reinterpret_cast<__m512i *>(hi + 2 * w), _mm512_load_si512(H));
            if (!dirty && (B[0] | B[1] | B[2] | B[3] | B[4] | B[5] | B[6] | B[7])) {
                dirty = true;
                std::memset(bad, 0, 8 * w);              // the clean blocks before this one
            }
            if (dirty) _mm512_stream_si512(reinterpret_cast<__m512i *>(bad + 2 * w), _mm512_load_si512(B));
        }
    }
    for (; w < full; ++w) {
        const __m512i x = _mm512_loadu_si512(bases + 64 * w);
        const __mmask64 good = RB_CLASSIFY(x);
        const uint64_t l = _mm512_mask_test_epi8_mask(good, x, c02), h = _mm512_mask_test_epi8_mask(good, x, c04), b = ~good;
        if (!dirty && b) { dirty = true; std::memset(bad, 0, 8 * w); }
        if (nt) {        // write-combining stores: the planes go to DRAM for the DMA engine, not into this core's cache
            _mm_stream_si64(reinterpret_cast<long long *>(lo + 2 * w), (long long)l);
            _mm_stream_si64(reinterpret_cast<long long *>(hi + 2 * w), (long long)h);
            if (dirty) _mm_stream_si64(reinterpret_cast<long long *>(bad + 2 * w), (long long)b);
        } else {
            std::memcpy(lo + 2 * w, &l, 8);
            std::memcpy(hi + 2 * w, &h, 8);
            if (dirty) std::memcpy(bad + 2 * w, &b, 8);
        }
    }
    if (nt) _mm_sfence();
    const size_t done = 64 * full, rest = n - done;
    for (size_t o = 0; o < rest; o += 32) {
        uint32_t b = 0;
        pack_word_scalar(bases + done + o, std::min<size_t>(32, rest - o), lo[2 * full + o / 32], hi[2 * full + o / 32], b);
        if (!dirty && b) { dirty = true; std::memset(bad, 0, 4 * (2 * full + o / 32)); }
        if (dirty) bad[2 * full + o / 32] = b;
    }
    if (lazy_dirty) *lazy_dirty = dirty;
#undef RB_CLASSIFY
}
#endif

int detect_isa()     // 0 scalar, 2 AVX2, 5 AVX-512 (BW + VBMI)
{
#if defined(__x86_64__)
    int cap = 5;
    if (const char *e = std::getenv("RB_HOST_PACK_ISA")) cap = std::atoi(e);
    if (const char *e = std::getenv("RB_HOST_PACK_SCALAR")) if (e[0] == '1') cap = 0;
    if (cap >= 5 && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vbmi")) return 5;
    if (cap >= 2 && __builtin_cpu_supports("avx2")) return 2;
#endif
    return 0;
}

const int g_isa = detect_isa();

// ---- a small pool of spinning-while-busy, sleeping-while-idle threads ---------------------------------
class Pool {
public:
    Pool()
    {
        int n = 0;
        if (const char *e = std::getenv("RB_HOST_THREADS")) n = std::atoi(e);
        if (n <= 0) {
            // share the cores between the ranks of one node (torchrun exports LOCAL_WORLD_SIZE), at most 16 each
            unsigned cores = std::max(1u, std::thread::hardware_concurrency()), ranks = 1;
            if (const char *e = std::getenv("LOCAL_WORLD_SIZE")) ranks = (unsigned)std::max(1, std::atoi(e));
            n = (int)std::min<unsigned>(std::max(1u, cores / ranks), 16u);
        }
        n_workers_ = std::max(0, n - 1);
        for (int i = 0; i < n_workers_; ++i) threads_.emplace_back([this] { worker(); });
    }
    ~Pool()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
    }
    int size() const { return n_workers_ + 1; }

    void run(size_t n_tasks, const std::function<void(size_t)> &fn, const std::function<void(size_t)> &poll)
    {
        std::unique_lock<std::mutex> busy(busy_, std::try_to_lock);
        if (!busy.owns_lock() || n_workers_ == 0 || n_tasks < 2) {        // pool taken by another call: go alone
            for (size_t i = 0; i < n_tasks; ++i) { fn(i); if (poll) poll(i + 1); }
            return;
        }
        done_flags_.assign(n_tasks, 0);
        fn_ = &fn;
        n_tasks_ = n_tasks;
        next_.store(0, std::memory_order_relaxed);
        finished_.store(0, std::memory_order_relaxed);
        {
            std::lock_guard<std::mutex> lk(mu_);
            ++generation_;
            job_open_ = true;
        }
        cv_.notify_all();
        size_t prefix = 0;
        auto advance = [&] {
            while (prefix < n_tasks && __atomic_load_n(&done_flags_[prefix], __ATOMIC_ACQUIRE)) ++prefix;
            if (poll) poll(prefix);
        };
        for (;;) {
            const size_t i = next_.fetch_add(1, std::memory_order_relaxed);
            if (i >= n_tasks) break;
            fn(i);
            __atomic_store_n(&done_flags_[i], 1, __ATOMIC_RELEASE);
            finished_.fetch_add(1, std::memory_order_release);
            advance();
        }
        while (prefix < n_tasks) { advance(); if (prefix < n_tasks) std::this_thread::yield(); }
        {
            std::lock_guard<std::mutex> lk(mu_);      // late wakers must not join a job that is over
            job_open_ = false;
        }
        // workers that joined may still be between "no task left" and going idle
        while (active_.load(std::memory_order_acquire) != 0) std::this_thread::yield();
        fn_ = nullptr;
    }

private:
    void worker()
    {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
                if (!job_open_) continue;             // woke up after the job had finished
                active_.fetch_add(1, std::memory_order_acq_rel);
            }
            for (;;) {
                const size_t i = next_.fetch_add(1, std::memory_order_relaxed);
                if (i >= n_tasks_) break;
                (*fn_)(i);
                __atomic_store_n(&done_flags_[i], 1, __ATOMIC_RELEASE);
                finished_.fetch_add(1, std::memory_order_release);
            }
            active_.fetch_sub(1, std::memory_order_acq_rel);
        }
    }

    int n_workers_ = 0;
    std::vector<std::thread> threads_;
    std::mutex mu_, busy_;
    std::condition_variable cv_;
    bool stop_ = false, job_open_ = false;
    uint64_t generation_ = 0;
    const std::function<void(size_t)> *fn_ = nullptr;
    size_t n_tasks_ = 0;
    std::vector<char> done_flags_;
    std::atomic<size_t> next_{0}, finished_{0};
    std::atomic<int> active_{0};
};

Pool &pool()
{
    static Pool *p = new Pool();      // leaked on purpose: worker threads must not be joined from a static destructor
    return *p;
}

}  // namespace

#if defined(__x86_64__)
namespace {
__attribute__((target("avx512f"))) uint64_t max_diff_avx512(const uint64_t *off, size_t n)
{
    __m512i m = _mm512_setzero_si512();
    size_t i = 0;
    for (; i + 8 <= n; i += 8)
        m = _mm512_max_epu64(m, _mm512_sub_epi64(_mm512_loadu_si512(off + i + 1), _mm512_loadu_si512(off + i)));
    uint64_t r = _mm512_reduce_max_epu64(m);
    for (; i < n; ++i) r = std::max(r, off[i + 1] - off[i]);
    return r;
}
}  // namespace
#endif

uint64_t max_read_length(const uint64_t *off, size_t n)
{
#if defined(__x86_64__)
    if (g_isa == 5) return max_diff_avx512(off, n);
#endif
    uint64_t r = 0;
    for (size_t i = 0; i < n; ++i) r = std::max(r, off[i + 1] - off[i]);
    return r;
}

bool pack_bases_lazy(const uint8_t *bases, size_t n, uint32_t *lo, uint32_t *hi, uint32_t *bad, bool streaming_stores)
{
#if defined(__x86_64__)
    if (g_isa == 5) { bool dirty = false; pack_avx512(bases, n, lo, hi, bad, streaming_stores, &dirty); return dirty; }
#endif
    // the other code paths write the whole plane (so the region is consistent either way) and report what it holds
    pack_bases(bases, n, lo, hi, bad, streaming_stores);
    const size_t nw = (n + 31) / 32;
    for (size_t w = 0; w < nw; ++w)
        if (bad[w]) return true;
    return false;
}

void pack_bases(const uint8_t *bases, size_t n, uint32_t *lo, uint32_t *hi, uint32_t *bad, bool streaming_stores)
{
#if defined(__x86_64__)
    if (g_isa == 5) { pack_avx512(bases, n, lo, hi, bad, streaming_stores, nullptr); return; }
    if (g_isa == 2) { pack_avx2(bases, n, lo, hi, bad); return; }
#endif
    pack_scalar(bases, n, lo, hi, bad);
}

int pack_isa() { return g_isa; }

void parallel_tasks(size_t n_tasks, const std::function<void(size_t)> &fn, const std::function<void(size_t)> &poll)
{
    pool().run(n_tasks, fn, poll);
}

int host_threads() { return pool().size(); }

#if defined(__x86_64__)
namespace {
__attribute__((target("avx2"))) uint64_t sum_avx2(const uint8_t *p, size_t n)
{
    __m256i a = _mm256_setzero_si256(), b = a, c = a, d = a;
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        _mm_prefetch(reinterpret_cast<const char *>(p + i) + 2048, _MM_HINT_T0);
        a = _mm256_xor_si256(a, _mm256_loadu_si256(reinterpret_cast<const __m256i *>(p + i)));
        b = _mm256_xor_si256(b, _mm256_loadu_si256(reinterpret_cast<const __m256i *>(p + i + 32)));
        c = _mm256_xor_si256(c, _mm256_loadu_si256(reinterpret_cast<const __m256i *>(p + i + 64)));
        d = _mm256_xor_si256(d, _mm256_loadu_si256(reinterpret_cast<const __m256i *>(p + i + 96)));
    }
    a = _mm256_xor_si256(_mm256_xor_si256(a, b), _mm256_xor_si256(c, d));
    alignas(32) uint64_t v[4];
    _mm256_store_si256(reinterpret_cast<__m256i *>(v), a);
    uint64_t r = v[0] ^ v[1] ^ v[2] ^ v[3];
    for (; i < n; ++i) r ^= p[i];
    return r;
}
}  // namespace
#endif

uint64_t stream_read(const uint8_t *buf, size_t n)
{
    constexpr size_t kTask = 128u << 10;
    const size_t n_tasks = (n + kTask - 1) / kTask;
    std::atomic<uint64_t> acc{0};
    auto task = [&](size_t t) {
        const size_t o = t * kTask, m = std::min(kTask, n - o);
        uint64_t r = 0;
#if defined(__x86_64__)
        if (g_isa >= 2) r = sum_avx2(buf + o, m);
        else
#endif
            for (size_t i = 0; i < m; ++i) r ^= buf[o + i];
        acc.fetch_xor(r, std::memory_order_relaxed);
    };
    pool().run(n_tasks, task, nullptr);
    return acc.load();
}

}  // namespace rb
