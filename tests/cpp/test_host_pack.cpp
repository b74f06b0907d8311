// Host packer (readbouncer_b200/csrc/host_pack.cpp): the three bit planes of a base string must match a plain
// restatement for every instruction-set path (env RB_HOST_PACK_ISA caps it), every tail length, every byte value.
#include "host_pack.hpp"

#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

static void restate(const uint8_t *b, size_t n, std::vector<uint32_t> &lo, std::vector<uint32_t> &hi, std::vector<uint32_t> &bad)
{
    const size_t nw = (n + 31) / 32;
    lo.assign(nw, 0); hi.assign(nw, 0); bad.assign(nw, 0);
    for (size_t i = 0; i < n; ++i) {
        const uint32_t c = b[i], u = c & 0xDFu;
        const bool ok = c < 0x80 && (u == 'A' || u == 'C' || u == 'G' || u == 'T' || u == 'U');
        if (ok) { lo[i >> 5] |= ((c >> 1) & 1u) << (i & 31); hi[i >> 5] |= ((c >> 2) & 1u) << (i & 31); }
        else bad[i >> 5] |= 1u << (i & 31);
    }
}

int main()
{
    std::mt19937_64 rng(7);
    int fails = 0;
    // all 256 byte values, then random mixes at awkward lengths
    std::vector<size_t> lens = {0, 1, 31, 32, 33, 63, 64, 65, 127, 128, 129, 250, 256, 1000, 4097, 131072, 131072 * 3 + 5};
    for (size_t n : lens) {
        std::vector<uint8_t> b(n + 64);
        for (size_t i = 0; i < n; ++i) {
            const uint64_t r = rng();
            b[i] = (r % 50 == 0) ? (uint8_t)(r >> 8) : (uint8_t)"ACGTacgtUuNn"[(r >> 8) % 12];
        }
        if (n >= 256) for (int c = 0; c < 256; ++c) b[c] = (uint8_t)c;
        std::vector<uint32_t> lo, hi, bad;
        restate(b.data(), n, lo, hi, bad);
        const size_t nw = (n + 31) / 32;
        std::vector<uint32_t> l2(nw + 1, 0xDEADBEEF), h2(nw + 1, 0xDEADBEEF), b2(nw + 1, 0xDEADBEEF);
        rb::pack_bases(b.data(), n, l2.data(), h2.data(), b2.data());
        if (std::memcmp(lo.data(), l2.data(), nw * 4) || std::memcmp(hi.data(), h2.data(), nw * 4) ||
            std::memcmp(bad.data(), b2.data(), nw * 4) || l2[nw] != 0xDEADBEEF || h2[nw] != 0xDEADBEEF || b2[nw] != 0xDEADBEEF) {
            std::printf("MISMATCH n=%zu\n", n);
            ++fails;
        }
        // lazy "not ACGT" plane: written completely when the range has such a base (first one early, in the middle, at the very
        // end, in the scalar tail), untouched-or-complete and reported clean when it has none; aligned planes with streaming
        // stores (the library's staging layout) and unaligned ones
        for (int variant = 0; variant < 5 && n; ++variant) {
            std::vector<uint8_t> c(b.begin(), b.begin() + n);
            for (size_t i = 0; i < n; ++i) if (bad[i >> 5] >> (i & 31) & 1u) c[i] = 'A';          // all ACGTU now
            size_t at = variant == 1 ? 0 : variant == 2 ? n / 2 : variant == 3 ? n - 1 : variant == 4 ? (n > 40 ? n - 40 : 0) : n;
            if (at < n) c[at] = 'N';
            std::vector<uint32_t> lo4, hi4, bad4;
            restate(c.data(), n, lo4, hi4, bad4);
            for (int aligned = 0; aligned < 2; ++aligned) {
                std::vector<uint32_t> buf(3 * (nw + 32) + 16, 0xDEADBEEF);
                uint32_t *l = reinterpret_cast<uint32_t *>((reinterpret_cast<uintptr_t>(buf.data()) + 63) & ~(uintptr_t)63) + (aligned ? 0 : 1);
                uint32_t *h = l + (nw + 15) / 16 * 16, *bd = h + (nw + 15) / 16 * 16;
                const bool dirty = rb::pack_bases_lazy(c.data(), n, l, h, bd, aligned != 0);
                bool ok = dirty == (at < n) && !std::memcmp(lo4.data(), l, nw * 4) && !std::memcmp(hi4.data(), h, nw * 4);
                if (dirty) ok = ok && !std::memcmp(bad4.data(), bd, nw * 4);
                if (!ok) { std::printf("LAZY MISMATCH n=%zu variant=%d aligned=%d\n", n, variant, aligned); ++fails; }
            }
        }
        // through the pool, in 128 K-base tasks, with a poll that checks the prefix is monotone
        const size_t task = 131072, nt = (n + task - 1) / task;
        std::vector<uint32_t> l3(nw + 1, 0), h3(nw + 1, 0), b3(nw + 1, 0);
        size_t last = 0, bad_poll = 0;
        rb::parallel_tasks(nt, [&](size_t t) {
            const size_t o = t * task, m = std::min(task, n - o);
            rb::pack_bases(b.data() + o, m, l3.data() + o / 32, h3.data() + o / 32, b3.data() + o / 32);
        }, [&](size_t done) { if (done < last || done > nt) ++bad_poll; last = done; });
        if (bad_poll || std::memcmp(lo.data(), l3.data(), nw * 4) || std::memcmp(hi.data(), h3.data(), nw * 4) ||
            std::memcmp(bad.data(), b3.data(), nw * 4)) {
            std::printf("POOL MISMATCH n=%zu\n", n);
            ++fails;
        }
    }
    std::printf("host_pack %s isa=%d threads=%d\n", fails ? "FAILED" : "OK", rb::pack_isa(), rb::host_threads());
    return fails ? 1 : 0;
}
