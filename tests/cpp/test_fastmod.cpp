// Exhaustive-on-edges check of rb::fast_mod (ibf_common.cuh) against the C `%` operator.
#include "../../readbouncer_b200/csrc/ibf_common.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>

static uint64_t rng_state = 0x1234567;
static uint64_t rnd()
{
    rng_state += 0x9E3779B97F4A7C15ULL;
    uint64_t z = rng_state;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

int main()
{
    std::vector<uint64_t> ds = {1, 2, 3, 5, 7, 64, 1236245, 1236269, 51929353, 0xFFFFFFFFull, 0x100000000ull,
                                0x100000001ull, (1ull << 63) - 1, 1ull << 63, (1ull << 63) + 1, ~0ull - 1, ~0ull};
    for (int i = 0; i < 2000; ++i) ds.push_back((rnd() >> (rnd() % 64)) | 1);
    uint64_t checked = 0;
    for (uint64_t d : ds) {
        const uint64_t m = rb::mod_magic(d);
        std::vector<uint64_t> vs = {0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, 2 * d + 1, ~0ull, ~0ull - 1, ~0ull - d, ~0ull / d * d,
                                    ~0ull / d * d - 1, ~0ull / d * d + 1, 1ull << 63, (1ull << 63) - 1};
        for (int i = 0; i < 3000; ++i) vs.push_back(rnd() >> (rnd() % 64));
        for (uint64_t q = 0; q < 64; ++q) { vs.push_back(q * d); vs.push_back(q * d + d - 1); vs.push_back((~0ull / d - q) * d); }
        for (uint64_t v : vs) {
            if (rb::fast_mod(v, d, m) != v % d) { std::printf("MISMATCH v=%llu d=%llu\n", (unsigned long long)v, (unsigned long long)d); return 1; }
            ++checked;
        }
    }
    // hash_row equals the reference arithmetic
    rb::HashParams hp = rb::make_hash_params(1236269, 13, 3);
    for (int i = 0; i < 100000; ++i) {
        uint64_t H = rnd() % 1220703125ull;
        for (int j = 0; j < 3; ++j) {
            uint64_t v = (((uint64_t)j) ^ (13ull * 0x90b45d39fb6da1faULL)) * H;
            v ^= v >> 27;
            if (rb::hash_row(H, hp.pre[j], hp.n_blocks, hp.magic) != v % 1236269ull) { std::puts("hash_row mismatch"); return 1; }
        }
    }
    std::printf("fast_mod OK (%llu cases)\n", (unsigned long long)checked);
    return 0;
}
