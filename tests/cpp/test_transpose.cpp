// rb::transpose32 (ibf_transpose.cuh) against the definition: bit i of y[j] == bit j of x[i].
#include "../../readbouncer_b200/csrc/ibf_transpose.cuh"
#include <cstdio>

static uint64_t s = 0x9E3779B97F4A7C15ULL;
static uint32_t rnd()
{
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    return (uint32_t)(s >> 16);
}

int main()
{
    for (int it = 0; it < 2000; ++it) {
        uint32_t x[32], y[32];
        for (int i = 0; i < 32; ++i) {
            x[i] = it == 0 ? 0u : it == 1 ? ~0u : it < 34 ? 1u << (it - 2) : it < 66 ? (i == it - 34 ? ~0u : 0u) : rnd() & (it & 1 ? rnd() : ~0u);
            y[i] = x[i];
        }
        rb::transpose32(y);
        for (int i = 0; i < 32; ++i)
            for (int j = 0; j < 32; ++j)
                if (((y[j] >> i) & 1u) != ((x[i] >> j) & 1u)) { std::printf("MISMATCH it=%d i=%d j=%d\n", it, i, j); return 1; }
    }
    std::printf("ok\n");
    return 0;
}
