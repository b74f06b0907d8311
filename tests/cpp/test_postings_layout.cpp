// rb::list_shape / rb::list_position (ibf_postings_layout.cuh): the dealing order is a bijection onto [0, n) for every
// list length, and ids sorted by bank are spread over the groups the lookup kernel walks (at most
// ceil(bank load / groups) ids of one bank per group while all groups take part).
#include "../../readbouncer_b200/csrc/ibf_postings_layout.cuh"
#include <algorithm>
#include <cstdio>
#include <vector>

static uint64_t s = 0x9E3779B97F4A7C15ULL;
static uint32_t rnd()
{
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    return (uint32_t)(s >> 16);
}

// group of a position, as the kernel walks the list: returns a group number unique within the list
static uint32_t group_of(uint32_t n, uint32_t pos)
{
    const uint32_t n_u = (n + 7) / 8, R = n_u / 32, tu = n_u % 32, E = rb::list_tail_ids_per_lane(tu);
    if (pos < 256 * R) return (pos / 256) * 8 + pos % 8;
    return 8 * R + (pos - 256 * R) % E;
}

int main()
{
    double sum_sorted = 0, sum_dealt = 0;
    uint64_t groups = 0;
    for (uint32_t n = 1; n <= 3000; ++n) {
        const rb::ListShape sh = rb::list_shape(n);
        std::vector<uint8_t> seen(n, 0);
        for (uint32_t c = 0; c < n; ++c) {
            const uint32_t p = rb::list_position(sh, c);
            if (p >= n || seen[p]) { std::printf("NOT A BIJECTION n=%u c=%u p=%u\n", n, c, p); return 1; }
            seen[p] = 1;
        }
        // lanes of the tail never exceed a warp
        const uint32_t n_u = (n + 7) / 8, tu = n_u % 32;
        if (sh.E && 8 * tu / sh.E > 32) { std::printf("TAIL TOO WIDE n=%u\n", n); return 1; }
        if (n % 7 != 0 && n > 600) continue;                       // conflict statistics on a subset
        // random distinct ids out of 31 008 bins, ascending
        std::vector<uint32_t> ids;
        while (ids.size() < n) ids.push_back(rnd() % 31008);
        std::sort(ids.begin(), ids.end());
        for (size_t i = 1; i < ids.size(); ++i) if (ids[i] <= ids[i - 1]) ids[i] = ids[i - 1] + 1;
        std::vector<uint32_t> by_bank(ids);
        std::stable_sort(by_bank.begin(), by_bank.end(), [](uint32_t a, uint32_t b) { return rb::counter_bank(a) < rb::counter_bank(b); });
        std::vector<uint32_t> dealt(n);
        for (uint32_t c = 0; c < n; ++c) dealt[rb::list_position(sh, c)] = by_bank[c];
        const uint32_t n_groups = sh.n_all;
        auto degree_sum = [&](const std::vector<uint32_t> &v, uint32_t *worst) {
            std::vector<uint32_t> cnt((size_t)n_groups * 32, 0);
            for (uint32_t p = 0; p < n; ++p) ++cnt[(size_t)group_of(n, p) * 32 + rb::counter_bank(v[p])];
            double t = 0;
            *worst = 0;
            for (uint32_t g = 0; g < n_groups; ++g) {
                uint32_t m = 0;
                for (int b = 0; b < 32; ++b) m = std::max(m, cnt[(size_t)g * 32 + b]);
                t += m;
                *worst = std::max(*worst, m);
            }
            return t;
        };
        uint32_t w0, w1;
        sum_sorted += degree_sum(ids, &w0);
        sum_dealt += degree_sum(dealt, &w1);
        groups += n_groups;
        uint32_t load[32] = {0}, max_load = 0;
        for (uint32_t id : ids) max_load = std::max(max_load, ++load[rb::counter_bank(id)]);
        const uint32_t min_groups = sh.n_big ? sh.n_big : n_groups;       // the fewest groups any id range is dealt over
        const uint32_t bound = (max_load + min_groups - 1) / min_groups + 1;   // + 1: a bank's run may cross a regime boundary
        if (w1 > bound) { std::printf("CONFLICT DEGREE n=%u worst=%u bound=%u\n", n, w1, bound); return 1; }
    }
    std::printf("mean wavefronts per group: ascending %.3f dealt %.3f\n", sum_sorted / groups, sum_dealt / groups);
    if (!(sum_dealt < 0.6 * sum_sorted)) { std::printf("NO GAIN\n"); return 1; }
    std::printf("ok\n");
    return 0;
}
