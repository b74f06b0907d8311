// rb::slot_shape / rb::slot_position / rb::slot_pad_id (ibf_postings_layout.cuh): the dealing order of the ids inside a
// postings SLOT is a bijection onto [0, n) for every list length a slot can hold, ids sorted by bank are spread over the
// groups the lookup kernel walks (group (r, e) = positions 128 r + 4 l + e, one ATOMS instruction), and the pads of one
// group fall into 32 different banks behind the bins.
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

static uint32_t group_of(uint32_t pos) { return (pos >> 7) * 4 + (pos & 3); }

int main()
{
    double sum_sorted = 0, sum_dealt = 0;
    uint64_t groups = 0;
    for (uint32_t n = 1; n <= rb::slot_capacity(rb::kSlotMaxBytes); ++n) {
        const rb::SlotShape sh = rb::slot_shape(n);
        std::vector<uint8_t> seen(n, 0);
        for (uint32_t c = 0; c < n; ++c) {
            const uint32_t p = rb::slot_position(sh, c);
            if (p >= n || seen[p]) { std::printf("NOT A BIJECTION n=%u c=%u p=%u\n", n, c, p); return 1; }
            seen[p] = 1;
        }
        if (n % 7 != 0 && n > 600) continue;
        std::vector<uint32_t> ids;
        while (ids.size() < n) ids.push_back(rnd() % 31008);
        std::sort(ids.begin(), ids.end());
        for (size_t i = 1; i < ids.size(); ++i) if (ids[i] <= ids[i - 1]) ids[i] = ids[i - 1] + 1;
        std::vector<uint32_t> by_bank(ids);
        std::stable_sort(by_bank.begin(), by_bank.end(), [](uint32_t a, uint32_t b) { return rb::counter_bank(a) < rb::counter_bank(b); });
        std::vector<uint32_t> dealt(n);
        for (uint32_t c = 0; c < n; ++c) dealt[rb::slot_position(sh, c)] = by_bank[c];
        const uint32_t n_groups = group_of(n - 1) / 4 * 4 + 4;
        auto degree_sum = [&](const std::vector<uint32_t> &v, uint32_t *worst) {
            std::vector<uint32_t> cnt((size_t)n_groups * 32, 0);
            for (uint32_t p = 0; p < n; ++p) ++cnt[(size_t)group_of(p) * 32 + rb::counter_bank(v[p])];
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
        const uint32_t min_groups = sh.n_big ? sh.n_big : 4;
        const uint32_t bound = (max_load + min_groups - 1) / min_groups + 1;
        if (w1 > bound) { std::printf("CONFLICT DEGREE n=%u worst=%u bound=%u\n", n, w1, bound); return 1; }
    }
    // pads: the 32 lanes of one instruction (positions 4 l + e of a round) hit 32 different counter words / banks
    for (uint32_t sentinel : {16u, 31008u, 65280u})
        for (uint32_t p0 = 0; p0 < 2048; p0 += 128)
            for (uint32_t e = 0; e < 4; ++e) {
                uint32_t banks = 0;
                for (uint32_t l = 0; l < 32; ++l) {
                    const uint32_t id = rb::slot_pad_id(sentinel, p0 + 4 * l + e);
                    if (id < sentinel || id >= sentinel + 128 || id > 0xFFFF) { std::printf("PAD OUT OF RANGE\n"); return 1; }
                    banks |= 1u << rb::counter_bank(id);
                }
                if (banks != 0xFFFFFFFFu) { std::printf("PADS SHARE A BANK\n"); return 1; }
            }
    std::printf("mean wavefronts per group: ascending %.3f dealt %.3f\n", sum_sorted / groups, sum_dealt / groups);
    if (!(sum_dealt < 0.6 * sum_sorted)) { std::printf("NO GAIN\n"); return 1; }
    std::printf("ok\n");
    return 0;
}
