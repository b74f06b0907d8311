// ibf_postings_layout.cuh -- order of the bin ids inside one postings list (host/device).
//
// The lookup kernel (ibf_postings.cu) counts with shared-memory atomics on packed 8-bit counters, four bins per
// 32-bit word: the bank of bin id is (id >> 2) & 31.  One ATOMS instruction of a warp serves one "group" of up
// to 32 ids, and its cost is the largest number of ids of the group that fall into one bank.  With ids in
// ascending order the banks of a group are random (3.5-way conflicts on average for 32 lanes, measured 2.9
// wavefronts per ATOMS, profiles/r1_g_postings_cfg3_ncu_full.json).  The ORDER of a list is free -- counting is
// commutative -- so the build deals the ids, sorted by bank, round-robin over the groups of the list: a bank
// with no more ids than the list has groups puts at most one id into each.
//
// How the kernel walks a list of n_u 16-byte units (8 ids each), R = n_u / 32, tu = n_u % 32:
//   full round r < R   lane l loads unit 32 r + l (LDG.128); group (r, e) = the e-th id of every lane, e < 8
//   tail (tu units)    E = 1, 2, 4 or 8 ids per lane (tu <= 4, 8, 16, 31): lane l < 8 tu / E loads the 2 E bytes at
//                      tail position l * E, group e = the e-th id of every lane -- so a short tail costs E ATOMS
//                      instructions with most lanes busy instead of 8 with tu lanes busy
// The real ids of a list occupy positions [0, n); the sentinel (n_bins_local) fills [n, 8 n_u).
#pragma once

#include "ibf_common.cuh"

namespace rb {

struct ListShape {
    uint32_t n_big;      // groups of the full rounds (8 R), 32 slots each
    uint32_t n_all;      // + E tail groups
    uint32_t E;          // ids per lane in the tail (0: no tail)
    uint32_t q;          // every tail group has at least q slots ...
    uint32_t n_a;        // ... so the first q * n_all ids are dealt over all groups,
    uint32_t n_b;        // the next n_big + (tail ids % E) over the groups that have a slot q,
    uint32_t tail0;      // and the rest over the full rounds' groups.  tail0 = first position of the tail
};

RB_HD uint32_t list_tail_ids_per_lane(uint32_t tail_units)
{
    return tail_units == 0 ? 0u : tail_units <= 4 ? 1u : tail_units <= 8 ? 2u : tail_units <= 16 ? 4u : 8u;
}

RB_HD ListShape list_shape(uint32_t n)          // n = real ids of the list
{
    ListShape s;
    const uint32_t n_u = (n + 7u) >> 3, R = n_u >> 5;
    s.E = list_tail_ids_per_lane(n_u & 31u);
    s.tail0 = 256u * R;
    const uint32_t m_tail = n - s.tail0;
    s.q = s.E ? m_tail / s.E : 0u;
    s.n_big = 8u * R;
    s.n_all = s.n_big + s.E;
    s.n_a = s.q * s.n_all;
    s.n_b = s.n_big + (s.E ? m_tail % s.E : 0u);
    return s;
}

// position (index into the list's ids) of the c-th id in dealing order, c < n
RB_HD uint32_t list_position(const ListShape &s, uint32_t c)
{
    uint32_t slot, g;
    if (c < s.n_a) { slot = c / s.n_all; g = c % s.n_all; }
    else if (c < s.n_a + s.n_b) { slot = s.q; g = c - s.n_a; }
    else { const uint32_t c2 = c - s.n_a - s.n_b; slot = s.q + 1u + c2 / s.n_big; g = c2 % s.n_big; }
    return g < s.n_big ? (g >> 3) * 256u + slot * 8u + (g & 7u) : s.tail0 + slot * s.E + (g - s.n_big);
}

RB_HD uint32_t counter_bank(uint32_t id) { return (id >> 2) & 31u; }

}  // namespace rb
