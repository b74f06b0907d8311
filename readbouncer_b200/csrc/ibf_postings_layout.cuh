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

// ---- slot layout (count_slots_kernel): every k-mer owns a fixed, 128-byte-aligned slot ------------------------------
// A slot is fetched whole by ONE bulk copy into shared memory, then walked by the warp in rounds: in round r lane l
// takes the four ids at positions 128 r + 4 l + e (one LDS.64), so one ATOMS instruction serves group (r, e) = the
// positions {128 r + 4 l + e : l < 32}.  The n real ids occupy positions [0, n); as above the build deals them, sorted
// by bank, round-robin over the groups, level by level (level t = lane t of every group that has one).
struct SlotShape {
    uint32_t n_big;      // groups of the full rounds (4 per round), 32 positions each
    uint32_t n_all;      // + the 4 groups of the last, partial round
    uint32_t q;          // every tail group has at least q positions
    uint32_t n_a;        // ids dealt over all groups (levels < q)
    uint32_t n_b;        // ids of level q: the full rounds' groups + the (n mod 4) tail groups that have one more
    uint32_t tail0;      // first position of the partial round
};

RB_HD SlotShape slot_shape(uint32_t n)
{
    SlotShape s;
    const uint32_t R = n >> 7, m = n & 127u;
    s.n_big = 4u * R;
    s.n_all = s.n_big + 4u;
    s.tail0 = 128u * R;
    s.q = m >> 2;
    s.n_a = s.q * s.n_all;
    s.n_b = s.n_big + (m & 3u);
    return s;
}

// position of the c-th id in dealing order, c < n
RB_HD uint32_t slot_position(const SlotShape &s, uint32_t c)
{
    uint32_t t, g;
    if (c < s.n_a) { t = c / s.n_all; g = c % s.n_all; }
    else if (c < s.n_a + s.n_b) { t = s.q; g = c - s.n_a; }
    else { const uint32_t c2 = c - s.n_a - s.n_b; t = s.q + 1u + c2 / s.n_big; g = c2 % s.n_big; }
    return g < s.n_big ? (g >> 2) * 128u + 4u * t + (g & 3u) : s.tail0 + 4u * t + (g - s.n_big);
}

// Padding id for the positions [n, end of the last round): lane l's pads live in their own counter word behind the bins
// (32 words, one per bank), so a group of pads costs one conflict-free ATOMS instead of a 32-way collision.
RB_HD uint32_t slot_pad_id(uint32_t sentinel, uint32_t pos) { return sentinel + 4u * ((pos >> 2) & 31u); }

constexpr uint32_t kSlotHeaderBytes = 8;      // u16 n (kSlotOverflow: the list lives in the overflow area), u16 0, u32 first overflow unit;
                                              // overflow slots: bytes 8..11 = u32 number of 16-byte units
constexpr uint32_t kSlotOverflow = 0xFFFFu;
constexpr uint32_t kSlotMaxBytes = 4096;
RB_HD uint32_t slot_capacity(uint32_t slot_bytes) { return (slot_bytes - kSlotHeaderBytes) / 2u; }

}  // namespace rb
