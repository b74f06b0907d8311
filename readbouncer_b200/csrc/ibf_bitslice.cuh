// ibf_bitslice.cuh -- bit-sliced (vertical) counters in registers, shared by the k-mer table kernels.
//
// A lane adds one-bit-per-bin masks into bit planes with carry-save adders, so the cost does not depend
// on how many bits are set; the 32 lanes are then summed by an exchange-and-halve butterfly of
// bit-sliced full adders (shuffle distance 16, 8, 4, 2, 1): every level a lane keeps one half of its span
// and receives its partner's copy of that half, so after five levels lane l holds the 9-plane counts of
// NWP consecutive mask bits.
#pragma once

#include "ibf_device.cuh"

namespace rb {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (c & (a ^ b)); }

// fold the upper/lower half of the WORDS held by a lane pair (distance OFF); NP -> NP+1 planes
template <int NW, int NP, int OFF>
__device__ __forceinline__ void fold_words(const uint32_t (&in)[NP][NW], uint32_t (&out)[NP + 1][NW / 2], int lane)
{
    const bool upper = (lane & OFF) != 0;
    uint32_t carry[NW / 2];
#pragma unroll
    for (int i = 0; i < NW / 2; ++i) carry[i] = 0;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
#pragma unroll
        for (int i = 0; i < NW / 2; ++i) {
            const uint32_t lo = in[p][i], hi = in[p][NW / 2 + i];
            const uint32_t mine = upper ? hi : lo;
            const uint32_t recv = __shfl_xor_sync(kFull, upper ? lo : hi, OFF);
            out[p][i] = mine ^ recv ^ carry[i];
            carry[i] = maj3(mine, recv, carry[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < NW / 2; ++i) out[NP][i] = carry[i];
}

// fold the upper/lower half of the BITS significant bits of a single word
template <int BITS, int NP, int OFF>
__device__ __forceinline__ void fold_bits(const uint32_t (&in)[NP], uint32_t (&out)[NP + 1], int lane)
{
    constexpr int H = BITS / 2;
    constexpr uint32_t LOW = (1u << H) - 1u;
    const bool upper = (lane & OFF) != 0;
    uint32_t carry = 0;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const uint32_t lo = in[p] & LOW, hi = (in[p] >> H) & LOW;
        const uint32_t mine = upper ? hi : lo;
        const uint32_t recv = __shfl_xor_sync(kFull, upper ? lo : hi, OFF);
        out[p] = mine ^ recv ^ carry;
        carry = maj3(mine, recv, carry);
    }
    out[NP] = carry;
}

// 32-lane sum of 4-plane per-lane counters -> 9-plane counters of NWP bits per lane
template <int NWP>
__device__ __forceinline__ void warp_fold(const uint32_t (&pl)[4][NWP], uint32_t (&res)[9], int lane)
{
    if constexpr (NWP == 8) {
        uint32_t a[5][4], b[6][2], c[7][1], d[7], e[8];
        fold_words<8, 4, 16>(pl, a, lane);
        fold_words<4, 5, 8>(a, b, lane);
        fold_words<2, 6, 4>(b, c, lane);
#pragma unroll
        for (int p = 0; p < 7; ++p) d[p] = c[p][0];
        fold_bits<32, 7, 2>(d, e, lane);
        fold_bits<16, 8, 1>(e, res, lane);
    } else {   // NWP == 4
        uint32_t a[5][2], b[6][1], c[6], d[7], e[8];
        fold_words<4, 4, 16>(pl, a, lane);
        fold_words<2, 5, 8>(a, b, lane);
#pragma unroll
        for (int p = 0; p < 6; ++p) c[p] = b[p][0];
        fold_bits<32, 6, 4>(c, d, lane);
        fold_bits<16, 7, 2>(d, e, lane);
        fold_bits<8, 8, 1>(e, res, lane);
    }
}

template <int NP>
__device__ __forceinline__ uint32_t bs_ge(const uint32_t (&pl)[NP], uint32_t thr)
{
    if (thr >> NP) return 0;
    uint32_t gt = 0, eq = ~0u;
#pragma unroll
    for (int p = NP - 1; p >= 0; --p) {
        const uint32_t tb = ((thr >> p) & 1u) ? ~0u : 0u;
        gt |= eq & pl[p] & ~tb;
        eq &= ~(pl[p] ^ tb);
    }
    return gt | eq;
}

// max over the counters selected by `sel` (bit-sliced numbers, MSB first); `sel` shrinks to the arg-max set
template <int NP>
__device__ __forceinline__ uint32_t bs_max(const uint32_t (&pl)[NP], uint32_t &sel)
{
    uint32_t val = 0;
#pragma unroll
    for (int p = NP - 1; p >= 0; --p) {
        const uint32_t t = sel & pl[p];
        if (t) { sel = t; val |= 1u << p; }
    }
    return val;
}

template <int NP>
__device__ __forceinline__ uint32_t bs_get(const uint32_t (&pl)[NP], int b)
{
    uint32_t c = 0;
#pragma unroll
    for (int p = 0; p < NP; ++p) c |= ((pl[p] >> b) & 1u) << p;
    return c;
}

}  // namespace rb
