// ibf_common.cuh -- shared host/device definitions of the B200 IBF engine.
//
// Restates the SeqAn-2 InterleavedBloomFilter arithmetic that ReadBouncer's
// src/IBF wraps (src/IBF/IBF.hpp:92-94; spec: SURVEY.md Appendix A):
//   k-mer value  H   = sum ord(s[j]) * 5^(k-1-j)  (mod 2^64), Dna5 ranks A0 C1 G2 T3 N4
//   row          r_i = ((pre_i * H) ^ ((pre_i * H) >> 27)) mod noOfBlocks,  pre_i = i ^ (k * seed)
//   bit (r, b) lives at bit  r * 64*binWidth + b  of a little-endian uint64 array.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace rb {

constexpr uint64_t kSeed = 0x90b45d39fb6da1faULL;
constexpr int kShift = 27;
constexpr uint64_t kInv5 = 0xCCCCCCCCCCCCCCCDULL;  // 5^-1 mod 2^64
constexpr int kMaxHash = 16;
constexpr int kMaxLut = 4;
constexpr uint32_t kLutSize = 65536;

#ifdef __CUDACC__
#define RB_HD __host__ __device__ __forceinline__
#else
#define RB_HD inline
#endif

RB_HD uint64_t mulhi64(uint64_t a, uint64_t b)
{
#ifdef __CUDA_ARCH__
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

// magic for fast_mod: floor((2^64 - 1) / d), d >= 1
inline uint64_t mod_magic(uint64_t d) { return ~0ULL / d; }

// Exact v mod d for any 64-bit v and d >= 1.  With m = floor((2^64-1)/d) the
// estimate q' = hi64(v*m) is floor(v/d) or one less, so one conditional
// subtraction finishes it (checked exhaustively on edge values in tests).
RB_HD uint64_t fast_mod(uint64_t v, uint64_t d, uint64_t magic)
{
    uint64_t q = mulhi64(v, magic);
    uint64_t r = v - q * d;
    return r >= d ? r - d : r;
}

struct HashParams {
    uint64_t n_blocks;        // rows (noOfBlocks); the modulus
    uint64_t magic;           // mod_magic(n_blocks)
    uint64_t pre[kMaxHash];   // i ^ (k * seed)
    uint64_t top;             // 5^(k-1) mod 2^64
    uint32_t k;
    uint32_t n_hash;
};

inline HashParams make_hash_params(uint64_t n_blocks, uint32_t k, uint32_t n_hash)
{
    HashParams hp{};
    hp.n_blocks = n_blocks;
    hp.magic = mod_magic(n_blocks);
    hp.k = k;
    hp.n_hash = n_hash;
    hp.top = 1;
    for (uint32_t j = 1; j < k; ++j) hp.top *= 5;
    for (uint32_t i = 0; i < (uint32_t)kMaxHash; ++i) hp.pre[i] = (uint64_t)i ^ ((uint64_t)k * kSeed);
    return hp;
}

RB_HD uint64_t hash_row(uint64_t H, uint64_t pre, uint64_t n_blocks, uint64_t magic)
{
    uint64_t v = pre * H;
    v ^= v >> kShift;
    return fast_mod(v, n_blocks, magic);
}

// char -> Dna5 rank (SeqAn-2 translate table): A/a 0, C/c 1, G/g 2, T/t/U/u 3, else 4
RB_HD uint32_t dna5(uint32_t c)
{
    uint32_t x = c & 0xDFu;
    return x == 'A' ? 0u : x == 'C' ? 1u : x == 'G' ? 2u : (x == 'T' || x == 'U') ? 3u : 4u;
}

// complement on Dna5 ranks; N stays N
RB_HD uint32_t comp5(uint32_t d) { return d < 4u ? 3u - d : 4u; }

// packed per-read summary (see rb_ibf.h)
RB_HD uint64_t pack_key(uint32_t count, uint32_t global_bin)
{
    return (1ULL << 48) | ((uint64_t)(count & 0xFFFFu) << 32) | (uint64_t)(~global_bin);
}

}  // namespace rb
