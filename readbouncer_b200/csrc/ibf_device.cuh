// ibf_device.cuh -- device-side helpers shared by the count kernels (internal).
#pragma once

#include "ibf_kernels.cuh"

namespace rb {

// ------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t warp_max_u64(uint64_t v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        uint64_t other = __shfl_xor_sync(0xffffffffu, v, o);
        v = other > v ? other : v;
    }
    return v;
}

// fold a key into a shared key array: device scope for the partial summaries of one launch, system scope when the array
// is shared by the kernels of several devices (CountArgs::keys_shared; NVLink peer atomics)
__device__ __forceinline__ void key_max(uint64_t *dst, uint64_t key, int shared_by_devices)
{
    if (shared_by_devices) atomicMax_system(reinterpret_cast<unsigned long long *>(dst), (unsigned long long)key);
    else atomicMax(reinterpret_cast<unsigned long long *>(dst), (unsigned long long)key);
}

__device__ __forceinline__ uint32_t read_flag_of(uint64_t len, uint32_t k)
{
    return len < k ? 1u : (len > 65535u ? 2u : 0u);
}

constexpr int kTileWarps = 8;
constexpr int kSegMax = 32;                    // k-mer positions per lane per chunk
constexpr int kChunkPos = 32 * kSegMax;        // positions per warp chunk
constexpr int kDigBytes = kChunkPos + 32;      // + (k - 1), k <= 32

template <int WT, bool A16>
__device__ __forceinline__ void load_tile(const uint64_t *__restrict__ p, uint64_t (&v)[WT])
{
    if constexpr (A16 && (WT % 2 == 0)) {
#pragma unroll
        for (int i = 0; i < WT / 2; ++i) {
            ulonglong2 t = __ldg(reinterpret_cast<const ulonglong2 *>(p) + i);
            v[2 * i] = t.x;
            v[2 * i + 1] = t.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < WT; ++i) v[i] = __ldg(p + i);
    }
}

template <int WT>
__device__ __forceinline__ void count_bits(const uint64_t (&m)[WT], uint32_t *cnt)
{
#pragma unroll
    for (int w = 0; w < WT; ++w) {
        uint64_t x = m[w];
        while (x) {
            int b = __ffsll((long long)x) - 1;
            atomicAdd(&cnt[64 * w + b], 1u);
            x &= x - 1;
        }
    }
}


// Per-read epilogue shared by the warp-per-read kernels: dense counts, threshold test
// (select_matches), max over passing bins (max_matches) and its lowest bin, then one key per
// threshold table.  Re-zeroes the shared-memory counters for the next read.
template <int WT>
__device__ __forceinline__ void tile_epilogue(const CountArgs &a, uint64_t read, uint64_t len, uint32_t flag,
                                              uint32_t wc, uint32_t *cntF, uint32_t *cntR, int lane, int multi_tile)
{
    uint64_t best[kMaxLut];
    uint32_t thr[kMaxLut];
#pragma unroll
    for (int t = 0; t < kMaxLut; ++t) {
        best[t] = 0;
        thr[t] = (t < (int)a.n_lut && flag == 0) ? (uint32_t)__ldg(a.lut + (size_t)t * kLutSize + len) : 0x10000u;
    }
    for (int b = lane; b < 64 * WT; b += 32) {
        const uint64_t lb = (uint64_t)wc * 64 + b;
        const uint32_t f = cntF[b], r = cntR[b];
        cntF[b] = 0;
        cntR[b] = 0;
        if (lb < a.fv.n_bins_local) {
            if (a.counts_fwd) a.counts_fwd[read * a.fv.n_bins_local + lb] = (uint16_t)f;
            if (a.counts_rev) a.counts_rev[read * a.fv.n_bins_local + lb] = (uint16_t)r;
            const uint32_t m = max(f, r);
#pragma unroll
            for (int t = 0; t < kMaxLut; ++t)
                if (f >= thr[t] || r >= thr[t]) {
                    uint64_t key = pack_key(m, (uint32_t)(a.fv.bin_begin + lb));
                    best[t] = key > best[t] ? key : best[t];
                }
        }
    }
#pragma unroll
    for (int t = 0; t < kMaxLut; ++t) {
        if (t < (int)a.n_lut) {
            uint64_t bk = warp_max_u64(best[t]);
            if (lane == 0) {
                uint64_t *dst = a.keys + (size_t)t * a.n_reads + read;
                if (multi_tile || a.keys_shared) { if (bk) key_max(dst, bk, a.keys_shared); }
                else *dst = bk;
            }
        }
    }
}

}  // namespace rb
