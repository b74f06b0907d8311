// ibf_ctable.cu -- one-k-mer table for MEDIUM filters (rows of 3..32 words, 129..2048 bins), entries loaded by lane groups.
//
// Same tabulated function as ibf_table.cu -- what seqan::count does per k-mer and strand (src/IBF/IBFClassify.cpp:149-150,
// SURVEY.md App. A.6):
//
//     table[x] = [ AND_i row(h_i(x)) : G words ][ AND_i row(h_i(revcomp(x))) : G words ]        x in [0, 4^k)
//
// with the row padded to G = 4, 8, 16 or 32 words, so an entry is G pieces of 16 bytes = 64 .. 512 bytes, aligned to its
// size.  What differs is who loads it.  A lane that fetches a whole 64..256-byte entry issues 4..16 separate 16-byte loads,
// i.e. 4..16 requests per k-mer position; here G ADJACENT LANES load one entry with ONE instruction (lane p takes piece p),
// which the memory system serves as one request per 128-byte line (profiles/r1_d_gather_sweep3_coop.jsonl) -- one or two
// requests per position and both strands, against two dependent ones per position AND strand for the postings lists these
// filters used before (ibf_postings.cu).  A lane then owns two words of one strand's mask for every position of its group
// and bumps the counters of their set bits in shared memory (a mask of a 1000-bin filter has ~10 false-positive bits, so
// this is a handful of atomics per entry, spread over the group).
//
// Windows with a non-ACGT base are not in the table: their lanes evaluate their two words from the original bit matrix
// (three row probes), so every output bit is the reference's.
#include "ibf_bitslice.cuh"

#include <cstdlib>

namespace rb {

namespace {

// ------------------------------------------------------------------------------------------
// build: one thread per (k-mer, piece)
// ------------------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(256) ctable_build_kernel(const FilterView fv, uint4 *__restrict__ table, const uint64_t n_kmers)
{
    const HashParams &hp = fv.hp;
    const uint32_t k = hp.k;
    const uint64_t n = n_kmers * G;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t x = t / G;
        const uint32_t p = (uint32_t)(t % G);
        const bool rev = p >= G / 2;
        const uint32_t w0 = 2u * (p % (G / 2));
        uint64_t H = 0, pw = 1;
        for (uint32_t j = 0; j < k; ++j) {
            const uint32_t d = (uint32_t)(x >> (2 * (k - 1 - j))) & 3u;          // base j of the k-mer
            if (!rev) H = H * 5 + d;
            else { H += (uint64_t)(3u - d) * pw; pw *= 5; }
        }
        uint64_t m0 = w0 < fv.stride ? ~0ULL : 0ULL, m1 = w0 + 1 < fv.stride ? ~0ULL : 0ULL;
        for (uint32_t i = 0; i < hp.n_hash; ++i) {
            const uint64_t *row = fv.words + hash_row(H, hp.pre[i], hp.n_blocks, hp.magic) * fv.stride;
            if (w0 < fv.stride) m0 &= __ldg(row + w0);
            if (w0 + 1 < fv.stride) m1 &= __ldg(row + w0 + 1);
        }
        table[t] = make_uint4((uint32_t)m0, (uint32_t)(m0 >> 32), (uint32_t)m1, (uint32_t)(m1 >> 32));
    }
}

// ------------------------------------------------------------------------------------------
// count: warp per read, G lanes per entry, 32/G runs of consecutive positions per warp
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void bump_word(uint32_t *cnt, uint32_t w)
{
    while (w) {
        const int b = __ffs((int)w) - 1;
        atomicAdd(cnt + b, 1u);
        w &= w - 1;
    }
}

// the lane's two mask words of a window that is not in the table, from the original rows
__device__ __forceinline__ uint4 piece_hashed(const FilterView &fv, const uint8_t *dig, const bool rev, const uint32_t w0)
{
    const HashParams &hp = fv.hp;
    uint64_t H = 0, pw = 1;
    for (uint32_t u = 0; u < hp.k; ++u) {
        const uint32_t d = dig[u];
        if (!rev) H = H * 5 + d;
        else { H += comp5(d) * pw; pw *= 5; }
    }
    uint64_t m0 = w0 < fv.stride ? ~0ULL : 0ULL, m1 = w0 + 1 < fv.stride ? ~0ULL : 0ULL;
    for (uint32_t i = 0; i < hp.n_hash; ++i) {
        const uint64_t *row = fv.words + hash_row(H, hp.pre[i], hp.n_blocks, hp.magic) * fv.stride;
        if (w0 < fv.stride) m0 &= __ldg(row + w0);
        if (w0 + 1 < fv.stride) m1 &= __ldg(row + w0 + 1);
    }
    return make_uint4((uint32_t)m0, (uint32_t)(m0 >> 32), (uint32_t)m1, (uint32_t)(m1 >> 32));
}

constexpr int kCtWarps = 8;

template <int G, int U>
__global__ void __launch_bounds__(kCtWarps * 32) count_ctable_kernel(const CountArgs a, const uint4 *__restrict__ table)
{
    constexpr int NG = 32 / G;                                          // entries per warp-wide load
    extern __shared__ __align__(16) uint32_t s_ct[];
    uint32_t *const s_cnt = s_ct;                                       // [kCtWarps][2][64 * G]
    uint8_t *const s_dig = reinterpret_cast<uint8_t *>(s_ct + kCtWarps * 2 * 64 * G);   // [kCtWarps][kDigBytes]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t total_warps = (uint64_t)gridDim.x * kCtWarps;
    const uint32_t k = a.fv.hp.k;
    const uint64_t kmask = (k >= 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
    uint8_t *const dig = s_dig + warp * kDigBytes;
    uint32_t *const cntF = s_cnt + warp * (2 * 64 * G), *const cntR = cntF + 64 * G;
    const uint32_t p = (uint32_t)lane % G, gi = (uint32_t)lane / G;
    const bool rev = p >= G / 2;
    const uint32_t w0 = 2u * (p % (G / 2));
    uint32_t *const cnt = (rev ? cntR : cntF) + 64 * w0;               // the 128 counters of my two words

    for (int b = lane; b < 2 * 64 * G; b += 32) cntF[b] = 0;
    __syncwarp();

    for (uint64_t read = (uint64_t)blockIdx.x * kCtWarps + warp; read < a.n_reads; read += total_warps) {
        const uint64_t off = a.read_off[read];
        const uint64_t len = a.read_off[read + 1] - off;
        const uint32_t flag = read_flag_of(len, k);
        if (lane == 0 && a.read_flag) a.read_flag[read] = (uint8_t)flag;

        if (flag == 0) {
            const uint32_t npos = (uint32_t)len - k + 1;
            for (uint32_t cs = 0; cs < npos; cs += kChunkPos) {
                const uint32_t cn = min((uint32_t)kChunkPos, npos - cs);
                __syncwarp();
                for (uint32_t i = lane; i < cn + k - 1; i += 32) dig[i] = (uint8_t)dna5(a.bases[off + cs + i]);
                __syncwarp();
                const uint32_t seg = (cn + NG - 1) / NG;
                const uint32_t j0 = gi * seg;
                const uint32_t j1 = min(j0 + seg, cn);
                if (j0 < j1) {
                    uint64_t x = 0;          // 2-bit packed k-mer
                    uint32_t nbad = 0;       // non-ACGT bases inside it
#pragma unroll 1                 // (nvcc 12.9's cicc crashes when it unrolls this loop here)
                    for (uint32_t u = 0; u < k; ++u) {
                        const uint32_t d = dig[j0 + u];
                        x = (x << 2) | (d & 3u);
                        nbad += d >> 2;
                    }
                    x &= kmask;
                    for (uint32_t j = j0; j < j1; j += U) {
                        uint4 v[U];
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            v[u] = make_uint4(0, 0, 0, 0);
                            if (j + u < j1) {
                                if (nbad == 0) v[u] = __ldg(table + x * G + p);
                                else v[u] = piece_hashed(a.fv, dig + j + u, rev, w0);
                                if (j + u + 1 < j1) {
                                    const uint32_t dout = dig[j + u], din = dig[j + u + k];
                                    x = ((x << 2) | (din & 3u)) & kmask;
                                    nbad += (din >> 2) - (dout >> 2);
                                }
                            }
                        }
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            bump_word(cnt, v[u].x);
                            bump_word(cnt + 32, v[u].y);
                            bump_word(cnt + 64, v[u].z);
                            bump_word(cnt + 96, v[u].w);
                        }
                    }
                }
            }
        }
        __syncwarp();
        tile_epilogue<G>(a, read, len, flag, 0, cntF, cntR, lane, 0);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// count, bit-sliced: the same loads, counters in registers
// ------------------------------------------------------------------------------------------
// ncu on the atomic-counter kernel above (profiles/r2_v_ctable_atomic_w4_ncu.json): 6 000 warp instructions per 250-base
// read, 60 % of them in the bit loops -- the masks of a 200-bin filter built by the reference's sizing rule have ~17 set bits
// per entry, found one at a time by 9 lanes of 32.  Here a lane adds the four 32-bit words of its piece into 7 bit planes with
// carry-save adders (17 logic operations per 4 positions and word, whatever the masks hold; 127 positions per lane and chunk),
// and the lanes that own the same piece (lane bits log2(G) and up) are summed by the exchange-and-halve butterfly of
// ibf_bitslice.cuh: 3 / 2 / 1 levels for G = 4 / 8 / 16, after which a lane holds the 10 / 9 / 8-plane counts of 16 / 32 / 64
// bins of one strand.  The lane G/2 further on holds the other strand of the same bins: one more exchange and every lane
// evaluates select_matches / max_matches for its bins without ever unpacking a counter.
constexpr int kCbPlanes = 7;                       // per-lane planes: up to 127 positions per lane and chunk
constexpr int kCbCap = (1 << kCbPlanes) - 1;

template <int G> struct CbGeom {
    static constexpr int LV = G == 4 ? 3 : G == 8 ? 2 : G == 16 ? 1 : 0;   // butterfly levels
    static constexpr int NPF = kCbPlanes + LV;                   // planes after the fold
    static constexpr int RW = G == 32 ? 4 : G == 16 ? 2 : 1;     // 32-bit words of bins a lane ends up with
    static constexpr int B = G == 4 ? 16 : 32;                   // bins per such word
    static constexpr int NG = 32 / G;
    static constexpr int CHUNK = NG * kCbCap;                    // positions per warp chunk: 1016 / 508 / 254 / 127
    // accumulator planes when the caller promises short reads: one chunk, or (G = 32) two chunks
    static constexpr int NPS = G == 32 ? 8 : NPF;
    static constexpr int SHORT_POS = G == 32 ? 254 : CHUNK;      // positions such a read may have
};

template <int G, int NPA>
__global__ void __launch_bounds__(kCtWarps * 32, 2) count_ctable_bs_kernel(const CountArgs a, const uint4 *__restrict__ table)
{
    using Ge = CbGeom<G>;
    constexpr int NG = Ge::NG, RW = Ge::RW, B = Ge::B, NPF = Ge::NPF;
    static_assert(NPA >= NPF, "accumulator narrower than one chunk's counts");
    __shared__ __align__(16) uint8_t s_dig[kCtWarps][kDigBytes];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t total_warps = (uint64_t)gridDim.x * kCtWarps;
    const uint32_t k = a.fv.hp.k;
    const uint64_t kmask = (k >= 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
    uint8_t *const dig = s_dig[warp];
    const uint32_t p = (uint32_t)lane % G, gi = (uint32_t)lane / G;
    const bool rev = p >= G / 2;
    const uint32_t w0 = 2u * (p % (G / 2));
    // the bins this lane reports after the fold: 32-bit word i of its piece (upper half of the words for lane bit 4, ...)
    uint32_t bin0[RW];
    if constexpr (G == 32) {
#pragma unroll
        for (int r = 0; r < 4; ++r) bin0[r] = 64u * w0 + 32u * r;
    } else if constexpr (G == 16) {
        bin0[0] = 64u * w0 + 64u * ((lane >> 4) & 1u);
        bin0[1] = bin0[0] + 32u;
    } else {
        bin0[0] = 64u * w0 + 64u * ((lane >> 4) & 1u) + 32u * ((lane >> 3) & 1u) + (G == 4 ? 16u * ((lane >> 2) & 1u) : 0u);
    }
    const uint64_t nbl = a.fv.n_bins_local;

    for (uint64_t read = (uint64_t)blockIdx.x * kCtWarps + warp; read < a.n_reads; read += total_warps) {
        const uint64_t off = a.read_off[read];
        const uint64_t len = a.read_off[read + 1] - off;
        uint32_t flag = read_flag_of(len, k);
        // a read longer than the caller's max_read_len promised does not fit the NPA-bit accumulator: flag 3, not classified
        if (flag == 0 && NPA < 16 && len - k + 1 > (uint64_t)Ge::SHORT_POS) flag = 3;
        if (lane == 0 && a.read_flag) a.read_flag[read] = (uint8_t)flag;

        uint32_t acc[RW][NPA];
#pragma unroll
        for (int r = 0; r < RW; ++r)
#pragma unroll
            for (int q = 0; q < NPA; ++q) acc[r][q] = 0;

        if (flag == 0) {
            const uint32_t npos = (uint32_t)len - k + 1;
            for (uint32_t cs = 0; cs < npos; cs += Ge::CHUNK) {
                const uint32_t cn = min((uint32_t)Ge::CHUNK, npos - cs);
                __syncwarp();
                for (uint32_t i = lane; i < cn + k - 1; i += 32) dig[i] = (uint8_t)dna5(a.bases[off + cs + i]);
                __syncwarp();
                const uint32_t seg = (cn + NG - 1) / NG;                  // <= kCbCap
                const uint32_t j0 = gi * seg;
                const uint32_t j1 = min(j0 + seg, cn);
                uint32_t pl[kCbPlanes][4];
#pragma unroll
                for (int q = 0; q < kCbPlanes; ++q)
#pragma unroll
                    for (int w = 0; w < 4; ++w) pl[q][w] = 0;
                if (j0 < j1) {
                    uint64_t x = 0;          // 2-bit packed k-mer
                    uint32_t nbad = 0;       // non-ACGT bases inside it
#pragma unroll 1                 // (nvcc 12.9's cicc crashes when it unrolls this loop here)
                    for (uint32_t u = 0; u < k; ++u) {
                        const uint32_t d = dig[j0 + u];
                        x = (x << 2) | (d & 3u);
                        nbad += d >> 2;
                    }
                    x &= kmask;
                    // four entries in flight per lane.  (Issuing the next four before the current four are added was measured
                    // 4-6 % slower on all four widths, profiles/r2_z_ctable_prefetch.jsonl: 16 warps per SM x 4 entries already
                    // cover the latency; what is left is the address chain of the sliding k-mer and the ~3 800 instructions.)
                    auto fetch4 = [&](const uint32_t jb, uint4 (&v)[4]) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            v[u] = make_uint4(0, 0, 0, 0);
                            if (jb + u < j1) {
                                if (nbad == 0) v[u] = __ldg(table + x * G + p);
                                else v[u] = piece_hashed(a.fv, dig + jb + u, rev, w0);
                                if (jb + u + 1 < j1) {
                                    const uint32_t dout = dig[jb + u], din = dig[jb + u + k];
                                    x = ((x << 2) | (din & 3u)) & kmask;
                                    nbad += (din >> 2) - (dout >> 2);
                                }
                            }
                        }
                    };
                    uint4 nxt[4];
                    fetch4(j0, nxt);
                    for (uint32_t j = j0; j < j1; j += 4) {
                        uint32_t m[4][4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) { m[u][0] = nxt[u].x; m[u][1] = nxt[u].y; m[u][2] = nxt[u].z; m[u][3] = nxt[u].w; }
                        // carry-save add of the four masks into planes ones / twos, the carry ripples through the rest
#pragma unroll
                        for (int w = 0; w < 4; ++w) {
                            const uint32_t t1 = maj3(pl[0][w], m[0][w], m[1][w]);
                            pl[0][w] ^= m[0][w] ^ m[1][w];
                            const uint32_t t2 = maj3(pl[0][w], m[2][w], m[3][w]);
                            pl[0][w] ^= m[2][w] ^ m[3][w];
                            uint32_t carry = maj3(pl[1][w], t1, t2);
                            pl[1][w] ^= t1 ^ t2;
#pragma unroll
                            for (int q = 2; q < kCbPlanes; ++q) {
                                const uint32_t c = pl[q][w] & carry;
                                pl[q][w] ^= carry;
                                carry = c;
                            }
                        }
                        if (j + 4 < j1) fetch4(j + 4, nxt);
                    }
                }
                // sum over the lanes that own the same piece
                uint32_t res[RW][NPF];
                if constexpr (G == 32) {                                   // the whole warp shares one piece layout: nothing to fold
#pragma unroll
                    for (int q = 0; q < NPF; ++q)
#pragma unroll
                        for (int r = 0; r < 4; ++r) res[r][q] = pl[q][r];
                } else {
                    uint32_t f1[kCbPlanes + 1][2];
                    fold_words<4, kCbPlanes, 16>(pl, f1, lane);
                    if constexpr (G == 16) {
#pragma unroll
                        for (int q = 0; q < NPF; ++q) { res[0][q] = f1[q][0]; res[1][q] = f1[q][1]; }
                    } else {
                        uint32_t f2[kCbPlanes + 2][1];
                        fold_words<2, kCbPlanes + 1, 8>(f1, f2, lane);
                        if constexpr (G == 8) {
#pragma unroll
                            for (int q = 0; q < NPF; ++q) res[0][q] = f2[q][0];
                        } else {
                            uint32_t f3[kCbPlanes + 2];
#pragma unroll
                            for (int q = 0; q < kCbPlanes + 2; ++q) f3[q] = f2[q][0];
                            fold_bits<32, kCbPlanes + 2, 4>(f3, res[0], lane);
                        }
                    }
                }
                // acc += res (bit-sliced ripple add)
#pragma unroll
                for (int r = 0; r < RW; ++r) {
                    uint32_t carry = 0;
#pragma unroll
                    for (int q = 0; q < NPA; ++q) {
                        const uint32_t v = q < NPF ? res[r][q] : 0u;
                        const uint32_t sum = acc[r][q] ^ v ^ carry;
                        carry = maj3(acc[r][q], v, carry);
                        acc[r][q] = sum;
                    }
                }
            }
        }

        // ---- epilogue: the other strand's planes of my bins, then threshold / max ----------------------------------
        uint64_t best[kMaxLut];
#pragma unroll
        for (int t = 0; t < kMaxLut; ++t) best[t] = 0;
#pragma unroll
        for (int r = 0; r < RW; ++r) {
            uint32_t oth[NPA];
#pragma unroll
            for (int q = 0; q < NPA; ++q) oth[q] = __shfl_xor_sync(kFull, acc[r][q], G / 2);
            uint32_t valid = 0;
            if (bin0[r] < nbl) valid = (nbl - bin0[r] >= (uint64_t)B) ? ((B == 32) ? ~0u : ((1u << B) - 1u)) : ((1u << (nbl - bin0[r])) - 1u);
            if (a.counts_fwd || a.counts_rev) {
                uint16_t *dst = rev ? a.counts_rev : a.counts_fwd;
                if (dst)
                    for (int b = 0; b < B; ++b)
                        if ((valid >> b) & 1u) dst[read * nbl + bin0[r] + b] = (uint16_t)bs_get<NPA>(acc[r], b);
            }
#pragma unroll
            for (int t = 0; t < kMaxLut; ++t) {
                if (t < (int)a.n_lut && flag == 0) {
                    const uint32_t thr = (uint32_t)__ldg(a.lut + (size_t)t * kLutSize + len);
                    const uint32_t pass = (bs_ge<NPA>(acc[r], thr) | bs_ge<NPA>(oth, thr)) & valid;
                    if (pass) {
                        uint32_t s1 = pass, s2 = pass;
                        const uint32_t m1 = bs_max<NPA>(acc[r], s1), m2 = bs_max<NPA>(oth, s2);
                        const uint32_t mx = max(m1, m2);
                        const uint32_t at = (m1 == mx ? s1 : 0u) | (m2 == mx ? s2 : 0u);
                        const uint64_t key = pack_key(mx, (uint32_t)(a.fv.bin_begin + bin0[r] + (__ffs((int)at) - 1)));
                        best[t] = key > best[t] ? key : best[t];
                    }
                }
            }
        }
#pragma unroll
        for (int t = 0; t < kMaxLut; ++t) {
            if (t < (int)a.n_lut) {
                const uint64_t bk = warp_max_u64(best[t]);
                if (lane == 0) {
                    uint64_t *dst = a.keys + (size_t)t * a.n_reads + read;
                    if (a.keys_shared) { if (bk) key_max(dst, bk, 1); }
                    else *dst = bk;
                }
            }
        }
    }
}

template <int G, int NPA>
void launch_count_bs(const CountArgs &a, const uint64_t *table, int sm_count, cudaStream_t st)
{
    static int occ = 0;
    if (occ == 0) {
        int o = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, count_ctable_bs_kernel<G, NPA>, kCtWarps * 32, 0);
        occ = o > 0 ? o : 1;
    }
    const uint64_t blocks_needed = (a.n_reads + kCtWarps - 1) / kCtWarps;
    const uint64_t max_x = (uint64_t)sm_count * occ * grid_waves();
    const uint32_t gx = (uint32_t)(blocks_needed < max_x ? blocks_needed : max_x);
    count_ctable_bs_kernel<G, NPA><<<gx ? gx : 1, kCtWarps * 32, 0, st>>>(a, reinterpret_cast<const uint4 *>(table));
}

template <int G>
void launch_count_bs_np(const CountArgs &a, const uint64_t *table, uint32_t max_read_len, int sm_count, cudaStream_t st)
{
    const bool short_reads = max_read_len != 0 && (max_read_len < a.fv.hp.k || max_read_len - a.fv.hp.k + 1 <= (uint32_t)CbGeom<G>::SHORT_POS);
    if (short_reads) launch_count_bs<G, CbGeom<G>::NPS>(a, table, sm_count, st);
    else launch_count_bs<G, 16>(a, table, sm_count, st);
}

template <int G, int U>
void launch_count_g(const CountArgs &a, const uint64_t *table, int sm_count, cudaStream_t st)
{
    const size_t smem = (size_t)kCtWarps * (2 * 64 * G * 4 + kDigBytes);
    static int occ = 0;
    if (occ == 0) {
        cudaFuncSetAttribute(count_ctable_kernel<G, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int o = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, count_ctable_kernel<G, U>, kCtWarps * 32, smem);
        occ = o > 0 ? o : 1;
    }
    const uint64_t blocks_needed = (a.n_reads + kCtWarps - 1) / kCtWarps;
    const uint64_t max_x = (uint64_t)sm_count * occ * grid_waves();
    const uint32_t gx = (uint32_t)(blocks_needed < max_x ? blocks_needed : max_x);
    count_ctable_kernel<G, U><<<gx ? gx : 1, kCtWarps * 32, smem, st>>>(a, reinterpret_cast<const uint4 *>(table));
}

int ct_inflight()
{
    const char *e = std::getenv("RB_CTABLE_U");                           // measurements, tests
    const int v = e ? std::atoi(e) : 0;
    return (v == 1 || v == 2 || v == 8) ? v : 4;
}

template <int G>
void launch_count_gu(const CountArgs &a, const uint64_t *table, int sm_count, cudaStream_t st)
{
    switch (ct_inflight()) {
    case 1: launch_count_g<G, 1>(a, table, sm_count, st); break;
    case 2: launch_count_g<G, 2>(a, table, sm_count, st); break;
    case 8: launch_count_g<G, 8>(a, table, sm_count, st); break;
    default: launch_count_g<G, 4>(a, table, sm_count, st); break;
    }
}

}  // namespace

// lanes per entry (= padded row words) for a row of `stride` words; 0: this layout does not apply
int ctable_lanes(uint64_t stride)
{
    return stride < 3 ? 0 : stride <= 4 ? 4 : stride <= 8 ? 8 : stride <= 16 ? 16 : stride <= 32 ? 32 : 0;
}

int launch_ctable_build(const FilterView &fv, uint64_t *table, uint64_t n_kmers, int sm_count, cudaStream_t st)
{
    const int G = ctable_lanes(fv.stride);
    const uint64_t blocks = (n_kmers * (uint64_t)G + 255) / 256, cap = (uint64_t)sm_count * 32;
    const uint32_t gx = (uint32_t)(blocks < cap ? blocks : cap);
    uint4 *t = reinterpret_cast<uint4 *>(table);
    switch (G) {
    case 4: ctable_build_kernel<4><<<gx, 256, 0, st>>>(fv, t, n_kmers); break;
    case 8: ctable_build_kernel<8><<<gx, 256, 0, st>>>(fv, t, n_kmers); break;
    case 16: ctable_build_kernel<16><<<gx, 256, 0, st>>>(fv, t, n_kmers); break;
    case 32: ctable_build_kernel<32><<<gx, 256, 0, st>>>(fv, t, n_kmers); break;
    default: return -1;
    }
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// variant 0: bit-sliced register counters; 1: shared-memory atomic counters (count kernel selector 4, RB_CTABLE_ATOMIC=1)
int launch_count_ctable(const CountArgs &a, const uint64_t *table, uint32_t max_read_len, int variant, int sm_count, cudaStream_t st)
{
    if (a.n_reads == 0) return 0;
    if (a.n_lut == 0 || a.n_lut > (uint32_t)kMaxLut) return -1;
    if (const char *e = std::getenv("RB_CTABLE_ATOMIC")) if (e[0] == '1') variant = 1;
    const int G = ctable_lanes(a.fv.stride);
    if (variant == 0) {
        switch (G) {
        case 4: launch_count_bs_np<4>(a, table, max_read_len, sm_count, st); break;
        case 8: launch_count_bs_np<8>(a, table, max_read_len, sm_count, st); break;
        case 16: launch_count_bs_np<16>(a, table, max_read_len, sm_count, st); break;
        case 32: launch_count_bs_np<32>(a, table, max_read_len, sm_count, st); break;
        default: return -1;
        }
        return cudaGetLastError() == cudaSuccess ? 1 : -1;
    }
    switch (G) {
    case 4: launch_count_gu<4>(a, table, sm_count, st); break;
    case 8: launch_count_gu<8>(a, table, sm_count, st); break;
    case 16: launch_count_gu<16>(a, table, sm_count, st); break;
    case 32: launch_count_gu<32>(a, table, sm_count, st); break;
    default: return -1;
    }
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace rb
