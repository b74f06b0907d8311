// ibf_table.cu -- direct k-mer table for narrow filters (<= 256 bins) on B200.
//
// What seqan::count does per k-mer and strand (src/IBF/IBFClassify.cpp:149-150 -> SURVEY.md
// Appendix A.6) is a pure function of the k-mer: AND of the h rows its hashes select.  On B200
// every random probe of a row narrower than 128 B costs one whole 128-byte HBM line (measured:
// ncu dram__bytes_read ~= 114 B per 16-byte probe, profiles/r1_a_*), so a 250-base chunk pays
// 2 strands x 238 k-mers x 3 probes = 1428 lines.  With 180 GB of HBM the function can simply be
// tabulated once per filter for every ACGT k-mer:
//
//     table[x] = { AND_i row(h_i(x)) , AND_i row(h_i(revcomp(x))) }      x in [0, 4^k)
//
// (2*W words; 2.1 GB for k=13 and 100 bins).  Classifying a chunk then costs ONE line per k-mer
// position for both strands -- 238 instead of 1428 -- and no hashing at all; results are the
// same bits the three probes would have produced.  Windows containing a non-ACGT base (Dna5
// rank 4) are not in the table and take the hashing path of the original filter, so the output
// stays bit-exact for every input.
#include "ibf_bitslice.cuh"

namespace rb {

// ------------------------------------------------------------------------------------------
// table build: one thread per k-mer value
// ------------------------------------------------------------------------------------------
// Entry y (a window of k+S-1 bases, 2 bits each) holds S consecutive k-mers: [t][fwd W words][rev W words].
template <int WT, int S>
__global__ void __launch_bounds__(256) table_build_kernel(const FilterView fv, uint64_t *__restrict__ table,
                                                          const uint64_t n_entries)
{
    const HashParams &hp = fv.hp;
    const uint32_t k = hp.k;
    const uint32_t wl = k + S - 1;                                   // window length in bases
    for (uint64_t y = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; y < n_entries;
         y += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t *e = table + y * (2 * WT * S);
#pragma unroll
        for (int t = 0; t < S; ++t) {
            uint64_t Hf = 0, Hr = 0, pw = 1;
            for (uint32_t j = 0; j < k; ++j) {
                uint32_t d = (uint32_t)(y >> (2 * (wl - 1 - (t + j)))) & 3u;   // base t+j of the window
                Hf = Hf * 5 + d;
                Hr += (uint64_t)(3u - d) * pw;
                pw *= 5;
            }
            uint64_t mf[WT], mr[WT];
#pragma unroll
            for (int w = 0; w < WT; ++w) { mf[w] = ~0ULL; mr[w] = ~0ULL; }
            for (uint32_t i = 0; i < hp.n_hash; ++i) {
                const uint64_t *pf = fv.words + hash_row(Hf, hp.pre[i], hp.n_blocks, hp.magic) * fv.stride;
                const uint64_t *pr = fv.words + hash_row(Hr, hp.pre[i], hp.n_blocks, hp.magic) * fv.stride;
#pragma unroll
                for (int w = 0; w < WT; ++w) { mf[w] &= __ldg(pf + w); mr[w] &= __ldg(pr + w); }
            }
#pragma unroll
            for (int w = 0; w < WT; ++w) { e[t * 2 * WT + w] = mf[w]; e[t * 2 * WT + WT + w] = mr[w]; }
        }
    }
}

// ------------------------------------------------------------------------------------------
// lookup kernel: warp per read, lane per run of consecutive k-mer positions
// ------------------------------------------------------------------------------------------
// Hashing path for windows that contain a non-ACGT base (rare): direct evaluation from the digits.
template <int WT>
__device__ inline void probe_hashed(const FilterView &fv, const uint8_t *dig, uint32_t j, uint64_t (&mf)[WT],
                                    uint64_t (&mr)[WT])
{
    const HashParams &hp = fv.hp;
    uint64_t Hf = 0, Hr = 0, pw = 1;
    for (uint32_t u = 0; u < hp.k; ++u) {
        uint32_t d = dig[j + u];
        Hf = Hf * 5 + d;
        Hr += comp5(d) * pw;
        pw *= 5;
    }
#pragma unroll
    for (int w = 0; w < WT; ++w) { mf[w] = ~0ULL; mr[w] = ~0ULL; }
#pragma unroll 1
    for (uint32_t i = 0; i < hp.n_hash; ++i) {
        const uint64_t *pf = fv.words + hash_row(Hf, hp.pre[i], hp.n_blocks, hp.magic) * fv.stride;
        const uint64_t *pr = fv.words + hash_row(Hr, hp.pre[i], hp.n_blocks, hp.magic) * fv.stride;
#pragma unroll
        for (int w = 0; w < WT; ++w) { mf[w] &= __ldg(pf + w); mr[w] &= __ldg(pr + w); }
    }
}

// one table entry = 2*WT words, 16*WT bytes, 16-byte aligned
template <int WT>
__device__ __forceinline__ void load_entry(const uint64_t *__restrict__ e, uint64_t (&mf)[WT], uint64_t (&mr)[WT])
{
    uint64_t v[2 * WT];
#pragma unroll
    for (int i = 0; i < WT; ++i) {
        ulonglong2 t = __ldg(reinterpret_cast<const ulonglong2 *>(e) + i);
        v[2 * i] = t.x;
        v[2 * i + 1] = t.y;
    }
#pragma unroll
    for (int w = 0; w < WT; ++w) { mf[w] = v[w]; mr[w] = v[WT + w]; }
}

template <int WT, int U>
__global__ void __launch_bounds__(kTileWarps * 32)
count_table_kernel(const CountArgs a, const uint64_t *__restrict__ table)
{
    __shared__ __align__(16) uint8_t s_dig[kTileWarps][kDigBytes];
    __shared__ uint32_t s_cnt[kTileWarps][2][64 * WT];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t total_warps = (uint64_t)gridDim.x * kTileWarps;
    const uint32_t k = a.fv.hp.k;
    const uint64_t kmask = (k >= 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
    uint8_t *dig = s_dig[warp];
    uint32_t *cntF = s_cnt[warp][0], *cntR = s_cnt[warp][1];

    for (int b = lane; b < 64 * WT; b += 32) { cntF[b] = 0; cntR[b] = 0; }
    __syncwarp();

    for (uint64_t read = (uint64_t)blockIdx.x * kTileWarps + warp; read < a.n_reads; read += total_warps) {
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
                const uint32_t seg = (cn + 31) >> 5;
                const uint32_t j0 = lane * seg;
                const uint32_t j1 = min(j0 + seg, cn);
                if (j0 < j1) {
                    uint64_t x = 0;          // 2-bit packed window
                    uint32_t nbad = 0;       // non-ACGT bases inside the window
                    for (uint32_t u = 0; u < k; ++u) {
                        uint32_t d = dig[j0 + u];
                        x = (x << 2) | (d & 3u);
                        nbad += d >> 2;
                    }
                    x &= kmask;
                    for (uint32_t j = j0; j < j1; j += U) {
                        uint64_t mf[U][WT], mr[U][WT];
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            if (j + u < j1) {
                                if (nbad == 0) load_entry<WT>(table + x * (2 * WT), mf[u], mr[u]);
                                else probe_hashed<WT>(a.fv, dig, j + u, mf[u], mr[u]);
                                if (j + u + 1 < j1) {
                                    uint32_t dout = dig[j + u], din = dig[j + u + k];
                                    x = ((x << 2) | (din & 3u)) & kmask;
                                    nbad += (din >> 2) - (dout >> 2);
                                }
                            }
                        }
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            if (j + u < j1) {
                                count_bits<WT>(mf[u], cntF);
                                count_bits<WT>(mr[u], cntR);
                            }
                        }
                    }
                }
            }
        }
        __syncwarp();
        tile_epilogue<WT>(a, read, len, flag, 0, cntF, cntR, lane, 0);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// lookup kernel with bit-sliced counting (rows of <= 2 words)
// ------------------------------------------------------------------------------------------
// The per-bin counters never leave registers.  Each lane adds the masks of its <= 15 k-mer positions
// into 4 bit planes per 32-bit mask word with carry-save adders (9 logic ops per 4 positions and word),
// so the cost does not depend on how many bits are set.  The 32 lanes are then summed by an
// exchange-and-halve butterfly of bit-sliced full adders (shuffle distance 16, 8, 4, 2, 1): every
// level a lane keeps one half of its span and receives its partner's copy of that half, so after five
// levels lane l holds the 9-plane counts of NWP consecutive mask bits.  Words are ordered
// q = 2*word32 + strand, which puts the forward and reverse counts of the same bins in lanes l and
// l ^ (32/NWP): one more exchange and each lane evaluates select_matches / max_matches for its bins.
constexpr int kBsSeg = 15;                  // positions per lane per chunk (4 planes hold 0..15)
constexpr int kBsChunk = 32 * kBsSeg;       // 480 positions per warp chunk
constexpr int kBsDig = kBsChunk + 32;
// WT: row words (1 or 2).  NPA: planes of the per-read accumulator (9 when every read is a single
// chunk, else 16).  S: consecutive k-mers per table entry (window table).
template <int WT, int NPA, int S>
__global__ void __launch_bounds__(kTileWarps * 32)
count_table_bs_kernel(const CountArgs a, const uint64_t *__restrict__ table)
{
    constexpr int NWP = 4 * WT;             // 32-bit mask words of both strands
    constexpr int B = NWP;                  // mask bits owned by a lane after the fold
    constexpr int LPW = 32 / B;             // lanes per mask word
    __shared__ __align__(16) uint8_t s_dig[kTileWarps][kBsDig];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t total_warps = (uint64_t)gridDim.x * kTileWarps;
    const uint32_t k = a.fv.hp.k;
    const uint32_t wl = k + S - 1;                                   // bases per table window
    const uint64_t wmask = (wl >= 32) ? ~0ULL : ((1ULL << (2 * wl)) - 1);
    uint8_t *dig = s_dig[warp];
    // bins this lane reports: strand = q & 1, 32-bit word ww = q >> 1 of that strand
    const int q = lane / LPW;
    const int strand = q & 1;
    const uint32_t bin0 = (uint32_t)(q >> 1) * 32u + (uint32_t)B * (uint32_t)(lane % LPW);

    for (uint64_t read = (uint64_t)blockIdx.x * kTileWarps + warp; read < a.n_reads; read += total_warps) {
        const uint64_t off = a.read_off[read];
        const uint64_t len = a.read_off[read + 1] - off;
        uint32_t flag = read_flag_of(len, k);
        // a read longer than the caller's max_read_len promised does not fit the NPA-bit accumulator: flag 3, not classified
        if (flag == 0 && NPA < 16 && len - k + 1 > (1u << NPA) - 1u) flag = 3;
        if (lane == 0 && a.read_flag) a.read_flag[read] = (uint8_t)flag;

        uint32_t acc[NPA];
#pragma unroll
        for (int p = 0; p < NPA; ++p) acc[p] = 0;

        if (flag == 0) {
            const uint32_t npos = (uint32_t)len - k + 1;
            for (uint32_t cs = 0; cs < npos; cs += kBsChunk) {
                const uint32_t cn = min((uint32_t)kBsChunk, npos - cs);
                __syncwarp();
                for (uint32_t i = lane; i < cn + k - 1; i += 32) dig[i] = (uint8_t)dna5(a.bases[off + cs + i]);
                __syncwarp();
                const uint32_t seg = (cn + 31) >> 5;
                const uint32_t j0 = lane * seg;
                const uint32_t j1 = min(j0 + seg, cn);
                uint32_t pl[4][NWP];
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int w = 0; w < NWP; ++w) pl[p][w] = 0;
                if (j0 < j1) {
                    // window of wl bases starting at j; bases past the end of the chunk read as 'A' (their
                    // k-mers are never consumed) -- the digit buffer is zero-padded below
                    const uint32_t nd = cn + k - 1;                  // valid digits in the chunk buffer
                    uint64_t x = 0;
                    uint32_t nbad = 0;
                    for (uint32_t u = 0; u < wl; ++u) {
                        uint32_t d = (j0 + u < nd) ? dig[j0 + u] : 0u;
                        x = (x << 2) | (d & 3u);
                        nbad += d >> 2;
                    }
                    x &= wmask;
                    for (uint32_t j = j0; j < j1; j += 4) {
                        uint64_t mf[4][WT], mr[4][WT];
#pragma unroll
                        for (int g = 0; g < 4 / S; ++g) {
                            const uint32_t jg = j + g * S;
#pragma unroll
                            for (int t = 0; t < S; ++t)
#pragma unroll
                                for (int w = 0; w < WT; ++w) { mf[g * S + t][w] = 0; mr[g * S + t][w] = 0; }
                            if (jg < j1) {
                                if (nbad == 0) {
                                    const uint64_t *e = table + x * (2 * WT * S);
#pragma unroll
                                    for (int t = 0; t < S; ++t) {
                                        uint64_t f[WT], r[WT];
                                        load_entry<WT>(e + t * 2 * WT, f, r);
                                        if (jg + t < j1) {
#pragma unroll
                                            for (int w = 0; w < WT; ++w) { mf[g * S + t][w] = f[w]; mr[g * S + t][w] = r[w]; }
                                        }
                                    }
                                } else {
#pragma unroll
                                    for (int t = 0; t < S; ++t)
                                        if (jg + t < j1) probe_hashed<WT>(a.fv, dig, jg + t, mf[g * S + t], mr[g * S + t]);
                                }
                                if (jg + S < j1) {                   // slide the window by S bases
#pragma unroll
                                    for (int t = 0; t < S; ++t) {
                                        const uint32_t dout = dig[jg + t];
                                        const uint32_t din = (jg + t + wl < nd) ? dig[jg + t + wl] : 0u;
                                        x = ((x << 2) | (din & 3u)) & wmask;
                                        nbad += (din >> 2) - (dout >> 2);
                                    }
                                }
                            }
                        }
                        // carry-save add of the four masks into planes ones/twos/fours/eights
#pragma unroll
                        for (int w = 0; w < NWP; ++w) {
                            uint32_t m[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const uint64_t v = (w & 1) ? mr[u][w >> 2] : mf[u][w >> 2];   // q = 2*word32 + strand
                                m[u] = ((w >> 1) & 1) ? (uint32_t)(v >> 32) : (uint32_t)v;
                            }
                            uint32_t t1 = maj3(pl[0][w], m[0], m[1]);
                            pl[0][w] ^= m[0] ^ m[1];
                            uint32_t t2 = maj3(pl[0][w], m[2], m[3]);
                            pl[0][w] ^= m[2] ^ m[3];
                            uint32_t f4 = maj3(pl[1][w], t1, t2);
                            pl[1][w] ^= t1 ^ t2;
                            uint32_t c8 = pl[2][w] & f4;
                            pl[2][w] ^= f4;
                            pl[3][w] ^= c8;
                        }
                    }
                }
                uint32_t res[9];
                warp_fold<NWP>(pl, res, lane);
                // acc += res (bit-sliced ripple add, NPA planes)
                uint32_t carry = 0;
#pragma unroll
                for (int p = 0; p < NPA; ++p) {
                    const uint32_t r = p < 9 ? res[p] : 0u;
                    const uint32_t s = acc[p] ^ r ^ carry;
                    carry = maj3(acc[p], r, carry);
                    acc[p] = s;
                }
            }
        }

        // ---- epilogue: bring the other strand's planes of my bins, then threshold / max -----------------
        uint32_t oth[NPA];
#pragma unroll
        for (int p = 0; p < NPA; ++p) oth[p] = __shfl_xor_sync(kFull, acc[p], LPW);
        const uint64_t nbl = a.fv.n_bins_local;
        uint32_t valid = 0;
        if (bin0 < nbl) valid = (nbl - bin0 >= (uint64_t)B) ? ((B == 32) ? ~0u : ((1u << B) - 1u)) : ((1u << (nbl - bin0)) - 1u);
        if (a.counts_fwd || a.counts_rev) {
            uint16_t *dst = strand == 0 ? a.counts_fwd : a.counts_rev;
            if (dst)
                for (int b = 0; b < B; ++b)
                    if ((valid >> b) & 1u) dst[read * nbl + bin0 + b] = (uint16_t)bs_get<NPA>(acc, b);
        }
        uint64_t best[kMaxLut];
#pragma unroll
        for (int t = 0; t < kMaxLut; ++t) {
            best[t] = 0;
            if (t < (int)a.n_lut && flag == 0) {
                const uint32_t thr = (uint32_t)__ldg(a.lut + (size_t)t * kLutSize + len);
                const uint32_t pass = (bs_ge<NPA>(acc, thr) | bs_ge<NPA>(oth, thr)) & valid;
                if (pass) {
                    // max_matches over the passing bins without a per-bin loop: bit-sliced max of each strand,
                    // then the lowest bin among those attaining the larger one
                    uint32_t s1 = pass, s2 = pass;
                    const uint32_t m1 = bs_max<NPA>(acc, s1), m2 = bs_max<NPA>(oth, s2);
                    const uint32_t m = max(m1, m2);
                    const uint32_t at = (m1 == m ? s1 : 0u) | (m2 == m ? s2 : 0u);
                    best[t] = pack_key(m, (uint32_t)(a.fv.bin_begin + bin0 + (__ffs((int)at) - 1)));
                }
            }
        }
#pragma unroll
        for (int t = 0; t < kMaxLut; ++t) {
            if (t < (int)a.n_lut) {
                const uint64_t bk = warp_max_u64(best[t]);
                if (lane == 0) a.keys[(size_t)t * a.n_reads + read] = bk;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
template <int WT, int S>
static void launch_build_wt(const FilterView &fv, uint64_t *table, uint64_t n_entries, int sm_count, cudaStream_t st)
{
    uint64_t blocks = (n_entries + 255) / 256;
    uint64_t cap = (uint64_t)sm_count * 32;
    table_build_kernel<WT, S><<<(uint32_t)(blocks < cap ? blocks : cap), 256, 0, st>>>(fv, table, n_entries);
}

// one k-mer per entry (span 1); wider windows live in ibf_wtable.cu
int launch_table_build(const FilterView &fv, uint64_t *table, uint64_t n_entries, int span, int sm_count, cudaStream_t st)
{
    if (span != 1) return -1;
    switch (fv.stride) {
    case 1: launch_build_wt<1, 1>(fv, table, n_entries, sm_count, st); break;
    case 2: launch_build_wt<2, 1>(fv, table, n_entries, sm_count, st); break;
    case 3: launch_build_wt<3, 1>(fv, table, n_entries, sm_count, st); break;
    case 4: launch_build_wt<4, 1>(fv, table, n_entries, sm_count, st); break;
    default: return -1;
    }
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

template <int WT, int U>
static void launch_table_wt(const CountArgs &a, const uint64_t *table, int sm_count, cudaStream_t st)
{
    static int occ = 0;
    if (occ == 0) {
        int o = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, count_table_kernel<WT, U>, kTileWarps * 32, 0);
        occ = o > 0 ? o : 1;
    }
    uint64_t blocks_needed = (a.n_reads + kTileWarps - 1) / kTileWarps;
    uint64_t max_x = (uint64_t)sm_count * occ * grid_waves();
    uint32_t gx = (uint32_t)(blocks_needed < max_x ? blocks_needed : max_x);
    count_table_kernel<WT, U><<<gx ? gx : 1, kTileWarps * 32, 0, st>>>(a, table);
}

template <int WT, int NPA, int S>
static void launch_table_bs(const CountArgs &a, const uint64_t *table, int sm_count, cudaStream_t st)
{
    static int occ = 0;
    if (occ == 0) {
        int o = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, count_table_bs_kernel<WT, NPA, S>, kTileWarps * 32, 0);
        occ = o > 0 ? o : 1;
    }
    uint64_t blocks_needed = (a.n_reads + kTileWarps - 1) / kTileWarps;
    uint64_t max_x = (uint64_t)sm_count * occ * grid_waves();
    uint32_t gx = (uint32_t)(blocks_needed < max_x ? blocks_needed : max_x);
    count_table_bs_kernel<WT, NPA, S><<<gx ? gx : 1, kTileWarps * 32, 0, st>>>(a, table);
}

template <int WT, int S>
static void launch_table_bs_np(const CountArgs &a, const uint64_t *table, bool single_chunk, int sm_count, cudaStream_t st)
{
    if (single_chunk) launch_table_bs<WT, 9, S>(a, table, sm_count, st);
    else launch_table_bs<WT, 16, S>(a, table, sm_count, st);
}

// entries a lane keeps in flight in the atomic-counter kernel for rows of 3-4 words (RB_TABLE_U: measurements)
static int table_u()
{
    static const int u = [] { const char *e = std::getenv("RB_TABLE_U"); const int v = e ? std::atoi(e) : 0; return (v == 2 || v == 4) ? v : 1; }();
    return u;
}

// variant 0: bit-sliced register counters (rows <= 2 words); 1: shared-memory atomic counters.
int launch_count_table(const CountArgs &a, const uint64_t *table, int span, uint32_t max_read_len, int variant,
                       int sm_count, cudaStream_t st)
{
    if (a.n_reads == 0) return 0;
    if (a.n_lut == 0 || a.n_lut > (uint32_t)kMaxLut || span != 1) return -1;
    if (variant == 0 && a.fv.stride <= 2) {
        const bool single_chunk = max_read_len != 0 && max_read_len < a.fv.hp.k + (uint32_t)kBsChunk;
        if (a.fv.stride == 1) launch_table_bs_np<1, 1>(a, table, single_chunk, sm_count, st);
        else launch_table_bs_np<2, 1>(a, table, single_chunk, sm_count, st);
        return cudaGetLastError() == cudaSuccess ? 1 : -1;
    }
    switch (a.fv.stride) {
    case 1: launch_table_wt<1, 2>(a, table, sm_count, st); break;
    case 2: launch_table_wt<2, 2>(a, table, sm_count, st); break;
    case 3: if (table_u() == 4) launch_table_wt<3, 4>(a, table, sm_count, st); else if (table_u() == 2) launch_table_wt<3, 2>(a, table, sm_count, st); else launch_table_wt<3, 1>(a, table, sm_count, st); break;
    case 4: if (table_u() == 4) launch_table_wt<4, 4>(a, table, sm_count, st); else if (table_u() == 2) launch_table_wt<4, 2>(a, table, sm_count, st); else launch_table_wt<4, 1>(a, table, sm_count, st); break;
    default: return -1;
    }
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace rb
