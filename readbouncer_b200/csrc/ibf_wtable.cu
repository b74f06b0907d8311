// ibf_wtable.cu -- window k-mer table with warp-cooperative entry loads (narrow filters, <= 128 bins).
//
// Measured on B200 (profiles/r1_d_gather_sweep3_coop.jsonl): an L2-missing random read costs the same
// ~1/43 G s whether it brings 16, 32, 64 or 128 contiguous bytes, PROVIDED the bytes are requested by
// adjacent lanes of one load instruction.  The unit of cost of the classify path for narrow rows is
// therefore the REQUEST, and the way to go faster is to make one request serve several k-mer positions.
//
// What seqan::count does per k-mer and strand (src/IBF/IBFClassify.cpp:149-150) is
// a pure function of the k-mer, so it is tabulated per WINDOW of L = k+S-1 bases:
//
//     entry(window) = [ slot t = 0..S-1 : AND_i row(h_i(kmer_t)) , AND_i row(h_i(revcomp(kmer_t))) ]
//
// G adjacent lanes load one entry (slot t -> lane t, 16*WT bytes per lane, G*16*WT <= 128 bytes per
// entry) and a 250-base chunk costs ceil(238/S) requests instead of 238 (one-position table) or 1428
// (hashed probes).  When L is odd the table is CANONICAL: a window and its reverse complement hold the
// same 2*S masks (slot order and strands swapped), and exactly one of the two has a middle base in
// {A, C}, so only those 4^L/2 windows are stored -- this is what lets S = 3 (L = 15 for k = 13) fit in
// HBM (68.7 GB for 100 bins).  Windows containing a non-ACGT base are not tabulated; their positions
// take the hashing path on the original bit matrix, so every output bit equals the reference's.
//
// Base codes inside the table index are (ascii >> 1) & 3 = A0 C1 T2 G3 (complement = flip the high bit);
// the index is "planar": the L low bits of the codes, then the L high bits (minus the middle one when
// canonical).  The two bit planes of a read are made with warp ballots, 32 bases per vote.
#include "ibf_bitslice.cuh"

namespace rb {

namespace {

int g_wtable_variant = 0;     // 0 auto, 1 force the warp-per-read kernel (tests, measurements)

struct WGeom {
    uint32_t L;        // window length in bases
    uint32_t mid;      // index of the middle base (canonical tables)
    uint32_t maskL;    // (1 << L) - 1
};

__host__ __device__ inline WGeom make_wgeom(uint32_t k, int span)
{
    WGeom g;
    g.L = k + (uint32_t)span - 1;
    g.mid = (g.L - 1) / 2;
    g.maskL = (g.L >= 32) ? ~0u : ((1u << g.L) - 1u);
    return g;
}

// rank (A0 C1 G2 T3) of a table code (A0 C1 T2 G3)
__device__ __forceinline__ uint32_t rank_of_code(uint32_t c) { return c ^ (c >> 1); }

// ------------------------------------------------------------------------------------------
// table build: one thread per (entry, slot)
// ------------------------------------------------------------------------------------------
template <int WT, int S, int G, bool CANON>
__global__ void __launch_bounds__(256) wtable_build_kernel(const FilterView fv, uint64_t *__restrict__ table,
                                                           const uint64_t n_entries)
{
    const HashParams &hp = fv.hp;
    const uint32_t k = hp.k;
    const WGeom wg = make_wgeom(k, S);
    const uint64_t n_threads = n_entries * G;
    for (uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; tid < n_threads;
         tid += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t idx = tid / G;
        const uint32_t t = (uint32_t)(tid % G);
        uint64_t *e = table + tid * (2 * WT);
        uint64_t mf[WT], mr[WT];
        if (t >= (uint32_t)S) {
#pragma unroll
            for (int w = 0; w < WT; ++w) { mf[w] = 0; mr[w] = 0; }
        } else {
            const uint32_t lo = (uint32_t)idx & wg.maskL;
            const uint32_t h = (uint32_t)(idx >> wg.L);
            const uint32_t hi = CANON ? ((h & ((1u << wg.mid) - 1u)) | ((h >> wg.mid) << (wg.mid + 1))) : h;
            uint64_t Hf = 0, Hr = 0, pw = 1;
            for (uint32_t j = 0; j < k; ++j) {
                const uint32_t c = (((hi >> (t + j)) & 1u) << 1) | ((lo >> (t + j)) & 1u);
                const uint32_t d = rank_of_code(c);
                Hf = Hf * 5 + d;
                Hr += (uint64_t)(3u - d) * pw;
                pw *= 5;
            }
#pragma unroll
            for (int w = 0; w < WT; ++w) { mf[w] = ~0ULL; mr[w] = ~0ULL; }
            for (uint32_t i = 0; i < hp.n_hash; ++i) {
                const uint64_t *pf = fv.words + hash_row(Hf, hp.pre[i], hp.n_blocks, hp.magic) * fv.stride;
                const uint64_t *pr = fv.words + hash_row(Hr, hp.pre[i], hp.n_blocks, hp.magic) * fv.stride;
#pragma unroll
                for (int w = 0; w < WT; ++w) { mf[w] &= __ldg(pf + w); mr[w] &= __ldg(pr + w); }
            }
        }
        if constexpr (WT == 2) {
            reinterpret_cast<ulonglong2 *>(e)[0] = make_ulonglong2(mf[0], mf[1]);
            reinterpret_cast<ulonglong2 *>(e)[1] = make_ulonglong2(mr[0], mr[1]);
        } else {
            reinterpret_cast<ulonglong2 *>(e)[0] = make_ulonglong2(mf[0], mr[0]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// lookup kernel
// ------------------------------------------------------------------------------------------
// one slot of an entry: 2*WT words = [fwd WT][rev WT], as 4*WT 32-bit words
template <int WT>
__device__ __forceinline__ void load_slot(const uint64_t *p, uint32_t (&v)[4 * WT])
{
    if constexpr (WT == 2) {
        asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "l"(p));
    } else {
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
                     : "l"(p));
    }
}

// hashing path for a position whose window holds a non-ACGT base: direct evaluation from the ASCII bases
template <int WT>
struct SlotWords { uint32_t v[4 * WT]; };

template <int WT>
__device__ __noinline__ SlotWords<WT> slot_hashed(const FilterView &fv, const uint8_t *__restrict__ bases)
{
    SlotWords<WT> out;
    uint32_t (&v)[4 * WT] = out.v;
    const HashParams &hp = fv.hp;
    uint64_t Hf = 0, Hr = 0, pw = 1;
    for (uint32_t u = 0; u < hp.k; ++u) {
        const uint32_t d = dna5(bases[u]);
        Hf = Hf * 5 + d;
        Hr += comp5(d) * pw;
        pw *= 5;
    }
    uint64_t mf[WT], mr[WT];
#pragma unroll
    for (int w = 0; w < WT; ++w) { mf[w] = ~0ULL; mr[w] = ~0ULL; }
#pragma unroll 1
    for (uint32_t i = 0; i < hp.n_hash; ++i) {
        const uint64_t *pf = fv.words + hash_row(Hf, hp.pre[i], hp.n_blocks, hp.magic) * fv.stride;
        const uint64_t *pr = fv.words + hash_row(Hr, hp.pre[i], hp.n_blocks, hp.magic) * fv.stride;
#pragma unroll
        for (int w = 0; w < WT; ++w) { mf[w] &= __ldg(pf + w); mr[w] &= __ldg(pr + w); }
    }
#pragma unroll
    for (int w = 0; w < WT; ++w) {
        v[2 * w] = (uint32_t)mf[w];
        v[2 * w + 1] = (uint32_t)(mf[w] >> 32);
        v[2 * WT + 2 * w] = (uint32_t)mr[w];
        v[2 * WT + 2 * w + 1] = (uint32_t)(mr[w] >> 32);
    }
    return out;
}

constexpr int kWtBallots = 9;                     // 288 bases per piece

template <int S, int G>
struct WtShape {
    static constexpr int EPW = 32 / G;            // entries per warp load
    static constexpr int PPI = EPW * S;           // k-mer positions per load iteration
    static constexpr int IT = (32 * kWtBallots - 15) / PPI > 10 ? 10 : (32 * kWtBallots - 15) / PPI;
    static constexpr int PIECE = PPI * IT;        // positions per piece
    // the 64-bit funnel (two ballot words) must cover every window (<= 16 bases) of every iteration
    static constexpr int max_start()
    {
        int m = 0;
        for (int it = 0; it < IT; ++it) m = ((PPI * it) & 31) > m ? ((PPI * it) & 31) : m;
        return m;
    }
    static_assert(max_start() + S * (EPW - 1) + 15 <= 63, "window leaves the funnel");
    static_assert(((PPI * (IT - 1)) >> 5) + 1 < kWtBallots, "ballot words");
    static_assert(IT <= 15, "4 planes hold 0..15");
};

// WT: row words (1 or 2).  S: k-mers per entry.  G: lanes per entry.  NPA: planes of the per-read accumulator
// (9 when every read is a single piece, else 16).
template <int WT, int S, int G, bool CANON, int NPA>
__global__ void __launch_bounds__(kTileWarps * 32, 2)
count_wtable_kernel(const CountArgs a, const uint64_t *__restrict__ table)
{
    using Sh = WtShape<S, G>;
    constexpr int NWP = 4 * WT;             // 32-bit mask words of both strands
    constexpr int B = NWP;                  // mask bits owned by a lane after the fold
    constexpr int LPW = 32 / B;             // lanes per mask word
    constexpr int IT = Sh::IT, PPI = Sh::PPI;
    __shared__ FilterView s_fv;             // for the out-of-line hashed path: one copy per CTA, no per-thread stack frame
    if (threadIdx.x == 0) s_fv = a.fv;
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t total_warps = (uint64_t)gridDim.x * kTileWarps;
    const uint32_t k = a.fv.hp.k;
    const WGeom wg = make_wgeom(k, S);
    const uint32_t g = (uint32_t)lane / G, t = (uint32_t)lane % G;
    // bins this lane reports after the fold: strand = q & 1, 32-bit word ww = q >> 1 of that strand
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
            for (uint32_t cs = 0; cs < npos; cs += Sh::PIECE) {
                const uint32_t cn = min((uint32_t)Sh::PIECE, npos - cs);      // positions of this piece
                const uint32_t avail = (uint32_t)len - cs;                    // bases from the piece start on
                const uint8_t *__restrict__ pb = a.bases + off + cs;
                // ---- the piece's bases as three bit planes (warp ballots) -----------------------------
                uint32_t lo[kWtBallots], hi[kWtBallots], bad[kWtBallots];
#pragma unroll
                for (int r = 0; r < kWtBallots; ++r) {
                    lo[r] = 0; hi[r] = 0; bad[r] = 0;
                    if (32u * r < avail && 32u * r < cn + wg.L - 1) {
                        const uint32_t i = 32u * r + lane;
                        uint32_t c = 'A';
                        if (i < avail) c = pb[i];
                        const uint32_t x = (c & 0xDFu) - 'A';                 // A 0, C 2, G 6, T 19, U 20
                        const bool ok = x < 21u && ((0x180045u >> x) & 1u);
                        lo[r] = __ballot_sync(kFull, (c >> 1) & 1u);
                        hi[r] = __ballot_sync(kFull, (c >> 2) & 1u);
                        bad[r] = __ballot_sync(kFull, !ok);
                    }
                }
                uint32_t pl[4][NWP];
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int w = 0; w < NWP; ++w) pl[p][w] = 0;

#pragma unroll
                for (int it0 = 0; it0 < IT; it0 += 4) {
                    uint32_t m[4][NWP];
                    bool flip[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int it = it0 + u;
                        flip[u] = false;
#pragma unroll
                        for (int w = 0; w < NWP; ++w) m[u][w] = 0;
                        if (it < IT && (uint32_t)(PPI * it) < cn) {
                            const int w0 = (PPI * it) >> 5;
                            const uint32_t sh = (uint32_t)((PPI * it) & 31) + (uint32_t)S * g;
                            const uint32_t p = (uint32_t)(PPI * it) + (uint32_t)S * g;     // first position of my entry
                            const uint32_t wlo = (uint32_t)((((uint64_t)lo[w0 + 1] << 32) | lo[w0]) >> sh) & wg.maskL;
                            uint32_t whi = (uint32_t)((((uint64_t)hi[w0 + 1] << 32) | hi[w0]) >> sh) & wg.maskL;
                            const uint32_t wbad = (uint32_t)((((uint64_t)bad[w0 + 1] << 32) | bad[w0]) >> sh) & wg.maskL;
                            uint32_t ilo = wlo;
                            bool fl = false;
                            if constexpr (CANON) {
                                fl = (whi >> wg.mid) & 1u;
                                if (fl) {
                                    ilo = __brev(wlo) >> (32 - wg.L);
                                    whi = __brev(~whi) >> (32 - wg.L);
                                }
                            }
                            const uint32_t pos = p + (fl ? (uint32_t)(S - 1) - t : t);   // the position my slot answers
                            if (t < (uint32_t)S && pos < cn) {
                                if (wbad == 0) {
                                    uint64_t idx;
                                    if constexpr (CANON)
                                        idx = (uint64_t)ilo | ((uint64_t)(whi & ((1u << wg.mid) - 1u)) << wg.L) |
                                              ((uint64_t)(whi >> (wg.mid + 1)) << (wg.L + wg.mid));
                                    else
                                        idx = (uint64_t)ilo | ((uint64_t)whi << wg.L);
                                    load_slot<WT>(table + (idx * G + t) * (2 * WT), m[u]);
                                    flip[u] = fl;
                                } else {
                                    const SlotWords<WT> hs = slot_hashed<WT>(s_fv, pb + pos);
#pragma unroll
                                    for (int w = 0; w < NWP; ++w) m[u][w] = hs.v[w];
                                }
                            }
                        }
                    }
                    // m[u] = [fwd words][rev words] of the slot's own k-mer; a flipped window answers for the
                    // reverse complement, so its strands trade places.  Plane word order q = 2*word32 + strand.
                    uint32_t mm[4][NWP];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
#pragma unroll
                        for (int w32 = 0; w32 < 2 * WT; ++w32) {
                            const uint32_t f = m[u][w32], r = m[u][2 * WT + w32];
                            mm[u][2 * w32] = flip[u] ? r : f;
                            mm[u][2 * w32 + 1] = flip[u] ? f : r;
                        }
                    }
#pragma unroll
                    for (int w = 0; w < NWP; ++w) {
                        uint32_t t1 = maj3(pl[0][w], mm[0][w], mm[1][w]);
                        pl[0][w] ^= mm[0][w] ^ mm[1][w];
                        uint32_t t2 = maj3(pl[0][w], mm[2][w], mm[3][w]);
                        pl[0][w] ^= mm[2][w] ^ mm[3][w];
                        uint32_t f4 = maj3(pl[1][w], t1, t2);
                        pl[1][w] ^= t1 ^ t2;
                        uint32_t c8 = pl[2][w] & f4;
                        pl[2][w] ^= f4;
                        pl[3][w] ^= c8;
                    }
                }
                uint32_t res[9];
                warp_fold<NWP>(pl, res, lane);
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
        for (int tt = 0; tt < kMaxLut; ++tt) {
            best[tt] = 0;
            if (tt < (int)a.n_lut && flag == 0) {
                const uint32_t thr = (uint32_t)__ldg(a.lut + (size_t)tt * kLutSize + len);
                const uint32_t pass = (bs_ge<NPA>(acc, thr) | bs_ge<NPA>(oth, thr)) & valid;
                if (pass) {
                    uint32_t s1 = pass, s2 = pass;
                    const uint32_t m1 = bs_max<NPA>(acc, s1), m2 = bs_max<NPA>(oth, s2);
                    const uint32_t mx = max(m1, m2);
                    const uint32_t at = (m1 == mx ? s1 : 0u) | (m2 == mx ? s2 : 0u);
                    best[tt] = pack_key(mx, (uint32_t)(a.fv.bin_begin + bin0 + (__ffs((int)at) - 1)));
                }
            }
        }
#pragma unroll
        for (int tt = 0; tt < kMaxLut; ++tt) {
            if (tt < (int)a.n_lut) {
                const uint64_t bk = warp_max_u64(best[tt]);
                if (lane == 0) a.keys[(size_t)tt * a.n_reads + read] = bk;
            }
        }
    }
}


// ------------------------------------------------------------------------------------------
// lookup kernel, short reads: one read per GROUP of G lanes (8 or 16 reads per warp)
// ------------------------------------------------------------------------------------------
// The warp-per-read kernel above spends most of its instructions outside the table loads: the 5-level
// butterfly that sums the 32 lanes' counters, the epilogue, the per-read setup.  For reads of at most
// 127*S positions (any 250-base chunk) the G lanes that load one entry keep the whole read instead:
// lane t adds the masks of slot t of every entry of ITS read into 7 bit planes, the fold has only
// log2(G) levels, and every instruction of fold and epilogue serves 32/G reads at once.
//
// The read's bases are loaded as aligned 16-byte blocks (lane t: blocks t, t+G, ...), turned into one
// 16-bit word per block and plane (low code bit, high code bit, "not ACGT") with SWAR arithmetic, and
// kept in shared memory as three bit streams; a block of 8 consecutive entries takes its windows from
// 64-bit pieces of those streams with compile-time shifts.
constexpr int kWgMaxBlocks = 36;                 // 16-base stream units per read (576 bits per plane)
constexpr int kWgMaxLen = 16 * kWgMaxBlocks - 31;   // longest read: the start can sit 31 bits into its first stream word
constexpr int kWgRow = kWgMaxBlocks + 12;        // + zero padding read by the last window block
constexpr int kWgIB = 8;                         // entries per window block
constexpr int kWgPlanes = 7;                     // per-lane counters hold 0..127

// 4 ASCII bases -> 4 low code bits, 4 high code bits (code = (c >> 1) & 3), and whether any of the four is
// not one of ACGTacgt (U/u are sent down the hashing path too, which reads them as T)
__device__ __forceinline__ void swar4(uint32_t x, uint32_t &lo4, uint32_t &hi4, uint32_t &bad)
{
    const uint32_t a = x >> 1;
    lo4 = ((a & 0x01010101u) * 0x01020408u) >> 24;
    hi4 = (((x >> 2) & 0x01010101u) * 0x01020408u) >> 24;
    const uint32_t y = a & 0x03030303u;                      // the codes, one per byte
    const uint32_t sel = __byte_perm(y | (y >> 4), 0u, 0x4420);   // code nibbles c0 c1 c2 c3 in the low 16 bits
    const uint32_t expect = __byte_perm(0x47544341u, 0u, sel);    // 'A' 'C' 'T' 'G' by code
    bad = ((x & 0xDFDFDFDFu) != expect) ? 0xFu : 0u;
}

// hashing path from the bit streams of a PACKED read (host-made planes: "bad" is exact per base and every bad
// base is Dna5 rank 4, because the packer codes U as T): k-mer starting at stream bit `bit0`
template <int WT>
__device__ __noinline__ SlotWords<WT> slot_hashed_streams(const FilterView &fv, const uint32_t *lo, const uint32_t *hi,
                                                          const uint32_t *bad, uint32_t bit0)
{
    SlotWords<WT> out;
    const HashParams &hp = fv.hp;
    uint64_t Hf = 0, Hr = 0, pw = 1;
    for (uint32_t u = 0; u < hp.k; ++u) {
        const uint32_t b = bit0 + u, w = b >> 5, sft = b & 31u;
        const uint32_t c = (((hi[w] >> sft) & 1u) << 1) | ((lo[w] >> sft) & 1u);
        const uint32_t d = ((bad[w] >> sft) & 1u) ? 4u : rank_of_code(c);
        Hf = Hf * 5 + d;
        Hr += comp5(d) * pw;
        pw *= 5;
    }
    uint64_t mf[WT], mr[WT];
#pragma unroll
    for (int w = 0; w < WT; ++w) { mf[w] = ~0ULL; mr[w] = ~0ULL; }
#pragma unroll 1
    for (uint32_t i = 0; i < hp.n_hash; ++i) {
        const uint64_t *pf = fv.words + hash_row(Hf, hp.pre[i], hp.n_blocks, hp.magic) * fv.stride;
        const uint64_t *pr = fv.words + hash_row(Hr, hp.pre[i], hp.n_blocks, hp.magic) * fv.stride;
#pragma unroll
        for (int w = 0; w < WT; ++w) { mf[w] &= __ldg(pf + w); mr[w] &= __ldg(pr + w); }
    }
#pragma unroll
    for (int w = 0; w < WT; ++w) {
        out.v[2 * w] = (uint32_t)mf[w];
        out.v[2 * w + 1] = (uint32_t)(mf[w] >> 32);
        out.v[2 * WT + 2 * w] = (uint32_t)mr[w];
        out.v[2 * WT + 2 * w + 1] = (uint32_t)(mr[w] >> 32);
    }
    return out;
}

template <int NP>
__device__ __forceinline__ void bs_max2(const uint32_t (&f)[NP], const uint32_t (&r)[NP], uint32_t (&mx)[NP])
{
    uint32_t gt = 0, eq = ~0u;                               // f > r, f == r so far (MSB first)
#pragma unroll
    for (int p = NP - 1; p >= 0; --p) {
        gt |= eq & f[p] & ~r[p];
        eq &= ~(f[p] ^ r[p]);
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) mx[p] = (f[p] & gt) | (r[p] & ~gt);
}

// Threshold test, max_matches and its lowest bin for the bins one lane holds after the fold:
// PAIRS (forward, reverse) word pairs of BITS significant bits each, NP planes.  select_matches is
// "fwd >= thr or rev >= thr", max_matches the largest max(fwd, rev) among the passing bins
// (src/IBF/IBFClassify.cpp:16-71); with M = the largest max(fwd, rev) over ALL bins both follow from
// M alone: some bin passes iff M >= thr, and then the bins attaining M pass.  So one (M, lowest bin)
// per read serves every threshold table.
template <int NP, int PAIRS, int BITS, int G>
__device__ __forceinline__ void group_epilogue(const CountArgs &a, uint64_t read, bool active, uint64_t len, uint32_t flag,
                                               const uint32_t (&cf)[PAIRS][NP], const uint32_t (&cr)[PAIRS][NP],
                                               const uint32_t (&bin0)[PAIRS], uint32_t t)
{
    const uint64_t nbl = a.fv.n_bins_local;
    uint32_t valid[PAIRS], sel[PAIRS], mx[PAIRS][NP];
#pragma unroll
    for (int pr = 0; pr < PAIRS; ++pr) {
        valid[pr] = 0;
        if (bin0[pr] < nbl)
            valid[pr] = (nbl - bin0[pr] >= (uint64_t)BITS) ? ((BITS == 32) ? ~0u : ((1u << BITS) - 1u))
                                                          : ((1u << (nbl - bin0[pr])) - 1u);
        bs_max2<NP>(cf[pr], cr[pr], mx[pr]);
        sel[pr] = valid[pr];
    }
    if (active && (a.counts_fwd || a.counts_rev)) {
#pragma unroll
        for (int pr = 0; pr < PAIRS; ++pr)
            for (int b = 0; b < BITS; ++b)
                if ((valid[pr] >> b) & 1u) {
                    if (a.counts_fwd) a.counts_fwd[read * nbl + bin0[pr] + b] = (uint16_t)bs_get<NP>(cf[pr], b);
                    if (a.counts_rev) a.counts_rev[read * nbl + bin0[pr] + b] = (uint16_t)bs_get<NP>(cr[pr], b);
                }
    }
    uint32_t val = 0;
#pragma unroll
    for (int p = NP - 1; p >= 0; --p) {
        uint32_t tt[PAIRS], any = 0;
#pragma unroll
        for (int pr = 0; pr < PAIRS; ++pr) { tt[pr] = sel[pr] & mx[pr][p]; any |= tt[pr]; }
        if (any) {
#pragma unroll
            for (int pr = 0; pr < PAIRS; ++pr) sel[pr] = tt[pr];
            val |= 1u << p;
        }
    }
    uint32_t key32 = 0;                                       // count << 16 | (0xFFFF - local bin); 0 = no valid bin here
#pragma unroll
    for (int pr = PAIRS - 1; pr >= 0; --pr)
        if (sel[pr]) key32 = (val << 16) | (0xFFFFu - (bin0[pr] + (uint32_t)__ffs((int)sel[pr]) - 1u));
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) key32 = max(key32, __shfl_xor_sync(kFull, key32, o));
    if (t == 0 && active) {
        const uint32_t M = key32 >> 16, bin = 0xFFFFu - (key32 & 0xFFFFu);
#pragma unroll
        for (int tt = 0; tt < kMaxLut; ++tt) {
            if (tt < (int)a.n_lut) {
                uint64_t key = 0;
                if (flag == 0) {
                    const uint32_t thr = (uint32_t)__ldg(a.lut + (size_t)tt * kLutSize + len);
                    if (M >= thr) key = pack_key(M, (uint32_t)(a.fv.bin_begin + bin));
                }
                a.keys[(size_t)tt * a.n_reads + read] = key;
            }
        }
    }
}

// PACKED: the bases arrive as the three bit streams already (a.pk_*, made by the host packer of the
// host-buffer API: 3 bits per base over PCIe instead of 8), bit i of the streams = base a.pk_base0 + i.
template <int WT, int S, int G, bool CANON, bool PACKED>
__global__ void __launch_bounds__(kTileWarps * 32, 2)
count_wgroup_kernel(const CountArgs a, const uint64_t *__restrict__ table)
{
    constexpr int NWP = 4 * WT;             // 32-bit mask words of both strands, order q = 2*word32 + strand
    constexpr int RPW = 32 / G;             // reads per warp
    constexpr int NPL = kWgPlanes;
    constexpr int IB = kWgIB;
    static_assert(S * (IB - 1) + 16 <= 64 && S * (IB - 1) < 32, "a window block must fit the 64-bit stream piece");
    __shared__ __align__(16) uint16_t s_stream[kTileWarps][RPW][3][kWgRow];
    // the out-of-line hashed path takes the filter by reference: one copy per CTA in shared memory instead of one per thread in
    // local memory (a 304-byte stack frame written by every thread at kernel start)
    __shared__ FilterView s_fv;
    if (threadIdx.x == 0) s_fv = a.fv;
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t g = (uint32_t)lane / G, t = (uint32_t)lane % G;
    const uint64_t total_warps = (uint64_t)gridDim.x * kTileWarps;
    const uint32_t k = a.fv.hp.k;
    const WGeom wg = make_wgeom(k, S);
    const uint32_t midmask = (1u << wg.mid) - 1u;
    // the position inside an entry that my slot answers, forward and flipped; slots >= S are padding
    const uint32_t tfwd = t < (uint32_t)S ? t : 0x40000000u;
    const uint32_t trev = t < (uint32_t)S ? (uint32_t)(S - 1) - t : 0x40000000u;
    const uint64_t *const my_table = table + (size_t)t * (2 * WT);
    uint16_t *const st_lo = s_stream[warp][g][0], *const st_hi = s_stream[warp][g][1], *const st_bad = s_stream[warp][g][2];

    for (uint64_t base = ((uint64_t)blockIdx.x * kTileWarps + warp) * RPW; base < a.n_reads; base += total_warps * RPW) {
        const uint64_t read = base + g;
        const bool active = read < a.n_reads;
        uint64_t off = 0, len = 0;
        if (active) { off = a.read_off[read]; len = a.read_off[read + 1] - off; }
        uint32_t flag = read_flag_of(len, k);
        const uint8_t *const pr = PACKED ? nullptr : a.bases + off;
        const uint64_t rel = PACKED ? off - a.pk_base0 : 0;                      // first base of the read in the streams
        const uint32_t sh = PACKED ? (uint32_t)(rel & 31u) : (uint32_t)(reinterpret_cast<uintptr_t>(pr) & 15u);
        // a read longer than the caller's max_read_len promised does not fit this kernel's counters: flag 3
        if (flag == 0 && (len - k + 1 > (uint64_t)S * ((1u << NPL) - 1u) || len > (uint64_t)kWgMaxLen)) flag = 3;
        if (t == 0 && active && a.read_flag) a.read_flag[read] = (uint8_t)flag;
        const uint32_t npos = (active && flag == 0) ? (uint32_t)len - k + 1 : 0u;

        // ---- bases -> bit streams ------------------------------------------------------------------
        __syncwarp();
        if constexpr (PACKED) {
            const uint32_t nw = npos ? (sh + (uint32_t)len + 31u) >> 5 : 0u;
            const uint64_t w0 = rel >> 5;
            uint32_t *const d_lo = reinterpret_cast<uint32_t *>(st_lo), *const d_hi = reinterpret_cast<uint32_t *>(st_hi),
                           *const d_bad = reinterpret_cast<uint32_t *>(st_bad);
            for (uint32_t j = t; j < nw + 4; j += G) {
                uint32_t l = 0, h = 0, b = 0;
                if (j < nw) { l = __ldg(a.pk_lo + w0 + j); h = __ldg(a.pk_hi + w0 + j); b = __ldg(a.pk_bad + w0 + j); }
                d_lo[j] = l; d_hi[j] = h; d_bad[j] = b;
            }
        } else {
            const uint4 *const pa = reinterpret_cast<const uint4 *>(pr - sh);
            const uint32_t nblk = npos ? (sh + (uint32_t)len + 15u) >> 4 : 0u;
            for (uint32_t j = t; j < nblk + 8; j += G) {
                uint32_t lo16 = 0, hi16 = 0, bad16 = 0;
                if (j < nblk) {
                    const uint4 v = __ldg(pa + j);
                    uint32_t l, h, b;
                    swar4(v.x, l, h, b); lo16 = l; hi16 = h; bad16 = b;
                    swar4(v.y, l, h, b); lo16 |= l << 4; hi16 |= h << 4; bad16 |= b << 4;
                    swar4(v.z, l, h, b); lo16 |= l << 8; hi16 |= h << 8; bad16 |= b << 8;
                    swar4(v.w, l, h, b); lo16 |= l << 12; hi16 |= h << 12; bad16 |= b << 12;
                }
                st_lo[j] = (uint16_t)lo16; st_hi[j] = (uint16_t)hi16; st_bad[j] = (uint16_t)bad16;
            }
        }
        __syncwarp();

        uint32_t pl[NPL][NWP];
#pragma unroll
        for (int p = 0; p < NPL; ++p)
#pragma unroll
            for (int w = 0; w < NWP; ++w) pl[p][w] = 0;

        const uint32_t n_blocks_mine = (npos + S * IB - 1) / (S * IB);
        const uint32_t n_blocks_warp = __reduce_max_sync(kFull, n_blocks_mine);
        for (uint32_t bi = 0; bi < n_blocks_warp; ++bi) {
            if (bi < n_blocks_mine) {
                // 64-bit pieces of the three streams, starting at the first base of this block's first window
                const uint32_t ob = sh + (uint32_t)(S * IB) * bi;
                const uint32_t wi = ob >> 5, r = ob & 31u;
                uint32_t lo0, lo1, hi0, hi1, bd0, bd1;
                {
                    const uint32_t *q = reinterpret_cast<const uint32_t *>(st_lo) + wi;
                    const uint32_t c0 = q[0], c1 = q[1], c2 = q[2];
                    lo0 = __funnelshift_r(c0, c1, r); lo1 = __funnelshift_r(c1, c2, r);
                    q = reinterpret_cast<const uint32_t *>(st_hi) + wi;
                    const uint32_t d0 = q[0], d1 = q[1], d2 = q[2];
                    hi0 = __funnelshift_r(d0, d1, r); hi1 = __funnelshift_r(d1, d2, r);
                    q = reinterpret_cast<const uint32_t *>(st_bad) + wi;
                    const uint32_t e0 = q[0], e1 = q[1], e2 = q[2];
                    bd0 = __funnelshift_r(e0, e1, r); bd1 = __funnelshift_r(e1, e2, r);
                }
                uint32_t f4[2][NWP];
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t m[4][NWP];
                    bool flip[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int su = S * (half * 4 + u);                        // compile-time shift
                        const uint32_t p0 = (uint32_t)(S * IB) * bi + (uint32_t)su;  // first position of the entry
                        flip[u] = false;
#pragma unroll
                        for (int w = 0; w < NWP; ++w) m[u][w] = 0;
                        const uint32_t wlo = __funnelshift_r(lo0, lo1, su) & wg.maskL;
                        uint32_t whi = __funnelshift_r(hi0, hi1, su) & wg.maskL;
                        const uint32_t wbad = __funnelshift_r(bd0, bd1, su) & wg.maskL;
                        uint32_t ilo = wlo;
                        bool fl = false;
                        if constexpr (CANON) {
                            fl = (whi >> wg.mid) & 1u;
                            if (fl) {
                                ilo = __brev(wlo) >> (32 - wg.L);
                                whi = __brev(~whi) >> (32 - wg.L);
                            }
                        }
                        const uint32_t pos = p0 + (fl ? trev : tfwd);
                        if (pos < npos) {
                            if (wbad == 0) {
                                uint32_t idx;
                                if constexpr (CANON) idx = ilo | ((whi & midmask) << wg.L) | ((whi >> (wg.mid + 1)) << (wg.L + wg.mid));
                                else idx = ilo | (whi << wg.L);
                                load_slot<WT>(my_table + (size_t)idx * (G * 2 * WT), m[u]);
                                flip[u] = fl;
                            } else {
                                SlotWords<WT> hs;
                                if constexpr (PACKED)
                                    hs = slot_hashed_streams<WT>(s_fv, reinterpret_cast<const uint32_t *>(st_lo),
                                                                 reinterpret_cast<const uint32_t *>(st_hi),
                                                                 reinterpret_cast<const uint32_t *>(st_bad), sh + pos);
                                else
                                    hs = slot_hashed<WT>(s_fv, pr + pos);
#pragma unroll
                                for (int w = 0; w < NWP; ++w) m[u][w] = hs.v[w];
                            }
                        }
                    }
                    // [fwd words][rev words] -> plane word order; a flipped window answers for the reverse
                    // complement, so its strands trade places
                    uint32_t mm[4][NWP];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int w32 = 0; w32 < 2 * WT; ++w32) {
                            const uint32_t f = m[u][w32], rv = m[u][2 * WT + w32];
                            mm[u][2 * w32] = flip[u] ? rv : f;
                            mm[u][2 * w32 + 1] = flip[u] ? f : rv;
                        }
#pragma unroll
                    for (int w = 0; w < NWP; ++w) {
                        const uint32_t t1 = maj3(pl[0][w], mm[0][w], mm[1][w]);
                        pl[0][w] ^= mm[0][w] ^ mm[1][w];
                        const uint32_t t2 = maj3(pl[0][w], mm[2][w], mm[3][w]);
                        pl[0][w] ^= mm[2][w] ^ mm[3][w];
                        f4[half][w] = maj3(pl[1][w], t1, t2);
                        pl[1][w] ^= t1 ^ t2;
                    }
                }
#pragma unroll
                for (int w = 0; w < NWP; ++w) {
                    uint32_t c = maj3(pl[2][w], f4[0][w], f4[1][w]);      // carry into the eights
                    pl[2][w] ^= f4[0][w] ^ f4[1][w];
#pragma unroll
                    for (int p = 3; p < NPL; ++p) {
                        const uint32_t c2 = pl[p][w] & c;
                        pl[p][w] ^= c;
                        c = c2;
                    }
                }
            }
        }

        // ---- fold over the G lanes of the group, then threshold / max ---------------------------------
        if constexpr (WT == 2 && G == 4) {
            uint32_t x1[NPL + 1][4], x2[NPL + 2][2];
            fold_words<8, NPL, 2>(pl, x1, lane);
            fold_words<4, NPL + 1, 1>(x1, x2, lane);
            uint32_t cf[1][NPL + 2], cr[1][NPL + 2];
#pragma unroll
            for (int p = 0; p < NPL + 2; ++p) { cf[0][p] = x2[p][0]; cr[0][p] = x2[p][1]; }
            const uint32_t bin0[1] = {32u * t};
            group_epilogue<NPL + 2, 1, 32, G>(a, read, active, len, flag, cf, cr, bin0, t);
        } else if constexpr (WT == 2 && G == 2) {
            uint32_t x1[NPL + 1][4];
            fold_words<8, NPL, 1>(pl, x1, lane);
            uint32_t cf[2][NPL + 1], cr[2][NPL + 1];
#pragma unroll
            for (int p = 0; p < NPL + 1; ++p) { cf[0][p] = x1[p][0]; cr[0][p] = x1[p][1]; cf[1][p] = x1[p][2]; cr[1][p] = x1[p][3]; }
            const uint32_t bin0[2] = {64u * t, 64u * t + 32u};
            group_epilogue<NPL + 1, 2, 32, G>(a, read, active, len, flag, cf, cr, bin0, t);
        } else if constexpr (WT == 1 && G == 4) {
            uint32_t x1[NPL + 1][2];
            fold_words<4, NPL, 2>(pl, x1, lane);
            uint32_t f8[NPL + 1], r8[NPL + 1], cf[1][NPL + 2], cr[1][NPL + 2];
#pragma unroll
            for (int p = 0; p < NPL + 1; ++p) { f8[p] = x1[p][0]; r8[p] = x1[p][1]; }
            fold_bits<32, NPL + 1, 1>(f8, cf[0], lane);
            fold_bits<32, NPL + 1, 1>(r8, cr[0], lane);
            const uint32_t bin0[1] = {32u * (t >> 1) + 16u * (t & 1u)};
            group_epilogue<NPL + 2, 1, 16, G>(a, read, active, len, flag, cf, cr, bin0, t);
        } else {
            static_assert(WT == 1 && G == 2, "unsupported window-table shape");
            uint32_t x1[NPL + 1][2];
            fold_words<4, NPL, 1>(pl, x1, lane);
            uint32_t cf[1][NPL + 1], cr[1][NPL + 1];
#pragma unroll
            for (int p = 0; p < NPL + 1; ++p) { cf[0][p] = x1[p][0]; cr[0][p] = x1[p][1]; }
            const uint32_t bin0[1] = {32u * t};
            group_epilogue<NPL + 1, 1, 32, G>(a, read, active, len, flag, cf, cr, bin0, t);
        }
    }
}

template <int WT, int S, int G, bool CANON>
void launch_build(const FilterView &fv, uint64_t *table, uint64_t n_entries, int sm_count, cudaStream_t st)
{
    const uint64_t blocks = (n_entries * G + 255) / 256;
    const uint64_t cap = (uint64_t)sm_count * 32;
    wtable_build_kernel<WT, S, G, CANON><<<(uint32_t)(blocks < cap ? blocks : cap), 256, 0, st>>>(fv, table, n_entries);
}

template <int WT, int S, int G, bool CANON, int NPA>
void launch_count_one(const CountArgs &a, const uint64_t *table, int sm_count, cudaStream_t st)
{
    static int occ = 0;
    if (occ == 0) {
        int o = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, count_wtable_kernel<WT, S, G, CANON, NPA>, kTileWarps * 32, 0);
        occ = o > 0 ? o : 1;
    }
    const uint64_t blocks_needed = (a.n_reads + kTileWarps - 1) / kTileWarps;
    const uint64_t max_x = (uint64_t)sm_count * occ * grid_waves();
    const uint32_t gx = (uint32_t)(blocks_needed < max_x ? blocks_needed : max_x);
    count_wtable_kernel<WT, S, G, CANON, NPA><<<gx ? gx : 1, kTileWarps * 32, 0, st>>>(a, table);
}

template <int WT, int S, int G, bool CANON, bool PACKED>
void launch_count_group(const CountArgs &a, const uint64_t *table, int sm_count, cudaStream_t st)
{
    static int occ = 0;
    if (occ == 0) {
        int o = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, count_wgroup_kernel<WT, S, G, CANON, PACKED>, kTileWarps * 32, 0);
        occ = o > 0 ? o : 1;
    }
    constexpr uint64_t reads_per_cta = (uint64_t)kTileWarps * (32 / G);
    const uint64_t blocks_needed = (a.n_reads + reads_per_cta - 1) / reads_per_cta;
    const uint64_t max_x = (uint64_t)sm_count * occ * grid_waves();
    const uint32_t gx = (uint32_t)(blocks_needed < max_x ? blocks_needed : max_x);
    count_wgroup_kernel<WT, S, G, CANON, PACKED><<<gx ? gx : 1, kTileWarps * 32, 0, st>>>(a, table);
}

template <int WT, int S, int G, bool CANON>
void launch_count_np(const CountArgs &a, const uint64_t *table, uint32_t max_read_len, int sm_count, cudaStream_t st)
{
    // short reads (every 250-base chunk): one read per group of G lanes
    if (a.pk_lo) {                                   // bit streams from the host packer: the caller checked the limits
        launch_count_group<WT, S, G, CANON, true>(a, table, sm_count, st);
        return;
    }
    if (wgroup_applicable(a.fv.hp.k, S, max_read_len) && g_wtable_variant != 1) {
        launch_count_group<WT, S, G, CANON, false>(a, table, sm_count, st);
        return;
    }
    const bool single = max_read_len != 0 && max_read_len < a.fv.hp.k + (uint32_t)WtShape<S, G>::PIECE;
    if (single) launch_count_one<WT, S, G, CANON, 9>(a, table, sm_count, st);
    else launch_count_one<WT, S, G, CANON, 16>(a, table, sm_count, st);
}

}  // namespace

// Geometry of the window table for a span (k-mers per entry): lanes per entry, canonical or not, entries.
// Returns false when the span is not supported for this filter.
bool wtable_geometry(uint64_t stride, uint32_t k, int span, int *lanes, int *canon, uint64_t *n_entries)
{
    if (stride < 1 || stride > 2 || span < 2 || span > 4) return false;
    const uint32_t L = k + (uint32_t)span - 1;
    if (L > 16 || L < 3) return false;
    const int G = span == 2 ? 2 : 4;
    const bool c = (L & 1u) != 0;
    if (lanes) *lanes = G;
    if (canon) *canon = c ? 1 : 0;
    if (n_entries) *n_entries = c ? (1ull << (2 * L - 1)) : (1ull << (2 * L));
    return true;
}

void set_wtable_variant(int v) { g_wtable_variant = v; }
int get_wtable_variant() { return g_wtable_variant; }

// reads of a launch fit the group-per-read kernel: known longest read, <= 127*span positions, <= kWgMaxLen bases
bool wgroup_applicable(uint32_t k, int span, uint32_t max_read_len)
{
    return max_read_len != 0 && max_read_len <= (uint32_t)kWgMaxLen &&
           (max_read_len < k || max_read_len - k + 1 <= (uint32_t)span * ((1u << kWgPlanes) - 1u));
}

#define RB_WT_DISPATCH(FN, ...)                                                                              \
    do {                                                                                                     \
        const int key = (int)stride * 100 + span * 10 + (canon ? 1 : 0);                                     \
        switch (key) {                                                                                       \
        case 120: FN<1, 2, 2, false>(__VA_ARGS__); break;                                                    \
        case 121: FN<1, 2, 2, true>(__VA_ARGS__); break;                                                     \
        case 130: FN<1, 3, 4, false>(__VA_ARGS__); break;                                                    \
        case 131: FN<1, 3, 4, true>(__VA_ARGS__); break;                                                     \
        case 140: FN<1, 4, 4, false>(__VA_ARGS__); break;                                                    \
        case 141: FN<1, 4, 4, true>(__VA_ARGS__); break;                                                     \
        case 220: FN<2, 2, 2, false>(__VA_ARGS__); break;                                                    \
        case 221: FN<2, 2, 2, true>(__VA_ARGS__); break;                                                     \
        case 230: FN<2, 3, 4, false>(__VA_ARGS__); break;                                                    \
        case 231: FN<2, 3, 4, true>(__VA_ARGS__); break;                                                     \
        case 240: FN<2, 4, 4, false>(__VA_ARGS__); break;                                                    \
        case 241: FN<2, 4, 4, true>(__VA_ARGS__); break;                                                     \
        default: return -1;                                                                                  \
        }                                                                                                    \
    } while (0)

int launch_wtable_build(const FilterView &fv, uint64_t *table, int span, int sm_count, cudaStream_t st)
{
    int lanes = 0, canon = 0;
    uint64_t n_entries = 0;
    const uint64_t stride = fv.stride;
    if (!wtable_geometry(stride, fv.hp.k, span, &lanes, &canon, &n_entries)) return -1;
    RB_WT_DISPATCH(launch_build, fv, table, n_entries, sm_count, st);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_count_wtable(const CountArgs &a, const uint64_t *table, int span, uint32_t max_read_len, int sm_count,
                        cudaStream_t st)
{
    if (a.n_reads == 0) return 0;
    if (a.n_lut == 0 || a.n_lut > (uint32_t)kMaxLut) return -1;
    int lanes = 0, canon = 0;
    const uint64_t stride = a.fv.stride;
    if (!wtable_geometry(stride, a.fv.hp.k, span, &lanes, &canon, nullptr)) return -1;
    RB_WT_DISPATCH(launch_count_np, a, table, max_read_len, sm_count, st);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace rb
