// ibf_postings.cu -- k-mer postings table for WIDE filters (rows of more than 4 words; BASELINE configs #3, #5).
//
// What seqan::count does per k-mer and strand (src/IBF/IBFClassify.cpp:149-150) is a pure
// function of the k-mer: the AND of the h rows its hashes select.  For a wide filter that AND is a row of
// thousands of bits -- 3 880 bytes for the 31 008 bins of a human reference -- but it is SPARSE: a bin's bit
// survives the AND only if the bin holds the k-mer or all h probes are false positives, about 1 % by the
// reference's own sizing (max_fp = 0.01, IBFBuild.cpp:404-413).  Tabulated densely the function needs 260 GB
// (k = 13); as postings -- the list of set bins per k-mer, 2 bytes each -- it needs ~50 GB and fits HBM.
//
//   ptr[x], ptr[x+1]   list of k-mer x in units of 8 ids (16 bytes), x = sum rank_j * 4^(k-1-j), ranks A0 C1 G2 T3
//   ids[8 * u + i]     local bin indices; lists are padded to a multiple of 8 with a sentinel (n_bins_local rounded up to
//                      a multiple of 16: its counter lies behind the quads of counter words that hold bins).  Inside a
//                      list the ids are dealt over the groups one ATOMS instruction serves so that few of them share a
//                      shared-memory bank (ibf_postings_layout.cuh); RB_POSTINGS_ORDER=0 keeps them ascending
//
// Classifying a 250-base chunk then reads 476 lists (~340 KB) instead of streaming 2 x 238 x 3 rows (5.5 MB), and
// counts with shared-memory atomics on packed 8-bit (or 16-bit) counters.  The reverse strand of k-mer x is the
// list of revcomp(x).  Windows containing a non-ACGT base take the hashing path on the bit matrix itself, so
// every output equals the reference's.
#include "ibf_device.cuh"
#include "ibf_postings_layout.cuh"

#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstdlib>
#include <vector>

namespace rb {

namespace {

// Threads per CTA (template parameter of the lookup kernel) follow from how many CTAs of counters fit the shared memory
// of an SM: 3 x 256, 2 x 384 or 1 x 768 -- always 24 warps per SM under the 85-register ceiling of 768 threads.
constexpr int kPostPiece = 512;               // k-mer positions staged per pass over a read
constexpr double kSub2Units = 3.5, kSub4Units = 7.5, kSub8Units = 16.0;   // mean list length (16-byte units) up to which 2 / 4 / 8 lanes own a list
constexpr int kPostInFlight = 4;              // lists a warp loads before it counts them (even: strands alternate)

// base-5 hashes of the forward and reverse-complement strand of the ACGT k-mer x (ranks, first base most significant)
__device__ __forceinline__ void kmer_hashes(uint64_t x, uint32_t k, uint64_t &Hf, uint64_t &Hr)
{
    Hf = 0; Hr = 0;
    uint64_t pw = 1;
    for (uint32_t j = 0; j < k; ++j) {
        const uint32_t d = (uint32_t)(x >> (2 * (k - 1 - j))) & 3u;
        Hf = Hf * 5 + d;
        Hr += (uint64_t)(3u - d) * pw;
        pw *= 5;
    }
}

// ------------------------------------------------------------------------------------------
// build: pass 1 counts, pass 2 fills; one warp per k-mer, lanes stride over the row words
// ------------------------------------------------------------------------------------------
// x = first + i * step for i < n (step > 1: a sample to estimate the table size)
// raw: write the number of ids instead of the number of 8-id units
__global__ void __launch_bounds__(256) postings_count_kernel(const FilterView fv, uint64_t first, uint64_t step, uint64_t n,
                                                             uint32_t *__restrict__ units, const int raw = 0)
{
    const HashParams &hp = fv.hp;
    const int lane = threadIdx.x & 31;
    const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t i = warp0; i < n; i += n_warps) {
        uint64_t Hf, Hr;
        kmer_hashes(first + i * step, hp.k, Hf, Hr);
        const uint64_t *rows[kMaxHash];
#pragma unroll
        for (int h = 0; h < kMaxHash; ++h)
            rows[h] = (uint32_t)h < hp.n_hash ? fv.words + hash_row(Hf, hp.pre[h], hp.n_blocks, hp.magic) * fv.stride : nullptr;
        uint32_t c = 0;
        for (uint64_t w = lane; w < fv.stride; w += 32) {
            uint64_t m = ~0ULL;
#pragma unroll
            for (int h = 0; h < kMaxHash; ++h)
                if ((uint32_t)h < hp.n_hash) m &= __ldg(rows[h] + w);
            c += (uint32_t)__popcll(m);
        }
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0) units[i] = raw ? c : (c + 7u) >> 3;
    }
}

// Padding id of the lists: the first id whose counter lies in the 16-byte quad of counter words AFTER the quads that hold
// bins, for 8-bit (4 per word) and 16-bit (2 per word) counters alike, so the epilogue scans whole quads of bins only.
__host__ __device__ inline uint32_t postings_sentinel(uint64_t n_bins_local) { return (uint32_t)((n_bins_local + 15) / 16 * 16); }

constexpr int kFillStage = 2048;              // ids of one list staged per warp for the reordering (longer lists stay ascending)

__global__ void __launch_bounds__(256) postings_fill_kernel(const FilterView fv, uint64_t n_kmers, const uint32_t *__restrict__ ptr,
                                                            uint16_t *__restrict__ ids, const int deal)
{
    __shared__ uint16_t s_stage[8][kFillStage];
    __shared__ uint32_t s_off[8][32];
    const HashParams &hp = fv.hp;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint16_t sentinel = (uint16_t)postings_sentinel(fv.n_bins_local);
    uint16_t *const stage = s_stage[wib];
    uint32_t *const off = s_off[wib];
    for (uint64_t x = warp0; x < n_kmers; x += n_warps) {
        uint64_t Hf, Hr;
        kmer_hashes(x, hp.k, Hf, Hr);
        const uint64_t *rows[kMaxHash];
#pragma unroll
        for (int h = 0; h < kMaxHash; ++h)
            rows[h] = (uint32_t)h < hp.n_hash ? fv.words + hash_row(Hf, hp.pre[h], hp.n_blocks, hp.magic) * fv.stride : nullptr;
        uint16_t *const out = ids + (uint64_t)ptr[x] * 8;
        const uint64_t end = (uint64_t)ptr[x + 1] * 8 - (uint64_t)ptr[x] * 8;
        const bool staged = deal && end <= (uint64_t)kFillStage;
        uint16_t *const dst = staged ? stage : out;
        uint64_t run = 0;                                           // ids written so far (same in all lanes)
        for (uint64_t w0 = 0; w0 < fv.stride; w0 += 32) {
            const uint64_t w = w0 + lane;
            uint64_t m = 0;
            if (w < fv.stride) {
                m = ~0ULL;
#pragma unroll
                for (int h = 0; h < kMaxHash; ++h)
                    if ((uint32_t)h < hp.n_hash) m &= __ldg(rows[h] + w);
            }
            const uint32_t c = (uint32_t)__popcll(m);
            uint32_t incl = c;                                      // inclusive prefix over the lanes
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            uint64_t pos = run + incl - c;
            while (m) {
                const int b = __ffsll((long long)m) - 1;
                m &= m - 1;
                dst[pos++] = (uint16_t)(w * 64 + b);
            }
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
        for (uint64_t p = run + lane; p < end; p += 32) out[p] = sentinel;
        if (!staged) continue;
        // ---- deal the staged (ascending) ids over the groups of the list, bank by bank (ibf_postings_layout.cuh) ----
        const uint32_t n = (uint32_t)run;
        off[lane] = 0;
        __syncwarp();
        for (uint32_t i = lane; i < n; i += 32) atomicAdd(&off[counter_bank(stage[i])], 1u);
        __syncwarp();
        {
            const uint32_t load = off[lane];
            uint32_t incl = load;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            __syncwarp();
            off[lane] = incl - load;                                // first dealing index of bank `lane`
        }
        __syncwarp();
        const ListShape shape = list_shape(n);
        for (uint32_t i0 = 0; i0 < n; i0 += 32) {                   // rank inside the bank = ascending order (deterministic)
            const uint32_t i = i0 + lane;
            const bool valid = i < n;
            const uint32_t id = valid ? stage[i] : 0u;
            const uint32_t b = counter_bank(id);
            const uint32_t same = __match_any_sync(0xffffffffu, valid ? b : 32u + lane);
            const uint32_t before = __popc(same & ((1u << lane) - 1u));
            const uint32_t c = valid ? off[b] + before : 0u;
            __syncwarp();
            if (valid && before == 0) off[b] += __popc(same);
            __syncwarp();
            if (valid) out[list_position(shape, c)] = (uint16_t)id;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// lookup: CTA per read, warps take (position, strand) pairs, counters in shared memory
// ------------------------------------------------------------------------------------------
// CB = counter bits (8: reads of <= 255 positions, 16: any read the API accepts)
// One increment.  `v` holds the id in its low 16 bits; whatever sits above is ignored (the byte offset of the counter word
// is a mask of the low bits, the shift amount is taken modulo 32 by the funnel shift), so the low id of a pair needs no
// extraction at all.
template <int CB>
__device__ __forceinline__ void bump(uint32_t *cnt, const uint32_t v)
{
    if (CB == 8)
        atomicAdd(reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(cnt) + (v & 0xFFFCu)), __funnelshift_l(0u, 1u, v << 3));
    else
        atomicAdd(reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(cnt) + ((v & 0xFFFEu) << 1)), __funnelshift_l(0u, 1u, v << 4));
}

template <int CB>
__device__ __forceinline__ void bump_pair(uint32_t *cnt, const uint32_t w)
{
    bump<CB>(cnt, w);
    bump<CB>(cnt, w >> 16);
}

template <int CB>
__device__ __forceinline__ void add_ids(uint32_t *cnt, const uint4 v)
{
    bump_pair<CB>(cnt, v.x);
    bump_pair<CB>(cnt, v.y);
    bump_pair<CB>(cnt, v.z);
    bump_pair<CB>(cnt, v.w);
}

// One list as the warp walks it (layout: ibf_postings_layout.cuh): the first full round and the tail are loaded up front
// (`fetch`), so that several lists are in flight before the first counter is touched; `count` adds them and walks the
// further rounds of a list of more than 511 ids.
struct ListRegs {
    uint4 full, tail;
    uint32_t u0, n_u;          // first unit, units (warp-uniform)
};

__device__ __forceinline__ void fetch_list(ListRegs &r, const uint4 *__restrict__ ids, const uint32_t u0, const uint32_t n_u, const int lane)
{
    r.u0 = u0; r.n_u = n_u;
    r.full = make_uint4(0, 0, 0, 0); r.tail = make_uint4(0, 0, 0, 0);
    if (n_u >= 32u) r.full = __ldg(ids + u0 + lane);
    const uint32_t tu = n_u & 31u, ub = u0 + (n_u & ~31u);
    if (tu > 16u) { if ((uint32_t)lane < tu) r.tail = __ldg(ids + ub + lane); }
    else if (tu > 8u) {
        if ((uint32_t)lane < 2u * tu) { const uint2 t = __ldg(reinterpret_cast<const uint2 *>(ids + ub) + lane); r.tail.x = t.x; r.tail.y = t.y; }
    } else if (tu > 4u) { if ((uint32_t)lane < 4u * tu) r.tail.x = __ldg(reinterpret_cast<const uint32_t *>(ids + ub) + lane); }
    else if (tu > 0u) { if ((uint32_t)lane < 8u * tu) r.tail.x = __ldg(reinterpret_cast<const uint16_t *>(ids + ub) + lane); }
}

template <int CB>
__device__ __forceinline__ void count_list(const ListRegs &r, uint32_t *cnt, const uint4 *__restrict__ ids, const int lane,
                                           const uint32_t sentinel)
{
    const uint32_t n_u = r.n_u;
    if (n_u >= 32u) {
        add_ids<CB>(cnt, r.full);
#pragma unroll 1
        for (uint32_t rr = 1; rr < (n_u >> 5); ++rr) add_ids<CB>(cnt, __ldg(ids + r.u0 + 32u * rr + lane));     // > 511 ids
    }
    const uint32_t tu = n_u & 31u;
    if (tu > 16u) { if ((uint32_t)lane < tu) add_ids<CB>(cnt, r.tail); }
    else if (tu > 8u) { if ((uint32_t)lane < 2u * tu) { bump_pair<CB>(cnt, r.tail.x); bump_pair<CB>(cnt, r.tail.y); } }
    else if (tu > 4u) {
        // the few sentinels of a short tail would meet in one instruction: skip them
        if ((uint32_t)lane < 4u * tu) {
            if ((r.tail.x & 0xFFFFu) != sentinel) bump<CB>(cnt, r.tail.x);
            if ((r.tail.x >> 16) != sentinel) bump<CB>(cnt, r.tail.x >> 16);
        }
    } else if (tu > 0u) { if ((uint32_t)lane < 8u * tu && r.tail.x != sentinel) bump<CB>(cnt, r.tail.x); }
}

// hashing path of one (position, strand): AND of the probed rows, one warp, counters by atomics
template <int CB>
__device__ __forceinline__ void add_hashed_inl(const FilterView &fv, const uint8_t *dig, uint32_t strand, uint32_t *cnt, int lane)
{
    constexpr int PER = 32 / CB;
    constexpr int SH = (PER == 4) ? 2 : 1;
    const HashParams &hp = fv.hp;
    uint64_t H = 0, pw = 1;
    for (uint32_t u = 0; u < hp.k; ++u) {
        const uint32_t d = dig[u];
        if (strand == 0) H = H * 5 + d;
        else { H += comp5(d) * pw; pw *= 5; }
    }
    for (uint64_t w = lane; w < fv.stride; w += 32) {
        uint64_t m = ~0ULL;
        for (uint32_t h = 0; h < hp.n_hash; ++h)
            m &= __ldg(fv.words + hash_row(H, hp.pre[h], hp.n_blocks, hp.magic) * fv.stride + w);
        while (m) {
            const uint32_t id = (uint32_t)(w * 64) + (uint32_t)__ffsll((long long)m) - 1u;
            m &= m - 1;
            atomicAdd(cnt + (id >> SH), 1u << ((id & (PER - 1)) * CB));
        }
    }
}

// Out of line for the lane-group kernels (inlined there it costs them 7-18 %: more registers in their hot loop); inlined into
// the warp-per-list kernel, whose CTAs then need no stack frame: 6.5 -> 5.8 ms per batch on config #3
// (profiles/r2_aq_add_hashed_inline.jsonl).
// TAG makes separate copies: ptxas gives a callee ONE register allocation, which all kernels that call it inherit -- called with
// the shared-memory filter view of the list kernel it needs 72-80 registers, and the slot kernel, which shares nothing else with
// that kernel, ran 17 % slower for it.
template <int CB, int TAG = 0>
__device__ __noinline__ void add_hashed(const FilterView &fv, const uint8_t *dig, uint32_t strand, uint32_t *cnt, int lane)
{
    add_hashed_inl<CB>(fv, dig, strand, cnt, lane);
}

template <int CB>
__device__ __forceinline__ uint32_t vmax(uint32_t a, uint32_t b) { return CB == 8 ? __vmaxu4(a, b) : __vmaxu2(a, b); }

// Per-read epilogue of the CTA-per-read kernels: M = max over bins of max(fwd, rev) and its lowest bin, one key per threshold
// table; the counters go back to zero.  Called by all threads after a __syncthreads().
template <int CB, int kPostThreads>
__device__ __forceinline__ void postings_epilogue(const CountArgs &a, const uint64_t read, const uint64_t len, const uint32_t flag,
                                                  uint32_t *const cntF, uint32_t *const cntR, uint32_t *const s_red)
{
    constexpr int PER = 32 / CB;
    constexpr uint32_t CMASK = (CB == 8) ? 0xFFu : 0xFFFFu;
    constexpr int kPostWarps = kPostThreads / 32;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint64_t nbl = a.fv.n_bins_local;
    // ---- epilogue: M = max over bins of max(fwd, rev), its lowest bin; counters back to zero ----------------
    // The sentinel's counter lies in the quad of words after the bins: it is not scanned, only cleared.
    const uint32_t sent_word = postings_sentinel(nbl) / PER;
    const uint32_t scan_quads = ((uint32_t)nbl + 4u * PER - 1u) / (4u * PER);
    if (tid == 0) { cntF[sent_word] = 0; cntR[sent_word] = 0; }
    if (a.counts_fwd || a.counts_rev) {                                 // dense counts on request (tests, tools)
        for (uint32_t w = tid; w * PER < nbl; w += kPostThreads) {
            const uint32_t f = cntF[w], r = cntR[w];
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const uint32_t bin = w * PER + i;
                if (bin < nbl) {
                    if (a.counts_fwd) a.counts_fwd[read * nbl + bin] = (uint16_t)((f >> (i * CB)) & CMASK);
                    if (a.counts_rev) a.counts_rev[read * nbl + bin] = (uint16_t)((r >> (i * CB)) & CMASK);
                }
            }
        }
        __syncthreads();
    }
    // One pass over quads of counter words (LDS.128 / STS.128): every thread keeps the largest count it has seen and
    // the first bin that had it (its bins ascend); a quad is unpacked only when the SWAR test says that one of its
    // counters may beat that.  A maximum below every threshold of the read yields key 0 whatever it is, so the search
    // starts at min(thr) - 1 and the unpacking is rare (random reads: never).
    constexpr uint32_t LOW = CB == 8 ? 0x7F7F7F7Fu : 0x7FFF7FFFu, ONES = CB == 8 ? 0x01010101u : 0x00010001u;
    constexpr uint32_t HALF = (CMASK + 1u) / 2u;                        // 128 or 32 768
    uint32_t thr_min = CMASK + 1u;                                      // not classified: nothing to find
    if (flag == 0)
        for (uint32_t t = 0; t < a.n_lut; ++t) thr_min = min(thr_min, (uint32_t)__ldg(a.lut + (size_t)t * kLutSize + len));
    uint32_t cur = min(thr_min, CMASK + 1u) - (thr_min ? 1u : 0u), cur_bin = 0;
    // counter c > cur  =>  c >= HALF, or (c & LOW) + (HALF - 1 - cur) carries into the top bit (cur < HALF);
    // for cur >= HALF only the first test is left: necessary, not sufficient -- the unpacking decides
    uint32_t add = (HALF - 1u - min(cur, HALF - 1u)) * ONES;
    uint4 *const qF = reinterpret_cast<uint4 *>(cntF), *const qR = reinterpret_cast<uint4 *>(cntR);
    for (uint32_t qd = tid; qd < scan_quads; qd += kPostThreads) {
        const uint4 f = qF[qd], r = qR[qd];
        qF[qd] = make_uint4(0, 0, 0, 0); qR[qd] = make_uint4(0, 0, 0, 0);
        const uint32_t gx = f.x | r.x, gy = f.y | r.y, gz = f.z | r.z, gw = f.w | r.w;
        const uint32_t h = (((gx & LOW) + add) | gx) | (((gy & LOW) + add) | gy) | (((gz & LOW) + add) | gz) | (((gw & LOW) + add) | gw);
        if (h & ~LOW) {
            const uint32_t m[4] = {vmax<CB>(f.x, r.x), vmax<CB>(f.y, r.y), vmax<CB>(f.z, r.z), vmax<CB>(f.w, r.w)};
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < PER; ++i) {
                    const uint32_t v = (m[j] >> (i * CB)) & CMASK;
                    if (v > cur) { cur = v; cur_bin = (4u * qd + j) * PER + i; }
                }
            add = (HALF - 1u - min(cur, HALF - 1u)) * ONES;
        }
    }
    // (count, lowest bin) as one key: larger count wins, then the smaller bin (bins < 65 535)
    uint32_t best_key = __reduce_max_sync(0xffffffffu, (cur << 16) | (0xFFFFu - cur_bin));
    if (lane == 0) s_red[warp] = best_key;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kPostWarps; ++i) best_key = max(best_key, s_red[i]);
    const uint32_t M = best_key >> 16, best_bin = 0xFFFFu - (best_key & 0xFFFFu);
    if (tid < (int)a.n_lut) {
        uint64_t key = 0;
        if (flag == 0) {
            const uint32_t thr = (uint32_t)__ldg(a.lut + (size_t)tid * kLutSize + len);
            if (M >= thr) key = pack_key(M, (uint32_t)(a.fv.bin_begin + best_bin));
        }
        uint64_t *const dst = a.keys + (size_t)tid * a.n_reads + read;
        if (a.keys_shared) { if (key) key_max(dst, key, 1); }          // all bin shards fold into one array (NVLink peer atomics)
        else *dst = key;
    }
}

template <int CB, int kPostThreads>
__global__ void __launch_bounds__(kPostThreads, 768 / kPostThreads)
count_postings_kernel(const CountArgs a, const uint32_t *__restrict__ ptr, const uint4 *__restrict__ ids, const uint32_t cnt_words)
{
    constexpr int PER = 32 / CB;
    constexpr uint32_t CMASK = (CB == 8) ? 0xFFu : 0xFFFFu;
    constexpr int kPostWarps = kPostThreads / 32;
    extern __shared__ __align__(16) uint32_t s_mem[];
    uint32_t *const cntF = s_mem, *const cntR = s_mem + cnt_words;
    uint32_t *const s_x = s_mem + 2 * cnt_words;                         // [kPostPiece] packed k-mer or ~0u (not ACGT)
    uint8_t *const s_dig = reinterpret_cast<uint8_t *>(s_x + kPostPiece); // [kPostPiece + 32] Dna5 ranks
    __shared__ uint32_t s_red[kPostWarps];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t k = a.fv.hp.k;
    const uint32_t kbits = 2 * k;
    const uint32_t kmask = kbits >= 32 ? ~0u : ((1u << kbits) - 1u);
    const uint64_t nbl = a.fv.n_bins_local;
    const uint32_t sentinel = postings_sentinel(nbl);

    for (uint32_t w = tid; w < 2 * cnt_words; w += kPostThreads) s_mem[w] = 0;

    for (uint64_t read = blockIdx.x; read < a.n_reads; read += gridDim.x) {
        const uint64_t off = a.read_off[read];
        const uint64_t len = a.read_off[read + 1] - off;
        uint32_t flag = read_flag_of(len, k);
        if (flag == 0 && CB == 8 && len - k + 1 > 255) flag = 3;         // longer than the caller's max_read_len promised
        if (tid == 0 && a.read_flag) a.read_flag[read] = (uint8_t)flag;
        __syncthreads();                                                   // counters are zero and visible

        if (flag == 0) {
            const uint32_t npos = (uint32_t)len - k + 1;
            for (uint32_t cs = 0; cs < npos; cs += kPostPiece) {
                const uint32_t cn = min((uint32_t)kPostPiece, npos - cs);
                __syncthreads();
                for (uint32_t i = tid; i < cn + k - 1; i += kPostThreads) s_dig[i] = (uint8_t)dna5(a.bases[off + cs + i]);
                __syncthreads();
                for (uint32_t j = tid; j < cn; j += kPostThreads) {
                    uint32_t x = 0, bad = 0;
                    for (uint32_t u = 0; u < k; ++u) {
                        const uint32_t d = s_dig[j + u];
                        x = (x << 2) | (d & 3u);
                        bad |= d >> 2;
                    }
                    s_x[j] = bad ? ~0u : (x & kmask);
                }
                __syncthreads();
                // (position, strand) pairs: 32 per warp round, list bounds fetched by the lanes, lists walked by the warp
                // every warp takes an equal, contiguous share of the pairs (even, so that strands alternate from its start) in
                // rounds of 32; the list bounds of the next round are requested before the current one is walked
                const uint32_t n_pairs = 2 * cn;
                const uint32_t share = 2u * ((cn + kPostWarps - 1) / kPostWarps);
                const uint32_t q_end = min(n_pairs, (warp + 1) * share);
                auto list_bounds = [&](const uint32_t q, uint32_t &p0, uint32_t &p1, bool &hashed) {
                    p0 = 0; p1 = 0; hashed = false;
                    if (q < q_end) {
                        const uint32_t x = s_x[q >> 1];
                        if (x == ~0u) hashed = true;
                        else {
                            uint32_t idx = x;
                            if (q & 1u) {                                  // reverse strand: the list of revcomp(x)
                                uint32_t v = __brev(~x);
                                v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
                                idx = v >> (32 - kbits);
                            }
                            p0 = __ldg(ptr + idx);
                            p1 = __ldg(ptr + idx + 1);
                        }
                    }
                };
                uint32_t p0, p1, np0 = 0, np1 = 0;
                bool hashed, nhashed = false;
                list_bounds(warp * share + lane, p0, p1, hashed);
                for (uint32_t q0 = warp * share; q0 < q_end; q0 += 32) {
                    if (q0 + 32 < q_end) list_bounds(q0 + 32 + lane, np0, np1, nhashed);
                    const uint32_t any_hashed = __ballot_sync(0xffffffffu, hashed);
                    const uint32_t n_here = min(32u, q_end - q0);
                    for (uint32_t t = 0; t < n_here; t += kPostInFlight) {  // kPostInFlight lists in flight
                        ListRegs lr[kPostInFlight];
#pragma unroll
                        for (int j = 0; j < kPostInFlight; ++j) {
                            const uint32_t tj = min(t + j, 31u);
                            const uint32_t u0 = __shfl_sync(0xffffffffu, p0, tj), u1 = __shfl_sync(0xffffffffu, p1, tj);
                            fetch_list(lr[j], ids, u0, t + j < n_here ? u1 - u0 : 0u, lane);
                        }
#pragma unroll
                        for (int j = 0; j < kPostInFlight; ++j)         // q0 and t are even: list j is strand j & 1
                            count_list<CB>(lr[j], (j & 1) ? cntR : cntF, ids, lane, sentinel);
                    }
                    if (any_hashed) {
                        for (uint32_t t = 0; t < n_here; ++t)
                            if ((any_hashed >> t) & 1u)
                                add_hashed_inl<CB>(a.fv, s_dig + ((q0 + t) >> 1), (q0 + t) & 1u, ((q0 + t) & 1u) ? cntR : cntF, lane);
                    }
                    p0 = np0; p1 = np1; hashed = nhashed;
                }
            }
        }
        __syncthreads();

        postings_epilogue<CB, kPostThreads>(a, read, len, flag, cntF, cntR, s_red);
    }
}

// ------------------------------------------------------------------------------------------
// lookup for SHORT lists: LG lanes per list, 32 / LG lists per warp step
// ------------------------------------------------------------------------------------------
// The kernel above spends a warp-wide step on every list: right for the ~360 ids per list of a human-sized filter, wasteful
// for a 4 000-bin one (46 ids = 6 units: most lanes idle, ~40 warp instructions and a dependent pair of requests per list,
// four lists in flight per warp -- 13.6 M chunks/s with the DRAM a third busy, profiles/r2_x_*).  Here a group of LG = 2, 4 or 8
// lanes owns a list: lane s loads unit s (and s + LG, ... of a longer list), 16, 8 or 4 lists per warp step, four steps' bounds
// and then four steps' units requested before the first counter is touched (64 .. 16 lists in flight per warp in 16
// registers).  The order of the ids inside a list does not matter (counting is commutative), so the same table serves both
// kernels; the launcher picks by the mean list length of the table.
// One increment unless the id is a pad (>= sentinel), as ONE predicated shared-memory reduction: written as `if (id <
// sentinel) bump(...)` the compiler wraps every increment into a divergence region (BSSY / BRA / BSYNC were 28 % of the
// instructions of the first version, profiles/r2_aj_slots_sub_w32_ncu.json).  cnt_s = shared-window address of the counters.
template <int CB>
__device__ __forceinline__ void bump_if_real(const uint32_t cnt_s, const uint32_t v, const uint32_t sentinel)
{
    const uint32_t id = v & 0xFFFFu;
    const uint32_t addr = cnt_s + (CB == 8 ? (v & 0xFFFCu) : ((v & 0xFFFEu) << 1));
    const uint32_t inc = __funnelshift_l(0u, 1u, v << (CB == 8 ? 3 : 4));
    asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %2, %3;\n\t@p red.shared.add.u32 [%0], %1;\n\t}"
                 :: "r"(addr), "r"(inc), "r"(id), "r"(sentinel) : "memory");
}

template <int CB>
__device__ __forceinline__ void add_ids_real(const uint32_t cnt_s, const uint4 v, const uint32_t sentinel)
{
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        bump_if_real<CB>(cnt_s, w[i], sentinel);
        bump_if_real<CB>(cnt_s, w[i] >> 16, sentinel);
    }
}

template <int CB, int kPostThreads, int LG>
__global__ void __launch_bounds__(kPostThreads, 768 / kPostThreads)
count_postings_sub_kernel(const CountArgs a, const uint32_t *__restrict__ ptr, const uint4 *__restrict__ ids, const uint32_t cnt_words)
{
    constexpr int kPostWarps = kPostThreads / 32;
    constexpr int NL = 32 / LG;                                          // lists per warp step
    constexpr int U = 4;                                                 // steps in flight
    extern __shared__ __align__(16) uint32_t s_mem[];
    uint32_t *const cntF = s_mem, *const cntR = s_mem + cnt_words;
    uint32_t *const s_x = s_mem + 2 * cnt_words;                         // [kPostPiece] packed k-mer or ~0u (not ACGT)
    uint8_t *const s_dig = reinterpret_cast<uint8_t *>(s_x + kPostPiece); // [kPostPiece + 32] Dna5 ranks
    __shared__ uint32_t s_red[kPostWarps];
    __shared__ FilterView s_fv;             // for the out-of-line hashed path: one copy per CTA, no per-thread stack frame
    if (threadIdx.x == 0) s_fv = a.fv;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t grp = (uint32_t)lane / LG, sub = (uint32_t)lane % LG;
    const uint32_t k = a.fv.hp.k;
    const uint32_t kbits = 2 * k;
    const uint32_t kmask = kbits >= 32 ? ~0u : ((1u << kbits) - 1u);
    const uint32_t sentinel = postings_sentinel(a.fv.n_bins_local);
    // pairs alternate strands and every step starts even: a lane always counts into the same strand's counters
    const uint32_t cnt_s = (uint32_t)__cvta_generic_to_shared((grp & 1u) ? cntR : cntF);

    for (uint32_t w = tid; w < 2 * cnt_words; w += kPostThreads) s_mem[w] = 0;

    for (uint64_t read = blockIdx.x; read < a.n_reads; read += gridDim.x) {
        const uint64_t off = a.read_off[read];
        const uint64_t len = a.read_off[read + 1] - off;
        uint32_t flag = read_flag_of(len, k);
        if (flag == 0 && CB == 8 && len - k + 1 > 255) flag = 3;         // longer than the caller's max_read_len promised
        if (tid == 0 && a.read_flag) a.read_flag[read] = (uint8_t)flag;
        __syncthreads();                                                   // counters are zero and visible

        if (flag == 0) {
            const uint32_t npos = (uint32_t)len - k + 1;
            for (uint32_t cs = 0; cs < npos; cs += kPostPiece) {
                const uint32_t cn = min((uint32_t)kPostPiece, npos - cs);
                __syncthreads();
                for (uint32_t i = tid; i < cn + k - 1; i += kPostThreads) s_dig[i] = (uint8_t)dna5(a.bases[off + cs + i]);
                __syncthreads();
                for (uint32_t j = tid; j < cn; j += kPostThreads) {
                    uint32_t x = 0, bad = 0;
                    for (uint32_t u = 0; u < k; ++u) {
                        const uint32_t d = s_dig[j + u];
                        x = (x << 2) | (d & 3u);
                        bad |= d >> 2;
                    }
                    s_x[j] = bad ? ~0u : (x & kmask);
                }
                __syncthreads();
                // every warp takes an equal, contiguous, even share of the (position, strand) pairs
                const uint32_t n_pairs = 2 * cn;
                const uint32_t share = 2u * ((cn + kPostWarps - 1) / kPostWarps);
                const uint32_t q_end = min(n_pairs, (warp + 1) * share);
                for (uint32_t qb = warp * share; qb < q_end; qb += NL * U) {
                    uint32_t p0[U], n_u[U], hashed = 0;
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const uint32_t q = qb + NL * u + grp;
                        p0[u] = 0; n_u[u] = 0;
                        if (q < q_end) {
                            const uint32_t x = s_x[q >> 1];
                            if (x == ~0u) hashed |= 1u << u;
                            else {
                                uint32_t idx = x;
                                if (q & 1u) {                              // reverse strand: the list of revcomp(x)
                                    uint32_t v = __brev(~x);
                                    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
                                    idx = v >> (32 - kbits);
                                }
                                p0[u] = __ldg(ptr + idx);
                                n_u[u] = __ldg(ptr + idx + 1) - p0[u];
                            }
                        }
                    }
                    uint4 v[U];
                    const uint32_t pads = sentinel | (sentinel << 16);
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        v[u] = make_uint4(pads, pads, pads, pads);         // lanes past the end of the list count nothing
                        if (sub < n_u[u]) v[u] = __ldg(ids + p0[u] + sub);
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        add_ids_real<CB>(cnt_s, v[u], sentinel);
                        // lists of more than LG units: the group walks on.  (Loading a lane's second unit up front as well was
                        // measured 10-25 % slower at every list length, profiles/r2_ab_postings_sub_sweep.jsonl.)
                        for (uint32_t o = LG; o < n_u[u]; o += LG)
                            if (o + sub < n_u[u]) add_ids_real<CB>(cnt_s, __ldg(ids + p0[u] + o + sub), sentinel);
                    }
                    // windows with a non-ACGT base: the whole warp evaluates the rows (rare)
                    if (__any_sync(0xffffffffu, hashed != 0))
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const uint32_t hm = __ballot_sync(0xffffffffu, (hashed >> u) & 1u);
                        for (uint32_t g = 0; g < (uint32_t)NL; ++g)
                            if ((hm >> (g * LG)) & 1u) {
                                const uint32_t q = qb + NL * u + g;
                                add_hashed<CB, 1>(s_fv, s_dig + (q >> 1), q & 1u, (q & 1u) ? cntR : cntF, lane);
                            }
                    }
                }
            }
        }
        __syncthreads();
        postings_epilogue<CB, kPostThreads>(a, read, len, flag, cntF, cntR, s_red);
    }
}

// ==========================================================================================================
// SLOT layout: every k-mer owns a fixed, 128-byte-aligned slot of slot_bytes (ibf_postings_layout.cuh)
// ==========================================================================================================
// Why: the list layout above costs a dependent pointer fetch (a whole 128-byte line for 8 useful bytes) before every list
// and lists start at random 16-byte offsets (a 720-byte list touches 6.5 lines): ncu measured 469 KB per chunk for 343 KB
// of ids, and 27 % of the warps' time waiting for list data with the register file capping the lists in flight
// (profiles/r1_m_postings_cfg3_ncu_full.json).  With a slot per k-mer the address is a multiplication, a list is ONE
// aligned bulk copy (cp.async.bulk global -> shared, completion on an mbarrier), and the lists in flight live in a
// shared-memory ring per warp instead of registers.  Lists longer than a slot go to an overflow area (the slot then
// holds its offset); the slot size is chosen from the sampled list-length distribution so that this is rare.

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// one bulk copy global -> shared (async proxy); bytes is a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Waits for the phase with the given parity; gives up after ~1 s of polling (a copy that never lands is a bug, not a
// reason to hang the device) and reports through *err.
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity, unsigned int *err)
{
    const uint32_t addr = smem_u32(bar);
    uint32_t ok = 0;
    for (uint32_t spin = 0; !ok; ++spin) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok)
                     : "r"(addr), "r"(parity)
                     : "memory");
        if (!ok && spin > (1u << 22)) { if (err) atomicExch(err, 1u); return false; }
    }
    return true;
}

// ---- build ------------------------------------------------------------------------------------------------
struct SlotTable {
    uint8_t *slots;            // [4^k][slot_bytes]
    uint16_t *ovf;             // overflow lists, 16-byte units of 8 ids, ascending, padded with the plain sentinel
    uint32_t slot_bytes;
    uint32_t ovf_cap_units;
    unsigned int *ovf_used;    // units handed out (device counter)
    unsigned int *err;         // set when the overflow area is exhausted
};

__global__ void __launch_bounds__(256) slots_fill_kernel(const FilterView fv, const uint64_t n_kmers, const SlotTable tb, const int deal)
{
    __shared__ uint16_t s_stage[8][kFillStage];
    __shared__ uint32_t s_off[8][32];
    const HashParams &hp = fv.hp;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t sentinel = postings_sentinel(fv.n_bins_local);
    const uint32_t cap = slot_capacity(tb.slot_bytes);
    uint16_t *const stage = s_stage[wib];
    uint32_t *const off = s_off[wib];
    for (uint64_t x = warp0; x < n_kmers; x += n_warps) {
        uint64_t Hf, Hr;
        kmer_hashes(x, hp.k, Hf, Hr);
        const uint64_t *rows[kMaxHash];
#pragma unroll
        for (int h = 0; h < kMaxHash; ++h)
            rows[h] = (uint32_t)h < hp.n_hash ? fv.words + hash_row(Hf, hp.pre[h], hp.n_blocks, hp.magic) * fv.stride : nullptr;
        // ascending ids into the stage (as many as it holds), n = all of them
        auto scan = [&](uint16_t *dst, uint64_t dst_cap) -> uint32_t {
            uint64_t run = 0;
            for (uint64_t w0 = 0; w0 < fv.stride; w0 += 32) {
                const uint64_t w = w0 + lane;
                uint64_t m = 0;
                if (w < fv.stride) {
                    m = ~0ULL;
#pragma unroll
                    for (int h = 0; h < kMaxHash; ++h)
                        if ((uint32_t)h < hp.n_hash) m &= __ldg(rows[h] + w);
                }
                const uint32_t c = (uint32_t)__popcll(m);
                uint32_t incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                uint64_t pos = run + incl - c;
                while (m) {
                    const int b = __ffsll((long long)m) - 1;
                    m &= m - 1;
                    if (pos < dst_cap) dst[pos] = (uint16_t)(w * 64 + b);
                    ++pos;
                }
                run += __shfl_sync(0xffffffffu, incl, 31);
            }
            return (uint32_t)run;
        };
        const uint32_t n = scan(stage, kFillStage);
        __syncwarp();
        uint8_t *const slot = tb.slots + x * tb.slot_bytes;
        uint16_t *const out = reinterpret_cast<uint16_t *>(slot + kSlotHeaderBytes);
        if (n > cap) {
            // ---- overflow: the whole list, ascending, in units of 8 ids ---------------------------------------
            const uint32_t units = (n + 7u) >> 3;
            uint32_t u0 = 0;
            if (lane == 0) u0 = atomicAdd(tb.ovf_used, units);
            u0 = __shfl_sync(0xffffffffu, u0, 0);
            const bool fits = (uint64_t)u0 + units <= tb.ovf_cap_units;
            if (lane == 0) {
                if (!fits) atomicExch(tb.err, 1u);
                reinterpret_cast<uint32_t *>(slot)[0] = fits ? kSlotOverflow : 0u;   // not fits: an empty list (the build is void anyway)
                reinterpret_cast<uint32_t *>(slot)[1] = u0;
                reinterpret_cast<uint32_t *>(slot)[2] = units;
            }
            if (fits) {
                uint16_t *const dst = tb.ovf + (uint64_t)u0 * 8;
                if (n <= (uint32_t)kFillStage) for (uint32_t i = lane; i < n; i += 32) dst[i] = stage[i];
                else scan(dst, n);
                for (uint32_t i = n + lane; i < 8u * units; i += 32) dst[i] = (uint16_t)sentinel;
            }
            __syncwarp();
            continue;
        }
        if (lane == 0) { reinterpret_cast<uint32_t *>(slot)[0] = n; reinterpret_cast<uint32_t *>(slot)[1] = 0u; }
        // pads up to the end of the last round the lookup walks (never past the slot)
        const uint32_t lim = min(cap, (n + 127u) & ~127u);
        for (uint32_t p = n + lane; p < lim; p += 32) out[p] = (uint16_t)slot_pad_id(sentinel, p);
        if (!deal) {
            for (uint32_t i = lane; i < n; i += 32) out[i] = stage[i];
            __syncwarp();
            continue;
        }
        // ---- deal the ascending ids over the groups of the slot, bank by bank ------------------------------------
        off[lane] = 0;
        __syncwarp();
        for (uint32_t i = lane; i < n; i += 32) atomicAdd(&off[counter_bank(stage[i])], 1u);
        __syncwarp();
        {
            const uint32_t load = off[lane];
            uint32_t incl = load;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            __syncwarp();
            off[lane] = incl - load;
        }
        __syncwarp();
        const SlotShape shape = slot_shape(n);
        for (uint32_t i0 = 0; i0 < n; i0 += 32) {
            const uint32_t i = i0 + lane;
            const bool valid = i < n;
            const uint32_t id = valid ? stage[i] : 0u;
            const uint32_t b = counter_bank(id);
            const uint32_t same = __match_any_sync(0xffffffffu, valid ? b : 32u + lane);
            const uint32_t before = __popc(same & ((1u << lane) - 1u));
            const uint32_t c = valid ? off[b] + before : 0u;
            __syncwarp();
            if (valid && before == 0) off[b] += __popc(same);
            __syncwarp();
            if (valid) out[slot_position(shape, c)] = (uint16_t)id;
        }
        __syncwarp();
    }
}

// ---- lookup -------------------------------------------------------------------------------------------------
template <int CB>
__device__ __forceinline__ void bump_checked(uint32_t *cnt, const uint32_t v, const uint32_t sentinel)
{
    if ((v & 0xFFFFu) != sentinel) bump<CB>(cnt, v);
}

// One slot, landed in shared memory: its ids into the counters of one strand.  Round r: lane l takes the four ids at
// positions 128 r + 4 l (one LDS.64); pads fill the last round, so only a slot-filling list ends in a partial one.
template <int CB>
__device__ __forceinline__ void consume_slot(const uint8_t *sp, uint32_t *cnt, const uint32_t cap, const int lane,
                                             const uint4 *__restrict__ ovf, const uint32_t sentinel)
{
    const uint2 hdr = *reinterpret_cast<const uint2 *>(sp);
    const uint32_t n = hdr.x & 0xFFFFu;
    if (n != kSlotOverflow) {
        const uint32_t lim = min(cap, (n + 127u) & ~127u);
        const uint8_t *ip = sp + kSlotHeaderBytes + 8u * lane;
        const uint32_t full = lim >> 7;
#pragma unroll 2
        for (uint32_t r = 0; r < full; ++r, ip += 256) {
            const uint2 v = *reinterpret_cast<const uint2 *>(ip);
            bump_pair<CB>(cnt, v.x);
            bump_pair<CB>(cnt, v.y);
        }
        if (4u * lane < (lim & 127u)) {
            const uint2 v = *reinterpret_cast<const uint2 *>(ip);
            bump_pair<CB>(cnt, v.x);
            bump_pair<CB>(cnt, v.y);
        }
    } else {                                                               // rare: the list lives in the overflow area
        const uint32_t units = *reinterpret_cast<const uint32_t *>(sp + 8);
        for (uint32_t u = lane; u < units; u += 32) {
            const uint4 v = __ldg(ovf + hdr.y + u);
            bump_checked<CB>(cnt, v.x, sentinel); bump_checked<CB>(cnt, v.x >> 16, sentinel);
            bump_checked<CB>(cnt, v.y, sentinel); bump_checked<CB>(cnt, v.y >> 16, sentinel);
            bump_checked<CB>(cnt, v.z, sentinel); bump_checked<CB>(cnt, v.z >> 16, sentinel);
            bump_checked<CB>(cnt, v.w, sentinel); bump_checked<CB>(cnt, v.w >> 16, sentinel);
        }
    }
}

// CB: counter bits.  THREADS per CTA: 256, 384 or 768 by how many CTAs of counters + rings fit an SM.
// Dynamic shared memory: [2 * cnt_words counters][2 * kPostPiece slot indices][digits][pad to 128]
//                        [warps * ring entries of 2 slots (forward, reverse strand of one position)][warps * ring mbarriers]
template <int CB, int THREADS>
__global__ void __launch_bounds__(THREADS, 768 / THREADS)
count_slots_kernel(const CountArgs a, const uint8_t *__restrict__ slots, const uint32_t slot_bytes, const uint4 *__restrict__ ovf,
                   const uint32_t cnt_words, const uint32_t ring, const uint32_t ring_off, unsigned int *err)
{
    constexpr int PER = 32 / CB;
    constexpr uint32_t CMASK = (CB == 8) ? 0xFFu : 0xFFFFu;
    constexpr int kWarps = THREADS / 32;
    extern __shared__ __align__(128) uint32_t s_slot_mem[];
    uint32_t *const s_mem = s_slot_mem;
    uint32_t *const cntF = s_mem, *const cntR = s_mem + cnt_words;
    uint32_t *const s_idx = s_mem + 2 * cnt_words;                         // [2 * kPostPiece]: slot index of (position, strand) or ~0u
    uint8_t *const s_dig = reinterpret_cast<uint8_t *>(s_idx + 2 * kPostPiece);
    __shared__ uint32_t s_red[kWarps];

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);                // the same value, provably warp-uniform
    const uint32_t entry_bytes = 2 * slot_bytes;
    uint8_t *const my_ring = reinterpret_cast<uint8_t *>(s_mem) + ring_off + (size_t)warp * ring * entry_bytes;
    uint64_t *const my_bar = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(s_mem) + ring_off + (size_t)kWarps * ring * entry_bytes) + warp * ring;
    const uint32_t k = a.fv.hp.k;
    const uint32_t kbits = 2 * k;
    const uint32_t kmask = kbits >= 32 ? ~0u : ((1u << kbits) - 1u);
    const uint64_t nbl = a.fv.n_bins_local;
    const uint32_t sentinel = postings_sentinel(nbl);
    const uint32_t cap = slot_capacity(slot_bytes);

    for (uint32_t w = tid; w < 2 * cnt_words; w += THREADS) s_mem[w] = 0;
    if (lane == 0)
        for (uint32_t j = 0; j < ring; ++j) mbar_init(my_bar + j, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint32_t phase_bits = 0;                                               // expected parity of each of this warp's ring entries
    bool dead = false;                                                     // a copy never landed: stop waiting, wind down

    for (uint64_t read = blockIdx.x; read < a.n_reads; read += gridDim.x) {
        if (*reinterpret_cast<volatile unsigned int *>(err)) break;        // some warp of the grid gave up (reported by the host call)
        const uint64_t off = a.read_off[read];
        const uint64_t len = a.read_off[read + 1] - off;
        uint32_t flag = read_flag_of(len, k);
        if (flag == 0 && CB == 8 && len - k + 1 > 255) flag = 3;         // longer than the caller's max_read_len promised
        if (tid == 0 && a.read_flag) a.read_flag[read] = (uint8_t)flag;
        __syncthreads();                                                   // counters are zero and visible

        if (flag == 0) {
            const uint32_t npos = (uint32_t)len - k + 1;
            for (uint32_t cs = 0; cs < npos; cs += kPostPiece) {
                const uint32_t cn = min((uint32_t)kPostPiece, npos - cs);
                __syncthreads();
                for (uint32_t i = tid; i < cn + k - 1; i += THREADS) s_dig[i] = (uint8_t)dna5(a.bases[off + cs + i]);
                __syncthreads();
                for (uint32_t j = tid; j < cn; j += THREADS) {
                    uint32_t x = 0, bad = 0;
                    for (uint32_t u = 0; u < k; ++u) {
                        const uint32_t d = s_dig[j + u];
                        x = (x << 2) | (d & 3u);
                        bad |= d >> 2;
                    }
                    x &= kmask;
                    uint32_t v = __brev(~x);                               // reverse strand: the slot of revcomp(x)
                    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
                    *reinterpret_cast<uint2 *>(s_idx + 2 * j) = bad ? make_uint2(~0u, ~0u) : make_uint2(x, v >> (32 - kbits));
                }
                __syncthreads();
                // every warp takes an equal, contiguous share of the positions and keeps `ring` of them in flight: both strands'
                // slots of a position land in one ring entry (two bulk copies, one mbarrier); the copies of position p + ring
                // are issued when position p has been counted
                const uint32_t share = (cn + kWarps - 1) / kWarps;
                const uint32_t p_begin = min(cn, warp * share), p_end = min(cn, p_begin + share);
                auto issue = [&](const uint32_t p, const uint32_t e) {
                    const uint2 ix = *reinterpret_cast<const uint2 *>(s_idx + 2 * p);
                    if (ix.x == ~0u) return;                               // window with a non-ACGT base: hashed, the entry stays idle
                    if (lane == 0) {
                        uint8_t *const dst = my_ring + (size_t)e * entry_bytes;
                        mbar_expect_tx(my_bar + e, entry_bytes);
                        bulk_g2s(dst, slots + (uint64_t)ix.x * slot_bytes, slot_bytes, my_bar + e);
                        bulk_g2s(dst + slot_bytes, slots + (uint64_t)ix.y * slot_bytes, slot_bytes, my_bar + e);
                    }
                };
                for (uint32_t j = 0; j < ring && p_begin + j < p_end; ++j) issue(p_begin + j, j);
                uint32_t e = 0;
#pragma unroll 1
                for (uint32_t p = p_begin; p < p_end; ++p) {
                    if (s_idx[2 * p] == ~0u) {
                        add_hashed<CB>(a.fv, s_dig + p, 0u, cntF, lane);
                        add_hashed<CB>(a.fv, s_dig + p, 1u, cntR, lane);
                    } else if (!dead) {
                        if (mbar_wait(my_bar + e, (phase_bits >> e) & 1u, err)) {
                            phase_bits ^= 1u << e;
                            const uint8_t *const sp = my_ring + (size_t)e * entry_bytes;
                            consume_slot<CB>(sp, cntF, cap, lane, ovf, sentinel);
                            consume_slot<CB>(sp + slot_bytes, cntR, cap, lane, ovf, sentinel);
                        } else
                            dead = true;
                    }
                    // the entry's ids are counted (every lane is past its loads): its next copies may land
                    __syncwarp();
                    if (p + ring < p_end) issue(p + ring, e);
                    if (++e == ring) e = 0;
                }
            }
        }
        __syncthreads();

        // ---- epilogue: M = max over bins of max(fwd, rev), its lowest bin; counters back to zero ----------------
        // The pads' counters (32 words for 8-bit, 64 for 16-bit counters) lie behind the quads of words that hold bins: they
        // are not scanned, only cleared.
        const uint32_t sent_word = sentinel / PER;
        const uint32_t scan_quads = ((uint32_t)nbl + 4u * PER - 1u) / (4u * PER);
        for (uint32_t w = tid; w < 128u / PER; w += THREADS) { cntF[sent_word + w] = 0; cntR[sent_word + w] = 0; }
        if (a.counts_fwd || a.counts_rev) {                                 // dense counts on request (tests, tools)
            for (uint32_t w = tid; w * PER < nbl; w += THREADS) {
                const uint32_t f = cntF[w], r = cntR[w];
#pragma unroll
                for (int i = 0; i < PER; ++i) {
                    const uint32_t bin = w * PER + i;
                    if (bin < nbl) {
                        if (a.counts_fwd) a.counts_fwd[read * nbl + bin] = (uint16_t)((f >> (i * CB)) & CMASK);
                        if (a.counts_rev) a.counts_rev[read * nbl + bin] = (uint16_t)((r >> (i * CB)) & CMASK);
                    }
                }
            }
            __syncthreads();
        }
        constexpr uint32_t LOW = CB == 8 ? 0x7F7F7F7Fu : 0x7FFF7FFFu, ONES = CB == 8 ? 0x01010101u : 0x00010001u;
        constexpr uint32_t HALF = (CMASK + 1u) / 2u;
        uint32_t thr_min = CMASK + 1u;
        if (flag == 0)
            for (uint32_t t = 0; t < a.n_lut; ++t) thr_min = min(thr_min, (uint32_t)__ldg(a.lut + (size_t)t * kLutSize + len));
        uint32_t cur = min(thr_min, CMASK + 1u) - (thr_min ? 1u : 0u), cur_bin = 0;
        uint32_t add = (HALF - 1u - min(cur, HALF - 1u)) * ONES;
        uint4 *const qF = reinterpret_cast<uint4 *>(cntF), *const qR = reinterpret_cast<uint4 *>(cntR);
        for (uint32_t qd = tid; qd < scan_quads; qd += THREADS) {
            const uint4 f = qF[qd], r = qR[qd];
            qF[qd] = make_uint4(0, 0, 0, 0); qR[qd] = make_uint4(0, 0, 0, 0);
            const uint32_t gx = f.x | r.x, gy = f.y | r.y, gz = f.z | r.z, gw = f.w | r.w;
            const uint32_t h = (((gx & LOW) + add) | gx) | (((gy & LOW) + add) | gy) | (((gz & LOW) + add) | gz) | (((gw & LOW) + add) | gw);
            if (h & ~LOW) {
                const uint32_t m[4] = {vmax<CB>(f.x, r.x), vmax<CB>(f.y, r.y), vmax<CB>(f.z, r.z), vmax<CB>(f.w, r.w)};
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int i = 0; i < PER; ++i) {
                        const uint32_t v = (m[j] >> (i * CB)) & CMASK;
                        if (v > cur) { cur = v; cur_bin = (4u * qd + j) * PER + i; }
                    }
                add = (HALF - 1u - min(cur, HALF - 1u)) * ONES;
            }
        }
        uint32_t best_key = __reduce_max_sync(0xffffffffu, (cur << 16) | (0xFFFFu - cur_bin));
        if (lane == 0) s_red[warp] = best_key;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kWarps; ++i) best_key = max(best_key, s_red[i]);
        const uint32_t M = best_key >> 16, best_bin = 0xFFFFu - (best_key & 0xFFFFu);
        if (tid < (int)a.n_lut) {
            uint64_t key = 0;
            if (flag == 0) {
                const uint32_t thr = (uint32_t)__ldg(a.lut + (size_t)tid * kLutSize + len);
                if (M >= thr) key = pack_key(M, (uint32_t)(a.fv.bin_begin + best_bin));
            }
            uint64_t *const dst = a.keys + (size_t)tid * a.n_reads + read;
            if (a.keys_shared) { if (key) key_max(dst, key, 1); }          // all bin shards fold into one array (NVLink peer atomics)
            else *dst = key;
        }
    }
}

// counters: whole quads of bins + the 128 pad ids' words (+ slack to a quad)
// ------------------------------------------------------------------------------------------
// lookup in slots of 128 / 256 bytes: 8 / 16 lanes load one slot with ONE instruction
// ------------------------------------------------------------------------------------------
// Short lists again (count_postings_sub_kernel), without the pointer: the slot of a k-mer is found by a multiplication, so
// a list costs ONE request of one or two lines instead of a dependent pair (a whole line for the 8 bytes of bounds, then
// 1.5+ lines for a list that starts at a random 16-byte offset).  Lane s of a group takes piece s of the slot; the first 8
// bytes are the header (n, or kSlotOverflow + the place of a list that did not fit), ids >= sentinel are pads.  Counters and
// epilogue are those of the list kernels.
template <int CB, int kPostThreads, int LG>
__global__ void __launch_bounds__(kPostThreads, 768 / kPostThreads)
count_slots_sub_kernel(const CountArgs a, const uint8_t *__restrict__ slots, const uint4 *__restrict__ ovf, const uint32_t cnt_words)
{
    constexpr int kPostWarps = kPostThreads / 32;
    constexpr int NL = 32 / LG;                                          // slots per warp step
    constexpr int U = 4;                                                 // steps in flight
    constexpr uint32_t kSlotB = 16u * LG;
    extern __shared__ __align__(16) uint32_t s_mem[];
    uint32_t *const cntF = s_mem, *const cntR = s_mem + cnt_words;
    uint32_t *const s_x = s_mem + 2 * cnt_words;                         // [kPostPiece] packed k-mer or ~0u (not ACGT)
    uint8_t *const s_dig = reinterpret_cast<uint8_t *>(s_x + kPostPiece); // [kPostPiece + 32] Dna5 ranks
    __shared__ uint32_t s_red[kPostWarps];
    // (the shared-memory filter view of the list kernel above made this one 17 % slower: measured, profiles/r2_as_*)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t grp = (uint32_t)lane / LG, sub = (uint32_t)lane % LG;
    const uint32_t k = a.fv.hp.k;
    const uint32_t kbits = 2 * k;
    const uint32_t kmask = kbits >= 32 ? ~0u : ((1u << kbits) - 1u);
    const uint32_t sentinel = postings_sentinel(a.fv.n_bins_local);
    // pairs alternate strands and every step starts even: a lane always counts into the same strand's counters
    const uint32_t cnt_s = (uint32_t)__cvta_generic_to_shared((grp & 1u) ? cntR : cntF);

    for (uint32_t w = tid; w < 2 * cnt_words; w += kPostThreads) s_mem[w] = 0;

    for (uint64_t read = blockIdx.x; read < a.n_reads; read += gridDim.x) {
        const uint64_t off = a.read_off[read];
        const uint64_t len = a.read_off[read + 1] - off;
        uint32_t flag = read_flag_of(len, k);
        if (flag == 0 && CB == 8 && len - k + 1 > 255) flag = 3;         // longer than the caller's max_read_len promised
        if (tid == 0 && a.read_flag) a.read_flag[read] = (uint8_t)flag;
        __syncthreads();                                                   // counters are zero and visible

        if (flag == 0) {
            const uint32_t npos = (uint32_t)len - k + 1;
            for (uint32_t cs = 0; cs < npos; cs += kPostPiece) {
                const uint32_t cn = min((uint32_t)kPostPiece, npos - cs);
                __syncthreads();
                for (uint32_t i = tid; i < cn + k - 1; i += kPostThreads) s_dig[i] = (uint8_t)dna5(a.bases[off + cs + i]);
                __syncthreads();
                for (uint32_t j = tid; j < cn; j += kPostThreads) {
                    uint32_t x = 0, bad = 0;
                    for (uint32_t u = 0; u < k; ++u) {
                        const uint32_t d = s_dig[j + u];
                        x = (x << 2) | (d & 3u);
                        bad |= d >> 2;
                    }
                    s_x[j] = bad ? ~0u : (x & kmask);
                }
                __syncthreads();
                const uint32_t n_pairs = 2 * cn;
                const uint32_t share = 2u * ((cn + kPostWarps - 1) / kPostWarps);
                const uint32_t q_end = min(n_pairs, (warp + 1) * share);
                for (uint32_t qb = warp * share; qb < q_end; qb += NL * U) {
                    uint4 v[U];
                    uint32_t hashed = 0;
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const uint32_t q = qb + NL * u + grp;
                        v[u] = make_uint4(0, 0, 0, 0);                     // header n = 0: nothing to count
                        if (q < q_end) {
                            const uint32_t x = s_x[q >> 1];
                            if (x == ~0u) hashed |= 1u << u;
                            else {
                                uint32_t idx = x;
                                if (q & 1u) {                              // reverse strand: the slot of revcomp(x)
                                    uint32_t r = __brev(~x);
                                    r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
                                    idx = r >> (32 - kbits);
                                }
                                v[u] = __ldg(reinterpret_cast<const uint4 *>(slots + (uint64_t)idx * kSlotB) + sub);
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        // header words of the group's slot (piece 0)
                        const uint32_t h0 = __shfl_sync(0xffffffffu, v[u].x, grp * LG), h1 = __shfl_sync(0xffffffffu, v[u].y, grp * LG);
                        const uint32_t h2 = __shfl_sync(0xffffffffu, v[u].z, grp * LG);
                        const uint32_t n = h0 & 0xFFFFu;
                        if (n == kSlotOverflow) {                          // the list lives in the overflow area: h1 = first unit, h2 = units
                            for (uint32_t o = sub; o < h2; o += LG) add_ids_real<CB>(cnt_s, __ldg(ovf + h1 + o), sentinel);
                        } else if (n != 0) {
                            uint4 w = v[u];
                            if (sub == 0) { w.x = sentinel | (sentinel << 16); w.y = w.x; }     // the header is not ids
                            add_ids_real<CB>(cnt_s, w, sentinel);
                        }
                    }
                    // windows with a non-ACGT base: the whole warp evaluates the rows (rare)
                    if (__any_sync(0xffffffffu, hashed != 0))
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const uint32_t hm = __ballot_sync(0xffffffffu, (hashed >> u) & 1u);
                        for (uint32_t g = 0; g < (uint32_t)NL; ++g)
                            if ((hm >> (g * LG)) & 1u) {
                                const uint32_t q = qb + NL * u + g;
                                add_hashed<CB>(a.fv, s_dig + (q >> 1), q & 1u, (q & 1u) ? cntR : cntF, lane);
                            }
                    }
                }
            }
        }
        __syncthreads();
        postings_epilogue<CB, kPostThreads>(a, read, len, flag, cntF, cntR, s_red);
    }
}

size_t slots_counter_words(uint64_t n_bins_local, int counter_bits)
{
    const uint32_t per = 32 / counter_bits;
    return (size_t)postings_sentinel(n_bins_local) / per + 128u / per + 4;
}

size_t postings_smem_bytes(uint64_t n_bins_local, int counter_bits, uint32_t *cnt_words)
{
    const uint32_t per = 32 / counter_bits;
    const uint32_t words = postings_sentinel(n_bins_local) / per + 4;     // whole quads of bins + the sentinel's quad
    if (cnt_words) *cnt_words = words;
    return (size_t)2 * words * 4 + kPostPiece * 4 + kPostPiece + 32;
}

}  // namespace

// Applicable: wide row, k-mer index fits 32 bits, bin ids fit 16 bits with the sentinel, 8-bit counters fit shared memory.
bool postings_applicable(const FilterView &fv)
{
    return fv.stride > 4 && fv.hp.k <= 15 && fv.n_bins_local <= 65520 && fv.hp.n_blocks > 0 &&
           postings_smem_bytes(fv.n_bins_local, 8, nullptr) <= 200u * 1024u;
}

// Average list length in 16-byte units over a sample of k-mers (size estimate before committing to the build).
// d_scratch: at least n_sample uint32.  Synchronises the stream.
int postings_sample_units(const FilterView &fv, uint32_t *d_scratch, uint32_t n_sample, double *mean_units, int sm_count,
                          cudaStream_t st)
{
    const uint64_t n_kmers = 1ull << (2 * fv.hp.k);
    // an odd step visits k-mers of every suffix (a power-of-two stride would fix the low bases of all samples)
    uint64_t step = n_kmers / n_sample ? n_kmers / n_sample : 1;
    if (step > 2) step |= 1;
    postings_count_kernel<<<sm_count * 8, 256, 0, st>>>(fv, step / 2, step, n_sample, d_scratch);
    std::vector<uint32_t> h(n_sample);
    if (cudaMemcpyAsync(h.data(), d_scratch, (size_t)n_sample * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
    double s = 0;
    for (uint32_t v : h) s += v;
    *mean_units = s / n_sample;
    return 1;
}

// exact 64-bit total of the list sizes (the scan below works in uint32: a table of 2^32 units or more must be refused, not wrapped)
__global__ void __launch_bounds__(256) sum_units_kernel(const uint32_t *__restrict__ units, uint64_t n, unsigned long long *total)
{
    unsigned long long t = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) t += units[i];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0 && t) atomicAdd(total, t);
}

// Pass 1 + scan: d_ptr[4^k + 1] (uint32 units).  Returns launches or < 0; *total_units = d_ptr[4^k] (host), stream synchronised.
int postings_build_ptr(const FilterView &fv, uint32_t *d_ptr, uint64_t *total_units, int sm_count, cudaStream_t st)
{
    const uint64_t n_kmers = 1ull << (2 * fv.hp.k);
    postings_count_kernel<<<sm_count * 16, 256, 0, st>>>(fv, 0, 1, n_kmers, d_ptr);
    if (cudaMemsetAsync(d_ptr + n_kmers, 0, 4, st) != cudaSuccess) return -1;
    // the table must stay below 2^32 units (64 GiB of ids): the caller's sampled estimate can be off, so count exactly
    {
        unsigned long long *d_total = nullptr, h_total = 0;
        if (cudaMallocAsync(&d_total, sizeof(h_total), st) != cudaSuccess) return -1;
        cudaMemsetAsync(d_total, 0, sizeof(h_total), st);
        sum_units_kernel<<<sm_count * 8, 256, 0, st>>>(d_ptr, n_kmers, d_total);
        cudaMemcpyAsync(&h_total, d_total, sizeof(h_total), cudaMemcpyDeviceToHost, st);
        cudaFreeAsync(d_total, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
        if (h_total >= 0xFFFFFFF0ull) return -1;
    }
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_ptr, d_ptr, n_kmers + 1, st);
    if (cudaMallocAsync(&d_tmp, tmp_bytes, st) != cudaSuccess) return -1;
    cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_ptr, d_ptr, n_kmers + 1, st);
    cudaFreeAsync(d_tmp, st);
    uint32_t last = 0;
    if (cudaMemcpyAsync(&last, d_ptr + n_kmers, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
    *total_units = last;
    return cudaGetLastError() == cudaSuccess ? 3 : -1;
}

int postings_fill(const FilterView &fv, const uint32_t *d_ptr, uint16_t *d_ids, int sm_count, cudaStream_t st)
{
    int deal = 1;                                               // RB_POSTINGS_ORDER=0: ascending ids (A/B measurements)
    if (const char *e = std::getenv("RB_POSTINGS_ORDER")) deal = e[0] != '0';
    postings_fill_kernel<<<sm_count * 16, 256, 0, st>>>(fv, 1ull << (2 * fv.hp.k), d_ptr, d_ids, deal);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// ---- slot table: host side ------------------------------------------------------------------------------------
bool slots_applicable(const FilterView &fv)
{
    return fv.stride > 4 && fv.hp.k <= 15 && fv.n_bins_local <= 65280 && fv.hp.n_blocks > 0 &&
           slots_counter_words(fv.n_bins_local, 8) * 8 + kPostPiece * 9 + 160 + 24 * (2 * 128 + 8) <= 200u * 1024u;
}

// List lengths (ids) of n_sample k-mers spread over all 4^k; d_scratch: n_sample uint32.  Synchronises the stream.
int slots_sample_lengths(const FilterView &fv, uint32_t *d_scratch, uint32_t n_sample, std::vector<uint32_t> *lengths, int sm_count,
                         cudaStream_t st)
{
    const uint64_t n_kmers = 1ull << (2 * fv.hp.k);
    // an odd step visits k-mers of every suffix (a power-of-two stride would fix the low bases of all samples)
    uint64_t step = n_kmers / n_sample ? n_kmers / n_sample : 1;
    if (step > 2) step |= 1;
    postings_count_kernel<<<sm_count * 8, 256, 0, st>>>(fv, step / 2, step, n_sample, d_scratch, 1);
    lengths->resize(n_sample);
    if (cudaMemcpyAsync(lengths->data(), d_scratch, (size_t)n_sample * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
    return 1;
}

// Slot size that minimises the expected bytes fetched per list -- slot + P(overflow) x (overflow lines + 1) -- among those whose
// table fits the budget; also the overflow area to reserve.  Returns false if nothing fits.
bool slots_choose(const std::vector<uint32_t> &lengths, uint32_t k, uint64_t budget, uint32_t *slot_bytes, uint64_t *ovf_units,
                  uint64_t *total_bytes)
{
    const double n_kmers = (double)(1ull << (2 * k));
    double best_cost = 0;
    bool found = false;
    for (uint32_t sb = 128; sb <= kSlotMaxBytes; sb += 128) {
        const uint32_t cap = slot_capacity(sb);
        double over = 0, over_units = 0, over_lines = 0;
        for (uint32_t n : lengths)
            if (n > cap) { over += 1; over_units += (n + 7) / 8; over_lines += ((n + 7) / 8 * 16 + 127) / 128 + 1; }
        const double p = over / (double)lengths.size();
        const double cost = sb + 128.0 * over_lines / (double)lengths.size();
        // reserve twice the sampled overflow (+ 1 M units): the sample is 65 536 k-mers
        const uint64_t units = (uint64_t)(2.0 * over_units / (double)lengths.size() * n_kmers) + (1u << 20);
        const uint64_t bytes = (uint64_t)(n_kmers * sb) + units * 16;
        if (bytes > budget || units > 0xFFFFFFF0ull) continue;
        (void)p;
        if (!found || cost < best_cost) { found = true; best_cost = cost; *slot_bytes = sb; *ovf_units = units; *total_bytes = bytes; }
    }
    return found;
}

// Fills the table; returns launches (1) or -1 (CUDA error) / -3 (overflow area exhausted: reserve more and retry).
int slots_fill(const FilterView &fv, uint8_t *d_slots, uint32_t slot_bytes, uint16_t *d_ovf, uint64_t ovf_units, unsigned int *d_ctr,
               int sm_count, cudaStream_t st)
{
    int deal = 1;
    if (const char *e = std::getenv("RB_POSTINGS_ORDER")) deal = e[0] != '0';
    if (cudaMemsetAsync(d_ctr, 0, 2 * sizeof(unsigned int), st) != cudaSuccess) return -1;
    SlotTable tb{};
    tb.slots = d_slots; tb.ovf = d_ovf; tb.slot_bytes = slot_bytes; tb.ovf_cap_units = (uint32_t)ovf_units;
    tb.ovf_used = d_ctr; tb.err = d_ctr + 1;
    slots_fill_kernel<<<sm_count * 16, 256, 0, st>>>(fv, 1ull << (2 * fv.hp.k), tb, deal);
    unsigned int h[2] = {0, 0};
    if (cudaMemcpyAsync(h, d_ctr, sizeof(h), cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
    if (cudaGetLastError() != cudaSuccess) return -1;
    return h[1] ? -3 : 1;
}

// returns launches, -1 on error, -2 if this launch's reads need more shared memory than an SM has
int launch_count_slots(const CountArgs &a, const uint8_t *d_slots, uint32_t slot_bytes, const uint16_t *d_ovf, uint32_t max_read_len,
                       int sm_count, unsigned int *d_err, cudaStream_t st)
{
    if (a.n_reads == 0) return 0;
    if (a.n_lut == 0 || a.n_lut > (uint32_t)kMaxLut) return -1;
    const uint32_t k = a.fv.hp.k;
    const bool narrow = max_read_len != 0 && (max_read_len < k || max_read_len - k + 1 <= 255);
    // slots of one or two lines: groups of 8 / 16 lanes load them straight into registers (RB_SLOTS_SUB=0: the ring kernel)
    if (slot_bytes <= 256) {
        const char *e = std::getenv("RB_SLOTS_SUB");
        if (!(e && e[0] == '0')) {
            uint32_t cw = 0;
            const size_t smem = postings_smem_bytes(a.fv.n_bins_local, narrow ? 8 : 16, &cw);
            if (smem > 220u * 1024u) return -2;
            const int fit = (int)std::min<size_t>(3, (227u * 1024u) / (smem + 1024u));
            const uint4 *ovf4 = reinterpret_cast<const uint4 *>(d_ovf);
            auto launch = [&](auto kernel, int threads) {
                cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                int occ = 1;
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem);
                if (occ < 1) occ = 1;
                const uint64_t capb = (uint64_t)sm_count * occ * grid_waves();
                const uint32_t gx = (uint32_t)(a.n_reads < capb ? a.n_reads : capb);
                kernel<<<gx, threads, smem, st>>>(a, d_slots, ovf4, cw);
            };
#define RB_SLOT_SUB_LAUNCH(LG)                                                                                                \
    do {                                                                                                                      \
        if (narrow) {                                                                                                         \
            if (fit >= 3) launch(count_slots_sub_kernel<8, 256, LG>, 256);                                                    \
            else if (fit == 2) launch(count_slots_sub_kernel<8, 384, LG>, 384);                                               \
            else launch(count_slots_sub_kernel<8, 768, LG>, 768);                                                             \
        } else {                                                                                                              \
            if (fit >= 3) launch(count_slots_sub_kernel<16, 256, LG>, 256);                                                   \
            else if (fit == 2) launch(count_slots_sub_kernel<16, 384, LG>, 384);                                              \
            else launch(count_slots_sub_kernel<16, 768, LG>, 768);                                                            \
        }                                                                                                                     \
    } while (0)
            if (slot_bytes == 128) RB_SLOT_SUB_LAUNCH(8); else RB_SLOT_SUB_LAUNCH(16);
#undef RB_SLOT_SUB_LAUNCH
            return cudaGetLastError() == cudaSuccess ? 1 : -1;
        }
    }
    const uint32_t cnt_words = (uint32_t)slots_counter_words(a.fv.n_bins_local, narrow ? 8 : 16);
    const size_t fixed = ((size_t)2 * cnt_words * 4 + kPostPiece * 8 + kPostPiece + 32 + 127) / 128 * 128;
    const size_t entry = 2 * (size_t)slot_bytes + 8;               // both strands' slots of one position + its mbarrier
    // CTAs per SM x threads: 3 x 256, 2 x 384 or 1 x 768 (always 24 warps per SM); take the shape that keeps the most
    // bytes in flight per SM (ring slots x 24 warps, at most 8 slots per warp), preferring more CTAs on a tie
    int ring_env = 0;
    if (const char *e = std::getenv("RB_SLOT_RING")) ring_env = std::atoi(e);
    int ctas_env = 0;
    if (const char *e = std::getenv("RB_SLOT_CTAS")) ctas_env = std::atoi(e);
    int best_ctas = 0, best_ring = 0;
    for (int ctas = 3; ctas >= 1; --ctas) {
        if (ctas_env && ctas != ctas_env) continue;
        const int warps = 24 / ctas;
        const size_t avail = (227u * 1024u) / ctas - 1024u;
        if (fixed + (size_t)warps * entry > avail) continue;
        int ring = (int)std::min<size_t>(8, (avail - fixed) / ((size_t)warps * entry));
        if (ring_env > 0) ring = std::min(ring, ring_env);
        if (ring > best_ring || (ring == best_ring && best_ctas == 0)) { best_ring = ring; best_ctas = ctas; }
        if (ring >= 2) break;                                     // two positions = four slots of a warp in flight cover the latency
    }
    if (!best_ctas) return -2;
    const int warps = 24 / best_ctas;
    const size_t smem = fixed + (size_t)warps * best_ring * entry;
    const uint4 *ovf = reinterpret_cast<const uint4 *>(d_ovf);
    auto launch = [&](auto kernel, int threads) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int occ = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem);
        if (occ < 1) occ = 1;
        const uint64_t capb = (uint64_t)sm_count * occ * grid_waves();
        const uint32_t gx = (uint32_t)(a.n_reads < capb ? a.n_reads : capb);
        kernel<<<gx, threads, smem, st>>>(a, d_slots, slot_bytes, ovf, cnt_words, (uint32_t)best_ring, (uint32_t)fixed, d_err);
    };
    if (narrow) {
        if (best_ctas == 3) launch(count_slots_kernel<8, 256>, 256);
        else if (best_ctas == 2) launch(count_slots_kernel<8, 384>, 384);
        else launch(count_slots_kernel<8, 768>, 768);
    } else {
        if (best_ctas == 3) launch(count_slots_kernel<16, 256>, 256);
        else if (best_ctas == 2) launch(count_slots_kernel<16, 384>, 384);
        else launch(count_slots_kernel<16, 768>, 768);
    }
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// counter width by the longest read of the launch; returns launches, -1 on error, -2 if this launch cannot use postings
int launch_count_postings(const CountArgs &a, const uint32_t *d_ptr, const uint16_t *d_ids, uint32_t max_read_len, double mean_units,
                          int sm_count, cudaStream_t st)
{
    if (a.n_reads == 0) return 0;
    if (a.n_lut == 0 || a.n_lut > (uint32_t)kMaxLut) return -1;
    const uint32_t k = a.fv.hp.k;
    const bool narrow = max_read_len != 0 && (max_read_len < k || max_read_len - k + 1 <= 255);
    uint32_t cnt_words = 0;
    const size_t smem = postings_smem_bytes(a.fv.n_bins_local, narrow ? 8 : 16, &cnt_words);
    if (smem > 220u * 1024u) return -2;
    const uint4 *ids = reinterpret_cast<const uint4 *>(d_ids);
    // CTAs of this much shared memory per SM (227 KB usable, 1 KB reserved per CTA) -> threads per CTA
    const int fit = (int)std::min<size_t>(3, (227u * 1024u) / (smem + 1024u));
    auto launch = [&](auto kernel, int threads) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int occ = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem);
        if (occ < 1) occ = 1;
        const uint64_t cap = (uint64_t)sm_count * occ * grid_waves();
        const uint32_t gx = (uint32_t)(a.n_reads < cap ? a.n_reads : cap);
        kernel<<<gx, threads, smem, st>>>(a, d_ptr, ids, cnt_words);
    };
    // lanes per list by the table's mean list length in 16-byte units (profiles/r2_ab_postings_sub_sweep.jsonl: 8 lanes win up
    // to ~12 units, the whole warp from ~23 on; 16 lanes per list lost everywhere); RB_POSTINGS_SUB = 0 / 2 / 4 / 8 forces one
    int lg = mean_units <= 0 ? 0 : mean_units <= kSub2Units ? 2 : mean_units <= kSub4Units ? 4 : mean_units <= kSub8Units ? 8 : 0;
    if (const char *e = std::getenv("RB_POSTINGS_SUB")) { const int v = std::atoi(e); if (v == 0 || v == 2 || v == 4 || v == 8) lg = v; }
#define RB_SUB_LAUNCH(LG)                                                                                                     \
    do {                                                                                                                      \
        if (narrow) {                                                                                                         \
            if (fit >= 3) launch(count_postings_sub_kernel<8, 256, LG>, 256);                                                 \
            else if (fit == 2) launch(count_postings_sub_kernel<8, 384, LG>, 384);                                            \
            else launch(count_postings_sub_kernel<8, 768, LG>, 768);                                                          \
        } else {                                                                                                              \
            if (fit >= 3) launch(count_postings_sub_kernel<16, 256, LG>, 256);                                                \
            else if (fit == 2) launch(count_postings_sub_kernel<16, 384, LG>, 384);                                           \
            else launch(count_postings_sub_kernel<16, 768, LG>, 768);                                                         \
        }                                                                                                                     \
        return cudaGetLastError() == cudaSuccess ? 1 : -1;                                                                    \
    } while (0)
    if (lg == 2) RB_SUB_LAUNCH(2);
    if (lg == 4) RB_SUB_LAUNCH(4);
    if (lg == 8) RB_SUB_LAUNCH(8);
#undef RB_SUB_LAUNCH
    if (narrow) {
        if (fit >= 3) launch(count_postings_kernel<8, 256>, 256);
        else if (fit == 2) launch(count_postings_kernel<8, 384>, 384);
        else launch(count_postings_kernel<8, 768>, 768);
    } else {
        if (fit >= 3) launch(count_postings_kernel<16, 256>, 256);
        else if (fit == 2) launch(count_postings_kernel<16, 384>, 384);
        else launch(count_postings_kernel<16, 768>, 768);
    }
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace rb
