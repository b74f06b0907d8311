// ibf_count.cu -- sm_100a kernels for the classify hot path.
//
// Replaces, for a whole batch of reads at once, what ReadBouncer does one read
// at a time in Read::count_matches / find_matches (src/IBF/IBFClassify.cpp:81-171):
//   seqan::count(filter, read), seqan::count(filter, revcomp(read)),
//   threshold lookup, select_matches (hit) and max_matches (max_count).
//
// Two kernels, both integer-only and memory-bound (no tensor cores):
//
//  count_tile_kernel    warp per read, lane per run of consecutive k-mers.  Each probe reads a
//                       WT-word (<= 32 B) tile of a row, so the access pattern is random 32 B
//                       sectors: the right shape for narrow filters (<= 256 bins; BASELINE
//                       configs #1, #2).  Per-bin counters live in shared memory.  With
//                       gridDim.y > 1 it walks column tiles of a wide filter (fallback path).
//
//  count_stream_kernel  CTA per (read, 256-word column block).  Rows are contiguous, so each
//                       k-mer streams 3 coalesced row segments per strand; every thread owns two
//                       words of the block and keeps its 128 bins x 2 strands in bit-sliced
//                       vertical counters in registers (NP planes), so counting costs 2 logic ops
//                       per plane per 64 bins and no shared-memory traffic (configs #3, #5).
#include "ibf_device.cuh"

#include <cstdlib>

namespace rb {

// ------------------------------------------------------------------------------------------
// tile kernel
// ------------------------------------------------------------------------------------------

// NH = 3: the reference's fixed hash_functions (src/IBF/IBFConfig.hpp:71), fully unrolled with
// U positions (2*3*U probes) in flight per lane.  NH = 0: runtime n_hash, one probe at a time.
template <int WT, bool A16, int NH, int U>
__global__ void __launch_bounds__(kTileWarps * 32)
count_tile_kernel(const CountArgs a, const uint32_t c0, const int multi_tile)
{
    __shared__ __align__(16) uint8_t s_dig[kTileWarps][kDigBytes];
    __shared__ uint32_t s_cnt[kTileWarps][2][64 * WT];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t wc = c0 + blockIdx.y * WT;          // first local row word of this tile
    const uint64_t total_warps = (uint64_t)gridDim.x * kTileWarps;
    const HashParams &hp = a.fv.hp;
    const uint32_t k = hp.k;
    const uint64_t *__restrict__ words = a.fv.words + wc;
    const uint64_t stride = a.fv.stride;
    uint8_t *dig = s_dig[warp];
    uint32_t *cntF = s_cnt[warp][0], *cntR = s_cnt[warp][1];

    for (int b = lane; b < 64 * WT; b += 32) { cntF[b] = 0; cntR[b] = 0; }
    __syncwarp();

    for (uint64_t read = (uint64_t)blockIdx.x * kTileWarps + warp; read < a.n_reads; read += total_warps) {
        const uint64_t off = a.read_off[read];
        const uint64_t len = a.read_off[read + 1] - off;
        const uint32_t flag = read_flag_of(len, k);
        if (lane == 0 && blockIdx.y == 0 && a.read_flag) a.read_flag[read] = (uint8_t)flag;

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
                    uint64_t Hf = 0, Hr = 0, pw = 1;
                    for (uint32_t u = 0; u < k; ++u) {
                        uint32_t d = dig[j0 + u];
                        Hf = Hf * 5 + d;
                        Hr += comp5(d) * pw;
                        pw *= 5;
                    }
                    if constexpr (NH == 3) {
                        for (uint32_t j = j0; j < j1; j += U) {
                            uint64_t v[U][2][3][WT];
#pragma unroll
                            for (int u = 0; u < U; ++u) {
                                if (j + u < j1) {
#pragma unroll
                                    for (int i = 0; i < 3; ++i) {
                                        uint64_t rf = hash_row(Hf, hp.pre[i], hp.n_blocks, hp.magic);
                                        uint64_t rr = hash_row(Hr, hp.pre[i], hp.n_blocks, hp.magic);
                                        load_tile<WT, A16>(words + rf * stride, v[u][0][i]);
                                        load_tile<WT, A16>(words + rr * stride, v[u][1][i]);
                                    }
                                    if (j + u + 1 < j1) {
                                        uint32_t dout = dig[j + u], din = dig[j + u + k];
                                        Hf = (Hf - dout * hp.top) * 5 + din;
                                        Hr = (Hr - comp5(dout)) * kInv5 + comp5(din) * hp.top;
                                    }
                                }
                            }
#pragma unroll
                            for (int u = 0; u < U; ++u) {
                                if (j + u < j1) {
                                    uint64_t mf[WT], mr[WT];
#pragma unroll
                                    for (int w = 0; w < WT; ++w) {
                                        mf[w] = v[u][0][0][w] & v[u][0][1][w] & v[u][0][2][w];
                                        mr[w] = v[u][1][0][w] & v[u][1][1][w] & v[u][1][2][w];
                                    }
                                    count_bits<WT>(mf, cntF);
                                    count_bits<WT>(mr, cntR);
                                }
                            }
                        }
                    } else {
                        for (uint32_t j = j0; j < j1; ++j) {
                            uint64_t mf[WT], mr[WT];
#pragma unroll
                            for (int w = 0; w < WT; ++w) { mf[w] = ~0ULL; mr[w] = ~0ULL; }
                            for (uint32_t i = 0; i < hp.n_hash; ++i) {
                                uint64_t rf = hash_row(Hf, hp.pre[i], hp.n_blocks, hp.magic);
                                uint64_t rr = hash_row(Hr, hp.pre[i], hp.n_blocks, hp.magic);
                                uint64_t vf[WT], vr[WT];
                                load_tile<WT, A16>(words + rf * stride, vf);
                                load_tile<WT, A16>(words + rr * stride, vr);
#pragma unroll
                                for (int w = 0; w < WT; ++w) { mf[w] &= vf[w]; mr[w] &= vr[w]; }
                            }
                            count_bits<WT>(mf, cntF);
                            count_bits<WT>(mr, cntR);
                            if (j + 1 < j1) {
                                uint32_t dout = dig[j], din = dig[j + k];
                                Hf = (Hf - dout * hp.top) * 5 + din;
                                Hr = (Hr - comp5(dout)) * kInv5 + comp5(din) * hp.top;
                            }
                        }
                    }
                }
            }
        }
        __syncwarp();

        tile_epilogue<WT>(a, read, len, flag, wc, cntF, cntR, lane, multi_tile);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// streaming kernel
// ------------------------------------------------------------------------------------------
constexpr int kStreamThreads = 128;
constexpr int kStreamCB = 2 * kStreamThreads;   // row words per column block
constexpr int kStreamPosChunk = 512;            // k-mer positions hashed per phase

// add the 64 one-bit values in `m` to the bit-sliced counters pl[0..NP)
template <int NP>
__device__ __forceinline__ void planes_add(uint64_t (&pl)[NP], uint64_t m)
{
    uint64_t c = m;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        uint64_t t = pl[p] & c;
        pl[p] ^= c;
        c = t;
    }
}

// bitmask of the 64 counters that are >= thr
template <int NP>
__device__ __forceinline__ uint64_t planes_ge(const uint64_t (&pl)[NP], uint32_t thr)
{
    if (thr >> NP) return 0;
    uint64_t gt = 0, eq = ~0ULL;
#pragma unroll
    for (int p = NP - 1; p >= 0; --p) {
        uint64_t tb = ((thr >> p) & 1u) ? ~0ULL : 0ULL;
        gt |= eq & pl[p] & ~tb;
        eq &= ~(pl[p] ^ tb);
    }
    return gt | eq;
}

template <int NP>
__device__ __forceinline__ uint32_t planes_get(const uint64_t (&pl)[NP], int b)
{
    uint32_t c = 0;
#pragma unroll
    for (int p = 0; p < NP; ++p) c |= (uint32_t)((pl[p] >> b) & 1ULL) << p;
    return c;
}

// V = 2: thread owns words {2t, 2t+1} of its column block (16-byte loads; needs an even row
// stride).  V = 1: thread owns words {t, t+128} (8-byte loads, any stride).
template <int NP, int V>
__global__ void __launch_bounds__(kStreamThreads)
count_stream_kernel(const CountArgs a)
{
    __shared__ uint8_t s_dig[kStreamPosChunk + 32];
    __shared__ uint32_t s_rows[kStreamPosChunk][6];
    __shared__ uint64_t s_red[kStreamThreads / 32][kMaxLut];

    const int tid = threadIdx.x;
    const uint64_t read = blockIdx.x;
    const HashParams &hp = a.fv.hp;
    const uint32_t k = hp.k;
    const uint64_t stride = a.fv.stride;
    const uint64_t off = a.read_off[read];
    const uint64_t len = a.read_off[read + 1] - off;
    uint32_t flag = read_flag_of(len, k);
    // a read longer than the caller's max_read_len promised does not fit this variant's NP-bit counters: flag 3, not classified
    if (flag == 0 && NP < 16 && len - k + 1 > (1u << NP) - 1u) flag = 3;
    if (tid == 0 && blockIdx.y == 0 && a.read_flag) a.read_flag[read] = (uint8_t)flag;

    const uint64_t cb0 = (uint64_t)blockIdx.y * kStreamCB;
    const uint64_t w0 = cb0 + (V == 2 ? 2 * tid : tid);
    const uint64_t w1 = cb0 + (V == 2 ? 2 * tid + 1 : tid + kStreamThreads);
    const bool ok0 = w0 < stride, ok1 = w1 < stride;
    const uint64_t *__restrict__ words = a.fv.words;

    uint64_t plF0[NP], plF1[NP], plR0[NP], plR1[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) { plF0[p] = 0; plF1[p] = 0; plR0[p] = 0; plR1[p] = 0; }

    if (flag == 0) {
        const uint32_t npos = (uint32_t)len - k + 1;
        for (uint32_t cs = 0; cs < npos; cs += kStreamPosChunk) {
            const uint32_t cn = min((uint32_t)kStreamPosChunk, npos - cs);
            __syncthreads();
            for (uint32_t i = tid; i < cn + k - 1; i += kStreamThreads) s_dig[i] = (uint8_t)dna5(a.bases[off + cs + i]);
            __syncthreads();
            // phase 1: the 6 row indices of every position of the chunk
            for (uint32_t j = tid; j < cn; j += kStreamThreads) {
                uint64_t Hf = 0, Hr = 0, pw = 1;
                for (uint32_t u = 0; u < k; ++u) {
                    uint32_t d = s_dig[j + u];
                    Hf = Hf * 5 + d;
                    Hr += comp5(d) * pw;
                    pw *= 5;
                }
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    s_rows[j][i] = (uint32_t)hash_row(Hf, hp.pre[i], hp.n_blocks, hp.magic);
                    s_rows[j][3 + i] = (uint32_t)hash_row(Hr, hp.pre[i], hp.n_blocks, hp.magic);
                }
            }
            __syncthreads();
            // phase 2: stream the rows; two positions (12 row segments) in flight per thread
            for (uint32_t j = 0; j < cn; j += 2) {
                uint64_t v0[2][6], v1[2][6];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const bool live = j + u < cn;
#pragma unroll
                    for (int i = 0; i < 6; ++i) {
                        v0[u][i] = 0; v1[u][i] = 0;
                        if (live) {
                            const uint64_t *rp = words + (uint64_t)s_rows[j + u][i] * stride;
                            if constexpr (V == 2) {
                                if (ok0) {
                                    ulonglong2 t = __ldg(reinterpret_cast<const ulonglong2 *>(rp + w0));
                                    v0[u][i] = t.x; v1[u][i] = t.y;
                                }
                            } else {
                                if (ok0) v0[u][i] = __ldg(rp + w0);
                                if (ok1) v1[u][i] = __ldg(rp + w1);
                            }
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    planes_add<NP>(plF0, v0[u][0] & v0[u][1] & v0[u][2]);
                    planes_add<NP>(plF1, v1[u][0] & v1[u][1] & v1[u][2]);
                    planes_add<NP>(plR0, v0[u][3] & v0[u][4] & v0[u][5]);
                    planes_add<NP>(plR1, v1[u][3] & v1[u][4] & v1[u][5]);
                }
            }
        }
    }

    // ---- epilogue -----------------------------------------------------------------------------
    const uint64_t nbl = a.fv.n_bins_local;
    uint64_t valid0 = 0, valid1 = 0;   // bins of my words that exist in the filter
    if (ok0 && w0 * 64 < nbl) valid0 = (nbl - w0 * 64 >= 64) ? ~0ULL : ((1ULL << (nbl - w0 * 64)) - 1);
    if (ok1 && w1 * 64 < nbl) valid1 = (nbl - w1 * 64 >= 64) ? ~0ULL : ((1ULL << (nbl - w1 * 64)) - 1);

    if (a.counts_fwd || a.counts_rev) {
        for (int b = 0; b < 64; ++b) {
            if ((valid0 >> b) & 1ULL) {
                size_t o = read * nbl + w0 * 64 + b;
                if (a.counts_fwd) a.counts_fwd[o] = (uint16_t)planes_get<NP>(plF0, b);
                if (a.counts_rev) a.counts_rev[o] = (uint16_t)planes_get<NP>(plR0, b);
            }
            if ((valid1 >> b) & 1ULL) {
                size_t o = read * nbl + w1 * 64 + b;
                if (a.counts_fwd) a.counts_fwd[o] = (uint16_t)planes_get<NP>(plF1, b);
                if (a.counts_rev) a.counts_rev[o] = (uint16_t)planes_get<NP>(plR1, b);
            }
        }
    }

    uint64_t best[kMaxLut];
#pragma unroll
    for (int t = 0; t < kMaxLut; ++t) {
        best[t] = 0;
        if (t < (int)a.n_lut && flag == 0) {
            const uint32_t thr = (uint32_t)__ldg(a.lut + (size_t)t * kLutSize + len);
            uint64_t pass0 = (planes_ge<NP>(plF0, thr) | planes_ge<NP>(plR0, thr)) & valid0;
            uint64_t pass1 = (planes_ge<NP>(plF1, thr) | planes_ge<NP>(plR1, thr)) & valid1;
            while (pass0) {
                int b = __ffsll((long long)pass0) - 1;
                pass0 &= pass0 - 1;
                uint32_t m = max(planes_get<NP>(plF0, b), planes_get<NP>(plR0, b));
                uint64_t key = pack_key(m, (uint32_t)(a.fv.bin_begin + w0 * 64 + b));
                best[t] = key > best[t] ? key : best[t];
            }
            while (pass1) {
                int b = __ffsll((long long)pass1) - 1;
                pass1 &= pass1 - 1;
                uint32_t m = max(planes_get<NP>(plF1, b), planes_get<NP>(plR1, b));
                uint64_t key = pack_key(m, (uint32_t)(a.fv.bin_begin + w1 * 64 + b));
                best[t] = key > best[t] ? key : best[t];
            }
        }
    }
#pragma unroll
    for (int t = 0; t < kMaxLut; ++t) {
        uint64_t bk = warp_max_u64(best[t]);
        if ((tid & 31) == 0) s_red[tid >> 5][t] = bk;
    }
    __syncthreads();
    if (tid < (int)a.n_lut) {
        uint64_t bk = 0;
#pragma unroll
        for (int w = 0; w < kStreamThreads / 32; ++w) bk = s_red[w][tid] > bk ? s_red[w][tid] : bk;
        if (bk) key_max(a.keys + (size_t)tid * a.n_reads + read, bk, a.keys_shared);
    }
}

// ------------------------------------------------------------------------------------------
// keys -> (max_count, hit, argmax_bin)
// ------------------------------------------------------------------------------------------
__global__ void keys_decode_kernel(const uint64_t *__restrict__ keys, uint64_t n, uint16_t *max_count,
                                   uint8_t *hit, uint32_t *argmax_bin)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t key = keys[i];
        if (max_count) max_count[i] = (uint16_t)((key >> 32) & 0xFFFFu);
        if (hit) hit[i] = (uint8_t)((key >> 48) & 1u);
        if (argmax_bin) argmax_bin[i] = key ? ~(uint32_t)(key & 0xFFFFFFFFu) : 0xFFFFFFFFu;
    }
}

// Piece of a batch -> the caller's batch-shaped result arrays: keys [n_lut][n] of reads [out_off, out_off+n) go to
// out[t * stride + out_off + i].  The outputs may be mapped pinned host memory (zero-copy results: consecutive
// threads write consecutive elements, so the stores leave the GPU as full PCIe write bursts).
__global__ void keys_decode_piece_kernel(const uint64_t *__restrict__ keys, const uint8_t *__restrict__ flag_in, uint64_t n,
                                         uint32_t n_lut, uint64_t stride, uint64_t out_off, uint16_t *max_count,
                                         uint8_t *hit, uint32_t *argmax_bin, uint8_t *flag_out)
{
    const uint64_t total = n * n_lut;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t t = j / n, i = j - t * n, o = t * stride + out_off + i;
        const uint64_t key = keys[j];
        if (max_count) max_count[o] = (uint16_t)((key >> 32) & 0xFFFFu);
        if (hit) hit[o] = (uint8_t)((key >> 48) & 1u);
        if (argmax_bin) argmax_bin[o] = key ? ~(uint32_t)(key & 0xFFFFFFFFu) : 0xFFFFFFFFu;
        if (t == 0 && flag_out) flag_out[out_off + i] = flag_in[i];
    }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
template <int WT, bool A16, int NH, int U>
static void launch_tile_one(const CountArgs &a, uint32_t c0, uint32_t n_tiles, int multi_tile, int sm_count,
                            cudaStream_t st)
{
    // persistent grid: exactly the resident CTAs of every SM, reads strided across their warps
    static int occ = 0;
    if (occ == 0) {
        int o = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, count_tile_kernel<WT, A16, NH, U>, kTileWarps * 32, 0);
        occ = o > 0 ? o : 1;
    }
    uint64_t blocks_needed = (a.n_reads + kTileWarps - 1) / kTileWarps;
    uint64_t max_x = (uint64_t)sm_count * occ * grid_waves();
    if (n_tiles > 1) max_x = (max_x + n_tiles - 1) / n_tiles;
    uint32_t gx = (uint32_t)(blocks_needed < max_x ? blocks_needed : max_x);
    if (gx == 0) gx = 1;
    dim3 grid(gx, n_tiles);
    count_tile_kernel<WT, A16, NH, U><<<grid, kTileWarps * 32, 0, st>>>(a, c0, multi_tile);
}

template <int WT, bool A16>
static void launch_tile_wt(const CountArgs &a, uint32_t c0, uint32_t n_tiles, int multi_tile, int sm_count,
                           cudaStream_t st)
{
    if (a.fv.hp.n_hash == 3) launch_tile_one<WT, A16, 3, (WT <= 2 ? 2 : 1)>(a, c0, n_tiles, multi_tile, sm_count, st);
    else launch_tile_one<WT, A16, 0, 1>(a, c0, n_tiles, multi_tile, sm_count, st);
}

static int launch_tile(const CountArgs &a, int sm_count, cudaStream_t st)
{
    const uint64_t W = a.fv.stride;
    const bool a16 = (W % 2 == 0);
    int launches = 0;
    if (W <= 4) {
        switch (W) {
        case 1: launch_tile_wt<1, false>(a, 0, 1, 0, sm_count, st); break;
        case 2: launch_tile_wt<2, true>(a, 0, 1, 0, sm_count, st); break;
        case 3: launch_tile_wt<3, false>(a, 0, 1, 0, sm_count, st); break;
        default: launch_tile_wt<4, true>(a, 0, 1, 0, sm_count, st); break;
        }
        return 1;
    }
    // wide filter through the tile kernel: column tiles of 4 words, then the remainder
    const uint32_t full = (uint32_t)(W / 4), rem = (uint32_t)(W % 4);
    for (uint32_t t0 = 0; t0 < full; t0 += 65535u) {
        uint32_t nt = full - t0 < 65535u ? full - t0 : 65535u;
        if (a16) launch_tile_wt<4, true>(a, t0 * 4, nt, 1, sm_count, st);
        else launch_tile_wt<4, false>(a, t0 * 4, nt, 1, sm_count, st);
        ++launches;
    }
    if (rem) {
        const uint32_t c0 = full * 4;
        if (rem == 1) launch_tile_wt<1, false>(a, c0, 1, 1, sm_count, st);
        else if (rem == 2) { if (a16) launch_tile_wt<2, true>(a, c0, 1, 1, sm_count, st); else launch_tile_wt<2, false>(a, c0, 1, 1, sm_count, st); }
        else launch_tile_wt<3, false>(a, c0, 1, 1, sm_count, st);
        ++launches;
    }
    return launches;
}

template <int NP>
static void launch_stream_np(const CountArgs &a, cudaStream_t st)
{
    // gridDim.x carries the reads (launch_count caps them at 2^31-1), gridDim.y the column blocks
    const uint32_t n_cb = (uint32_t)((a.fv.stride + kStreamCB - 1) / kStreamCB);
    dim3 grid((uint32_t)a.n_reads, n_cb);
    if (a.fv.stride % 2 == 0) count_stream_kernel<NP, 2><<<grid, kStreamThreads, 0, st>>>(a);
    else count_stream_kernel<NP, 1><<<grid, kStreamThreads, 0, st>>>(a);
}

static int launch_stream(const CountArgs &a, uint32_t max_read_len, cudaStream_t st)
{
    const uint32_t k = a.fv.hp.k;
    uint32_t max_pos = (max_read_len == 0 || max_read_len > 65535u) ? 65535u : max_read_len;
    max_pos = max_pos >= k ? max_pos - k + 1 : 0;
    if (max_pos <= 255) launch_stream_np<8>(a, st);
    else if (max_pos <= 1023) launch_stream_np<10>(a, st);
    else launch_stream_np<16>(a, st);
    return 1;
}

int grid_waves()
{
    static const int waves = [] {
        const char *e = std::getenv("RB_GRID_WAVES");
        const int v = e ? std::atoi(e) : 0;
        return v > 0 ? (v > 256 ? 256 : v) : 16;
    }();
    return waves;
}

int launch_count(const CountArgs &a, uint32_t max_read_len, int which, int sm_count, cudaStream_t st)
{
    if (a.n_reads == 0) return 0;
    if (a.n_reads > 0x7FFFFFFFull) return -1;
    if (a.n_lut == 0 || a.n_lut > (uint32_t)kMaxLut) return -1;
    const bool stream_ok = a.fv.hp.n_hash == 3 && a.fv.hp.n_blocks <= 0xFFFFFFFFull;
    bool use_stream;
    if (which == 1) use_stream = false;
    else if (which == 2) use_stream = stream_ok;
    else use_stream = stream_ok && a.fv.stride > 4;
    int launches = 0;
    if ((use_stream || a.fv.stride > 4) && !a.keys_shared) {
        // both multi-block paths combine their partial summaries with atomicMax (a shared key array is zeroed by its owner)
        cudaMemsetAsync(a.keys, 0, sizeof(uint64_t) * a.n_lut * a.n_reads, st);
    }
    launches += use_stream ? launch_stream(a, max_read_len, st) : launch_tile(a, sm_count, st);
    return cudaGetLastError() == cudaSuccess ? launches : -1;
}

int launch_keys_decode(const uint64_t *keys, uint64_t n, uint16_t *max_count, uint8_t *hit,
                       uint32_t *argmax_bin, cudaStream_t st)
{
    if (n == 0) return 0;
    uint64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    keys_decode_kernel<<<(uint32_t)blocks, 256, 0, st>>>(keys, n, max_count, hit, argmax_bin);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_keys_decode_piece(const uint64_t *keys, const uint8_t *flag_in, uint64_t n, uint32_t n_lut, uint64_t stride,
                             uint64_t out_off, uint16_t *max_count, uint8_t *hit, uint32_t *argmax_bin, uint8_t *flag_out,
                             cudaStream_t st)
{
    if (n == 0) return 0;
    uint64_t blocks = (n * n_lut + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    keys_decode_piece_kernel<<<(uint32_t)blocks, 256, 0, st>>>(keys, flag_in, n, n_lut, stride, out_off, max_count, hit,
                                                             argmax_bin, flag_out);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace rb
