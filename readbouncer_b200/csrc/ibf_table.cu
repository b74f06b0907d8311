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
#include "ibf_device.cuh"

namespace rb {

// ------------------------------------------------------------------------------------------
// table build: one thread per k-mer value
// ------------------------------------------------------------------------------------------
template <int WT>
__global__ void __launch_bounds__(256) table_build_kernel(const FilterView fv, uint64_t *__restrict__ table,
                                                          const uint64_t n_entries)
{
    const HashParams &hp = fv.hp;
    const uint32_t k = hp.k;
    for (uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; x < n_entries;
         x += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t Hf = 0, Hr = 0, pw = 1;
        for (uint32_t j = 0; j < k; ++j) {
            uint32_t d = (uint32_t)(x >> (2 * (k - 1 - j))) & 3u;   // j-th base of the k-mer
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
        uint64_t *e = table + x * (2 * WT);
#pragma unroll
        for (int w = 0; w < WT; ++w) { e[w] = mf[w]; e[WT + w] = mr[w]; }
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
// launchers
// ------------------------------------------------------------------------------------------
template <int WT>
static void launch_build_wt(const FilterView &fv, uint64_t *table, uint64_t n_entries, int sm_count, cudaStream_t st)
{
    uint64_t blocks = (n_entries + 255) / 256;
    uint64_t cap = (uint64_t)sm_count * 32;
    table_build_kernel<WT><<<(uint32_t)(blocks < cap ? blocks : cap), 256, 0, st>>>(fv, table, n_entries);
}

int launch_table_build(const FilterView &fv, uint64_t *table, uint64_t n_entries, int sm_count, cudaStream_t st)
{
    switch (fv.stride) {
    case 1: launch_build_wt<1>(fv, table, n_entries, sm_count, st); break;
    case 2: launch_build_wt<2>(fv, table, n_entries, sm_count, st); break;
    case 3: launch_build_wt<3>(fv, table, n_entries, sm_count, st); break;
    case 4: launch_build_wt<4>(fv, table, n_entries, sm_count, st); break;
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
    uint64_t max_x = (uint64_t)sm_count * occ;
    uint32_t gx = (uint32_t)(blocks_needed < max_x ? blocks_needed : max_x);
    count_table_kernel<WT, U><<<gx ? gx : 1, kTileWarps * 32, 0, st>>>(a, table);
}

int launch_count_table(const CountArgs &a, const uint64_t *table, int sm_count, cudaStream_t st)
{
    if (a.n_reads == 0) return 0;
    if (a.n_lut == 0 || a.n_lut > (uint32_t)kMaxLut) return -1;
    switch (a.fv.stride) {
    case 1: launch_table_wt<1, 2>(a, table, sm_count, st); break;
    case 2: launch_table_wt<2, 2>(a, table, sm_count, st); break;
    case 3: launch_table_wt<3, 1>(a, table, sm_count, st); break;
    case 4: launch_table_wt<4, 1>(a, table, sm_count, st); break;
    default: return -1;
    }
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace rb
