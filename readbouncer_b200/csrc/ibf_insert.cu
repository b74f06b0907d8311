// ibf_insert.cu -- sm_100a kernels for the IBF build.
//
// Replaces seqan::insertKmer(filter, fragment, bin) as called once per reference
// fragment by IBF::add_sequences_to_filter (src/IBF/IBFBuild.cpp:143-215): every
// k-mer of the fragment sets bit (row(h_i(kmer)), bin) for each hash function i.
//
// Two paths, same bits:
//  * insert_kernel: all k-mers of a fragment land in the same bit column of different random rows, so
//    the direct form is random 8-byte read-modify-writes: fire-and-forget 64-bit OR reductions (RED.OR
//    at L2).  Each one moves a 128-byte HBM line in and out: ~25 G RMW/s, the random-line ceiling.
//  * column build (insert_column_kernel + transpose_merge_kernel), for filters whose bit column
//    (noOfBlocks bits; 154 KB at the reference's default fragment_size = 100 000) fits in shared memory:
//    one CTA per bin sets the bits of that bin's fragments in a shared-memory bit column (ATOMS.OR),
//    writes it to a bin-major scratch matrix with coalesced stores, and a second kernel transposes
//    32 bins x 32 rows bit tiles in registers and ORs 64-byte row segments into the interleaved
//    matrix.  HBM sees each filter byte about four times, sequentially, instead of one random line
//    RMW per (k-mer, hash).
#include "ibf_kernels.cuh"
#include "ibf_transpose.cuh"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cub/device/device_scan.cuh>

namespace rb {

constexpr int kInsThreads = 128;
constexpr int kInsPerThread = 16;                       // consecutive k-mers rolled by one thread
constexpr int kInsChunk = kInsThreads * kInsPerThread;  // k-mer positions per CTA step

// blockIdx.x strides over fragments, blockIdx.y over the 2048-position chunks of a fragment.
__global__ void __launch_bounds__(kInsThreads) insert_kernel(const InsertArgs a)
{
    __shared__ uint8_t s_dig[kInsChunk + 32];
    const int tid = threadIdx.x;
    const HashParams &hp = a.hp;
    const uint32_t k = hp.k;

    for (uint64_t f = blockIdx.x; f < a.n_frags; f += gridDim.x) {
        const uint64_t fb = a.frag_begin[f], fe = a.frag_end[f], bin = a.frag_bin[f];
        if (fe < fb + k) continue;                       // shorter than k: inserts nothing
        if (bin >= a.n_bins) {                           // out-of-range bin (quirk Q3): reported, not written
            if (tid == 0 && blockIdx.y == 0) atomicExch(a.error_flag, 1u);
            continue;
        }
        if (bin < a.bin_begin || bin >= a.bin_end) continue;   // another shard's bin
        const uint64_t lb = bin - a.bin_begin;
        uint64_t *__restrict__ col = a.words + (lb >> 6);
        const unsigned long long bit = 1ULL << (lb & 63);
        const uint64_t npos = fe - fb - k + 1;

        for (uint64_t cs = (uint64_t)blockIdx.y * kInsChunk; cs < npos; cs += (uint64_t)gridDim.y * kInsChunk) {
            const uint32_t cn = (uint32_t)(npos - cs < (uint64_t)kInsChunk ? npos - cs : (uint64_t)kInsChunk);
            __syncthreads();
            for (uint32_t i = tid; i < cn + k - 1; i += kInsThreads) s_dig[i] = (uint8_t)dna5(a.bases[fb + cs + i]);
            __syncthreads();
            const uint32_t j0 = tid * kInsPerThread;
            const uint32_t j1 = min(j0 + (uint32_t)kInsPerThread, cn);
            if (j0 < j1) {
                uint64_t H = 0;
                for (uint32_t u = 0; u < k; ++u) H = H * 5 + s_dig[j0 + u];
                for (uint32_t j = j0; j < j1; ++j) {
                    for (uint32_t i = 0; i < hp.n_hash; ++i) {
                        uint64_t row = hash_row(H, hp.pre[i], hp.n_blocks, hp.magic);
                        atomicOr((unsigned long long *)(col + row * a.stride), bit);
                    }
                    if (j + 1 < j1) H = (H - (uint64_t)s_dig[j] * hp.top) * 5 + s_dig[j + k];
                }
            }
        }
    }
}

// ---- column build -------------------------------------------------------------------------------------
constexpr int kColThreads = 1024;
constexpr int kColPerThread = 16;
// k-mer positions staged per CTA step: with the misalignment (<= 15) and the k - 1 (<= 31) trailing bases a chunk
// spans at most kColThreads aligned 16-byte blocks, one per thread
constexpr int kColChunk = kColThreads * kColPerThread - 48;
constexpr int kColMaxSmem = 227 * 1024;
constexpr int kTrThreads = 512;                            // 16 warps x 32 bins = 512 bins (64-byte row segments)
constexpr int kTrBins = 512;
constexpr int kTrRows = 1024;                              // 32 lanes x 32 rows
constexpr int kTrWarpStride = 32 * 33 + 1;                 // padded so both the scatter and the row read-out spread over banks

struct ColumnArgs {
    InsertArgs in;
    uint32_t *bin_ptr;       // [n_local_bins + 1] fragment-list bounds (exclusive scan of the counts)
    uint32_t *bin_fill;      // [n_local_bins] cursor while the lists are filled
    uint32_t *frag_list;     // [n_frags] fragment indices grouped by local bin
    uint32_t *scratch;       // [pass bins][col_words32] bin-major bit columns
    uint64_t col_words32;    // 32-bit words per column (even, so columns are 8-byte aligned)
    uint64_t col_words32_pad; // the same rounded up to whole 16-byte units (shared-memory copy)
    uint64_t pass_bin0, pass_bin1;   // local bins of this pass
};

// fragments per local bin; out-of-range bins raise the error flag exactly like insert_kernel
__global__ void column_count_kernel(const ColumnArgs c)
{
    const InsertArgs &a = c.in;
    for (uint64_t f = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; f < a.n_frags; f += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t fb = a.frag_begin[f], fe = a.frag_end[f], bin = a.frag_bin[f];
        if (fe < fb + a.hp.k) continue;
        if (bin >= a.n_bins) { atomicExch(a.error_flag, 1u); continue; }
        if (bin < a.bin_begin || bin >= a.bin_end) continue;
        atomicAdd(&c.bin_ptr[bin - a.bin_begin], 1u);
    }
}

__global__ void column_fill_kernel(const ColumnArgs c)
{
    const InsertArgs &a = c.in;
    for (uint64_t f = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; f < a.n_frags; f += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t fb = a.frag_begin[f], fe = a.frag_end[f], bin = a.frag_bin[f];
        if (fe < fb + a.hp.k || bin >= a.n_bins || bin < a.bin_begin || bin >= a.bin_end) continue;
        const uint64_t lb = bin - a.bin_begin;
        c.frag_list[c.bin_ptr[lb] + atomicAdd(&c.bin_fill[lb], 1u)] = (uint32_t)f;
    }
}

// Four ASCII bases -> four Dna5 ranks (one per byte), branch-free: code = bits 1..2 of the upper-cased byte
// (A 0, C 1, T/U 2, G 3), rank = code ^ (code >> 1); a byte is valid iff it equals the letter its code names
// (U differs from T in bit 0 only), anything else is rank 4.  Same mapping as dna5() in ibf_common.cuh.
__device__ __forceinline__ uint32_t dna5x4(uint32_t c)
{
    const uint32_t x = c & 0xDFDFDFDFu;
    const uint32_t code = (x >> 1) & 0x03030303u;
    const uint32_t hi = (code >> 1) & 0x01010101u;                 // code 2 or 3
    const uint32_t t = hi & ~code;                                  // code == 2 (T/U), in bit 0 of the byte
    const uint32_t expect = 0x41414141u + 2u * code + 15u * t;      // 'A' 'C' 'T' 'G'
    const uint32_t diff = (x ^ expect) & ~t;
    const uint32_t bad = ((diff | ((diff & 0x7F7F7F7Fu) + 0x7F7F7F7Fu)) >> 7) & 0x01010101u;
    const uint32_t rank = code ^ hi;
    return (rank & ~(3u * bad)) | (bad << 2);
}

// One CTA per local bin of the pass: the bin's bit column lives in shared memory while its fragments are hashed.
// Bases are fetched as aligned 16-byte blocks one chunk ahead (the loads of chunk c+1 are in flight while chunk c
// is hashed), turned into rank bytes with SWAR arithmetic, and each thread rolls 16 consecutive k-mers.
template <int NH>   // number of hash functions; 0 = read it from the parameters
__global__ void __launch_bounds__(kColThreads, 1) insert_column_kernel(const ColumnArgs c)
{
    extern __shared__ __align__(16) uint32_t s_col[];
    const InsertArgs &a = c.in;
    const HashParams &hp = a.hp;
    const uint32_t k = hp.k;
    const uint32_t n_hash = NH ? (uint32_t)NH : hp.n_hash;
    const int tid = threadIdx.x;
    uint4 *s_dig4 = reinterpret_cast<uint4 *>(s_col + c.col_words32_pad);
    const uint8_t *s_dig = reinterpret_cast<const uint8_t *>(s_dig4);

    for (uint64_t lb = c.pass_bin0 + blockIdx.x; lb < c.pass_bin1; lb += gridDim.x) {
        const uint32_t l0 = c.bin_ptr[lb], l1 = c.bin_ptr[lb + 1];
        if (l0 == l1) continue;                               // no fragment: the merge kernel skips this column
        __syncthreads();
        for (uint64_t i = tid; i < c.col_words32_pad / 4; i += kColThreads) reinterpret_cast<uint4 *>(s_col)[i] = make_uint4(0, 0, 0, 0);
        for (uint32_t li = l0; li < l1; ++li) {
            const uint64_t f = c.frag_list[li];
            const uint64_t fb = a.frag_begin[f], fe = a.frag_end[f];
            const uint64_t npos = fe - fb - k + 1;
            const uintptr_t A = reinterpret_cast<uintptr_t>(a.bases) + fb;
            // a chunk's bytes [A + cs, A + cs + cn + k - 1) sit in at most kColThreads aligned 16-byte blocks
            uint4 pre = make_uint4(0, 0, 0, 0);
            {
                const uint32_t cn = (uint32_t)(npos < (uint64_t)kColChunk ? npos : (uint64_t)kColChunk);
                const uint32_t mis = (uint32_t)(A & 15);
                if ((uint32_t)tid * 16 < mis + cn + k - 1) pre = __ldg(reinterpret_cast<const uint4 *>(A - mis) + tid);
            }
            for (uint64_t cs = 0; cs < npos; cs += kColChunk) {
                const uint32_t cn = (uint32_t)(npos - cs < (uint64_t)kColChunk ? npos - cs : (uint64_t)kColChunk);
                const uint32_t mis = (uint32_t)((A + cs) & 15);
                __syncthreads();
                s_dig4[tid] = make_uint4(dna5x4(pre.x), dna5x4(pre.y), dna5x4(pre.z), dna5x4(pre.w));
                __syncthreads();
                if (cs + kColChunk < npos) {                       // next chunk's block, consumed after the hashing below
                    const uint64_t rest = npos - cs - kColChunk;
                    const uint32_t cn2 = (uint32_t)(rest < (uint64_t)kColChunk ? rest : (uint64_t)kColChunk);
                    const uintptr_t A2 = A + cs + kColChunk;
                    const uint32_t mis2 = (uint32_t)(A2 & 15);
                    if ((uint32_t)tid * 16 < mis2 + cn2 + k - 1) pre = __ldg(reinterpret_cast<const uint4 *>(A2 - mis2) + tid);
                }
                const uint32_t j0 = tid * kColPerThread;
                const uint32_t j1 = min(j0 + (uint32_t)kColPerThread, cn);
                if (j0 < j1) {
                    const uint8_t *d = s_dig + mis;
                    uint64_t H = 0;
                    for (uint32_t u = 0; u < k; ++u) H = H * 5 + d[j0 + u];
                    for (uint32_t j = j0; j < j1; ++j) {
                        if (NH) {
#pragma unroll
                            for (int i = 0; i < (NH ? NH : 1); ++i) {
                                const uint32_t row = (uint32_t)hash_row(H, hp.pre[i], hp.n_blocks, hp.magic);
                                atomicOr(&s_col[row >> 5], 1u << (row & 31));
                            }
                        } else {
                            for (uint32_t i = 0; i < n_hash; ++i) {
                                const uint32_t row = (uint32_t)hash_row(H, hp.pre[i], hp.n_blocks, hp.magic);
                                atomicOr(&s_col[row >> 5], 1u << (row & 31));
                            }
                        }
                        if (j + 1 < j1) H = (H - (uint64_t)d[j] * hp.top) * 5 + d[j + k];
                    }
                }
            }
        }
        __syncthreads();
        uint32_t *dst = c.scratch + (lb - c.pass_bin0) * c.col_words32;       // 8-byte aligned (col_words32 is even)
        for (uint64_t i = tid; i < c.col_words32 / 2; i += kColThreads)
            reinterpret_cast<uint2 *>(dst)[i] = reinterpret_cast<const uint2 *>(s_col)[i];
    }
}

// Columns too long for shared memory (fragment_size in the millions: one genome per bin): the same bin-major scratch
// columns, filled by 32-bit RED.OR from CTAs that walk the fragment lists in bin order.  The CTAs in flight at any time
// cover a few million k-mer positions, i.e. a handful of columns (~7 MB whatever the fragment size), so the reductions
// resolve in L2 (measured 180 G RED/s against 25 G/s for lines that miss) and every scratch line is written back once.
__global__ void __launch_bounds__(kInsThreads) insert_gcolumn_kernel(const ColumnArgs c, const uint32_t chunks_per_entry)
{
    __shared__ uint8_t s_dig[kInsChunk + 32];
    const InsertArgs &a = c.in;
    const HashParams &hp = a.hp;
    const uint32_t k = hp.k;
    const int tid = threadIdx.x;
    const uint32_t entry = blockIdx.x / chunks_per_entry, lane_chunk = blockIdx.x % chunks_per_entry;
    if (entry >= c.bin_ptr[a.bin_end - a.bin_begin]) return;
    const uint64_t f = c.frag_list[entry];
    const uint64_t lb = a.frag_bin[f] - a.bin_begin;
    if (lb < c.pass_bin0 || lb >= c.pass_bin1) return;
    const uint64_t fb = a.frag_begin[f], fe = a.frag_end[f];
    const uint64_t npos = fe - fb - k + 1;
    uint32_t *__restrict__ col = c.scratch + (lb - c.pass_bin0) * c.col_words32;
    for (uint64_t cs = (uint64_t)lane_chunk * kInsChunk; cs < npos; cs += (uint64_t)chunks_per_entry * kInsChunk) {
        const uint32_t cn = (uint32_t)(npos - cs < (uint64_t)kInsChunk ? npos - cs : (uint64_t)kInsChunk);
        __syncthreads();
        for (uint32_t i = tid; i < cn + k - 1; i += kInsThreads) s_dig[i] = (uint8_t)dna5(a.bases[fb + cs + i]);
        __syncthreads();
        const uint32_t j0 = tid * kInsPerThread;
        const uint32_t j1 = min(j0 + (uint32_t)kInsPerThread, cn);
        if (j0 < j1) {
            uint64_t H = 0;
            for (uint32_t u = 0; u < k; ++u) H = H * 5 + s_dig[j0 + u];
            for (uint32_t j = j0; j < j1; ++j) {
                for (uint32_t i = 0; i < hp.n_hash; ++i) {
                    const uint64_t row = hash_row(H, hp.pre[i], hp.n_blocks, hp.magic);
                    atomicOr(&col[row >> 5], 1u << (row & 31));
                }
                if (j + 1 < j1) H = (H - (uint64_t)s_dig[j] * hp.top) * 5 + s_dig[j + k];
            }
        }
    }
}

// Bin-major scratch columns -> OR into the row-interleaved matrix.  CTA tile: 512 bins x 1024 rows.  Warp w holds 32 bins:
// lane l loads rows 32l..32l+31 of each (32 coalesced 128-byte loads), transposes its 32x32 bit tile in registers and
// scatters 32 row words into shared memory; then 8 threads per row OR one 64-byte segment into the matrix.
__global__ void __launch_bounds__(kTrThreads) transpose_merge_kernel(const ColumnArgs c)
{
    extern __shared__ __align__(16) uint32_t s_t[];            // [16 warps][kTrWarpStride]
    const InsertArgs &a = c.in;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t row0 = (uint64_t)blockIdx.x * kTrRows;
    const uint64_t bin0 = c.pass_bin0 + (uint64_t)blockIdx.y * kTrBins;     // local bin; pass_bin0 is a multiple of 512
    const uint64_t wb = bin0 + warp * 32;
    const uint64_t cw = row0 / 32 + lane;                                     // this lane's word of every column

    uint32_t x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const uint64_t lb = wb + i;
        uint32_t v = 0;
        if (lb < c.pass_bin1 && cw < c.col_words32 && c.bin_ptr[lb] != c.bin_ptr[lb + 1])
            v = __ldcs(c.scratch + (lb - c.pass_bin0) * c.col_words32 + cw);
        x[i] = v;
    }
    transpose32(x);
#pragma unroll
    for (int j = 0; j < 32; ++j) s_t[warp * kTrWarpStride + lane * 33 + j] = x[j];
    __syncthreads();

    // 8 threads per row, one 64-bit matrix word each (two adjacent warps' 32-bin words)
    const uint64_t word0 = bin0 / 64;
    for (int rr = threadIdx.x >> 3; rr < kTrRows; rr += kTrThreads / 8) {
        const uint64_t row = row0 + rr;
        const int part = threadIdx.x & 7;
        const uint64_t w = word0 + part;
        if (row >= a.hp.n_blocks || w >= a.stride) continue;
        const int so = (rr >> 5) * 33 + (rr & 31);
        const uint64_t v = (uint64_t)s_t[(2 * part) * kTrWarpStride + so] | ((uint64_t)s_t[(2 * part + 1) * kTrWarpStride + so] << 32);
        if (v) {
            uint64_t *p = a.words + row * a.stride + w;
            *p |= v;
        }
    }
}

static std::atomic<int> g_insert_variant{0};
void set_insert_variant(int v) { g_insert_variant.store(v); }
int get_insert_variant() { return g_insert_variant.load(); }

// bytes of shared memory the column kernel needs for this filter, or 0 if the column does not fit
static size_t column_smem_bytes(const InsertArgs &a, uint64_t *col_words32)
{
    const uint64_t cw = ((a.hp.n_blocks + 63) / 64) * 2;
    *col_words32 = cw;
    const uint64_t need = ((cw + 3) & ~3ull) * 4 + (uint64_t)kColThreads * 16;
    return a.hp.n_blocks < (1ull << 32) && need <= (uint64_t)kColMaxSmem ? (size_t)need : 0;
}

constexpr uint64_t kGColMaxBytes = 48ull << 20;    // longest column the L2-resident scratch build takes

// 0 = direct RED.OR only, 1 = shared-memory columns, 2 = scratch columns in global memory (L2-resident reductions)
int insert_column_mode(const InsertArgs &a)
{
    uint64_t cw;
    if (a.n_frags >= (1ull << 32) || a.bin_end - a.bin_begin >= (1ull << 31)) return 0;
    if (column_smem_bytes(a, &cw) != 0) return 1;
    return a.hp.n_blocks < (1ull << 32) && cw * 4 <= kGColMaxBytes ? 2 : 0;
}

// Returns launches, -1 on a CUDA error, -2 when the scratch memory could not be had (caller falls back to insert_kernel).
static int launch_insert_columns(const InsertArgs &a, uint64_t max_frag_len, int sm_count, ScratchBuf *scr, cudaStream_t st)
{
    ColumnArgs c{};
    c.in = a;
    const size_t smem = column_smem_bytes(a, &c.col_words32);     // 0: the column lives in global scratch (L2)
    c.col_words32_pad = (c.col_words32 + 3) & ~3ull;
    auto kernel = a.hp.n_hash == 3 ? insert_column_kernel<3> : insert_column_kernel<0>;
    const uint64_t n_local = a.bin_end - a.bin_begin;
    const uint64_t col_bytes = c.col_words32 * 4;
    // scratch: whole 512-bin groups, as many as fit half of the free HBM (one pass for a 4.8 GB human filter)
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return -1;
    const uint64_t groups = (n_local + kTrBins - 1) / kTrBins;
    uint64_t budget = (free_b + scr->cap) / 2;             // the cached block is ours to reuse
    if (const char *e = std::getenv("RB_INSERT_SCRATCH_MB")) {            // tests: force several passes
        const uint64_t mb = std::strtoull(e, nullptr, 10);
        if (mb) budget = std::min<uint64_t>(budget, mb << 20);
    }
    uint64_t groups_per_pass = budget / (col_bytes * kTrBins);
    if (groups_per_pass == 0) return -2;
    if (groups_per_pass > groups) groups_per_pass = groups;
    // the last group of a pass may be partial: never allocate more columns than exist
    const uint64_t scratch_cols = groups_per_pass * kTrBins < n_local ? groups_per_pass * kTrBins : n_local;

    // one cached allocation (kept by the filter handle between calls, so only the first build pays for it):
    // bin_ptr [n_local + 1] | bin_fill [n_local] | frag_list [n_frags] | CUB scan storage | scratch columns
    const uint64_t meta_words = (n_local + 1) + n_local + a.n_frags;
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)(n_local + 1), st);
    const size_t meta_bytes = (meta_words * 4 + 255) & ~(size_t)255, tmp_pad = (tmp_bytes + 255) & ~(size_t)255;
    const size_t need = meta_bytes + tmp_pad + scratch_cols * col_bytes;
    if (need > scr->cap) {
        if (scr->p) cudaFree(scr->p);           // synchronises: nothing in flight uses the old block afterwards
        scr->p = nullptr; scr->cap = 0;
        if (cudaMalloc(&scr->p, need) != cudaSuccess) { cudaGetLastError(); scr->p = nullptr; return -2; }
        scr->cap = need;
    }
    uint8_t *const base = static_cast<uint8_t *>(scr->p);
    uint32_t *const d_meta = reinterpret_cast<uint32_t *>(base);
    void *const d_tmp = base + meta_bytes;
    c.scratch = reinterpret_cast<uint32_t *>(base + meta_bytes + tmp_pad);
    c.bin_ptr = d_meta; c.bin_fill = d_meta + n_local + 1; c.frag_list = c.bin_fill + n_local;
    int launches = 0;
    bool ok = cudaMemsetAsync(d_meta, 0, (2 * n_local + 1) * 4, st) == cudaSuccess;
    const uint32_t fgrid = (uint32_t)std::min<uint64_t>((a.n_frags + 255) / 256, (uint64_t)sm_count * 8);
    if (ok) {
        column_count_kernel<<<fgrid, 256, 0, st>>>(c);
        cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, c.bin_ptr, c.bin_ptr, (int)(n_local + 1), st);
        column_fill_kernel<<<fgrid, 256, 0, st>>>(c);
        launches += 3;
    }
    constexpr size_t tr_smem = 16 * kTrWarpStride * sizeof(uint32_t);
    if (ok)     // per-device attributes; setting them again is harmless
        ok = (!smem || cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kColMaxSmem) == cudaSuccess) &&
             cudaFuncSetAttribute(transpose_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tr_smem) == cudaSuccess;
    for (uint64_t g0 = 0; ok && g0 < groups; g0 += groups_per_pass) {
        c.pass_bin0 = g0 * kTrBins;
        c.pass_bin1 = std::min<uint64_t>((g0 + groups_per_pass) * kTrBins, n_local);
        const uint64_t nb = c.pass_bin1 - c.pass_bin0;
        if (smem) {
            kernel<<<(uint32_t)std::min<uint64_t>(nb, 1u << 20), kColThreads, smem, st>>>(c);
        } else {
            ok = cudaMemsetAsync(c.scratch, 0, nb * col_bytes, st) == cudaSuccess;
            uint64_t chunks = max_frag_len ? (max_frag_len + kInsChunk - 1) / kInsChunk : 64;
            chunks = std::max<uint64_t>(1, std::min<uint64_t>(chunks, 0x7FFFFFFFull / a.n_frags));
            if (ok) insert_gcolumn_kernel<<<(uint32_t)(a.n_frags * chunks), kInsThreads, 0, st>>>(c, (uint32_t)chunks);
        }
        const uint64_t row_tiles = (a.hp.n_blocks + kTrRows - 1) / kTrRows;
        const uint64_t bin_tiles = (nb + kTrBins - 1) / kTrBins;
        transpose_merge_kernel<<<dim3((uint32_t)row_tiles, (uint32_t)bin_tiles), kTrThreads, tr_smem, st>>>(c);
        launches += 2;
        ok = cudaGetLastError() == cudaSuccess;
    }
    return ok && cudaGetLastError() == cudaSuccess ? launches : -1;
}

int launch_insert(const InsertArgs &a, uint64_t max_frag_len, int sm_count, ScratchBuf *scr, cudaStream_t st)
{
    if (a.n_frags == 0) return 0;
    // column build (variant 0 auto, 1 always RED.OR, 2 column build whenever applicable).  Auto: shared-memory columns
    // need one bin per SM to fill the GPU; scratch columns need enough k-mers to pay for the scratch and the merge pass.
    const int variant = get_insert_variant();
    const int mode = variant == 1 ? 0 : insert_column_mode(a);
    const uint64_t n_local = a.bin_end - a.bin_begin;
    if (mode != 0 && scr && (variant == 2 || (mode == 1 && a.n_frags >= (uint64_t)sm_count && n_local >= (uint64_t)sm_count) ||
                      (mode == 2 && a.n_frags * (max_frag_len ? max_frag_len : 1) >= (64ull << 20)))) {
        int n = launch_insert_columns(a, max_frag_len, sm_count, scr, st);
        if (n != -2) return n;
    }
    uint32_t gx = (uint32_t)(a.n_frags < 16384 ? a.n_frags : 16384);
    uint64_t chunks = max_frag_len ? (max_frag_len + kInsChunk - 1) / kInsChunk : 16;
    // enough CTAs for ~16 per SM, but never more chunk lanes than the longest fragment has chunks
    uint64_t want = ((uint64_t)sm_count * 16 + gx - 1) / gx;
    uint64_t gy = want < chunks ? want : chunks;
    if (gy < 1) gy = 1;
    if (gy > 65535) gy = 65535;
    insert_kernel<<<dim3(gx, (uint32_t)gy), kInsThreads, 0, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace rb
