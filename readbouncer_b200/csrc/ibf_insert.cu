// ibf_insert.cu -- sm_100a kernel for the IBF build.
//
// Replaces seqan::insertKmer(filter, fragment, bin) as called once per reference
// fragment by IBF::add_sequences_to_filter (src/IBF/IBFBuild.cpp:143-215): every
// k-mer of the fragment sets bit (row(h_i(kmer)), bin) for each hash function i.
// All k-mers of a fragment land in the same bit column of different random rows,
// so the kernel is random 8-byte read-modify-writes: fire-and-forget 64-bit OR
// reductions (RED.OR at L2), no return value, no intra-word contention.
#include "ibf_kernels.cuh"

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

int launch_insert(const InsertArgs &a, uint64_t max_frag_len, int sm_count, cudaStream_t st)
{
    if (a.n_frags == 0) return 0;
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
