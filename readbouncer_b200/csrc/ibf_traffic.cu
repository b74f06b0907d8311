// ibf_traffic.cu -- measurement aid: the DRAM bytes one count launch has to fetch, from the table geometry.
//
// bench.py's roofline needs the bytes the dominant kernel really moves (the reference's row probes are
// no longer what is fetched).  They follow from the layout: HBM delivers whole 128-byte lines (measured: a random
// 16-byte probe moves 114 bytes, profiles/r1_a_tile_cfg2_ncu_full.json), so the traffic of a launch is the number of
// distinct lines each table access touches, assuming nothing is found in L2 (tables are tens of GB).  One warp per
// read; the numbers are checked against ncu's dram__bytes in profiles/ (r2_*_traffic_check.json).
#include "ibf_device.cuh"

namespace rb {

struct TrafficArgs {
    const uint8_t *bases;
    const uint64_t *read_off;
    uint64_t n_reads;
    uint32_t k, n_hash;
    int kind;                  // 0 hashed row probes, 1 dense k-mer / window table, 2 postings (ptr + lists), 3 postings in slots
    int span;                  // positions per table entry (kind 1)
    uint32_t entry_bytes;      // kind 1: bytes per entry; kind 0: bytes per row; kind 3: bytes per slot
    const uint32_t *ptr;       // kind 2: list bounds; kind 3: the slots (header words)
    unsigned long long *out;   // [0] table bytes at line granularity, [1] table accesses (requests), [2] bases read
};

__device__ __forceinline__ uint64_t lines_of(uint64_t byte0, uint64_t nbytes)
{
    return nbytes ? ((byte0 + nbytes - 1) >> 7) - (byte0 >> 7) + 1 : 0;
}

__global__ void __launch_bounds__(256) traffic_kernel(const TrafficArgs a)
{
    const int lane = threadIdx.x & 31;
    const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    unsigned long long bytes = 0, reqs = 0, nb = 0;
    for (uint64_t read = warp0; read < a.n_reads; read += n_warps) {
        const uint64_t off = a.read_off[read], len = a.read_off[read + 1] - off;
        if (lane == 0) nb += len;
        if (read_flag_of(len, a.k) != 0) continue;
        const uint32_t npos = (uint32_t)len - a.k + 1;
        if (a.kind == 0) {
            if (lane == 0) { reqs += 2ull * npos * a.n_hash; bytes += 2ull * npos * a.n_hash * ((a.entry_bytes + 127u) / 128u) * 128u; }
        } else if (a.kind == 1) {
            const uint64_t n_ent = (npos + a.span - 1) / a.span;
            if (lane == 0) { reqs += n_ent; bytes += n_ent * ((a.entry_bytes + 127u) / 128u) * 128u; }
        } else {
            const uint32_t kbits = 2 * a.k;
            for (uint32_t p = lane; p < npos; p += 32) {
                uint32_t x = 0, bad = 0;
                for (uint32_t u = 0; u < a.k; ++u) {
                    const uint32_t d = dna5(a.bases[off + p + u]);
                    x = (x << 2) | (d & 3u);
                    bad |= d >> 2;
                }
                if (bad) continue;                                  // hashed path: rare, not counted
                uint32_t v = __brev(~x);
                v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
                const uint32_t idx[2] = {x & (kbits >= 32 ? ~0u : (1u << kbits) - 1u), v >> (32 - kbits)};
                for (int s = 0; s < 2; ++s) {
                    if (a.kind == 2) {
                        const uint32_t p0 = a.ptr[idx[s]], p1 = a.ptr[idx[s] + 1];
                        bytes += 128ull * (lines_of(4ull * idx[s], 8) + lines_of(16ull * p0, 16ull * (p1 - p0)));
                        reqs += 2;
                    } else {
                        // slots are line-aligned multiples of 128 bytes; a list that overflowed its slot costs its units too
                        const uint32_t *hdr = reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(a.ptr) + (uint64_t)idx[s] * a.entry_bytes);
                        bytes += a.entry_bytes;
                        reqs += 1;
                        if ((hdr[0] & 0xFFFFu) == 0xFFFFu) { bytes += 128ull * lines_of(16ull * hdr[1], 16ull * hdr[2]); reqs += 1; }
                    }
                }
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
        reqs += __shfl_xor_sync(0xffffffffu, reqs, o);
        nb += __shfl_xor_sync(0xffffffffu, nb, o);
    }
    if (lane == 0) { atomicAdd(a.out, bytes); atomicAdd(a.out + 1, reqs); atomicAdd(a.out + 2, nb); }
}

int launch_traffic(const uint8_t *bases, const uint64_t *read_off, uint64_t n_reads, uint32_t k, uint32_t n_hash, int kind, int span,
                   uint32_t entry_bytes, const uint32_t *ptr, unsigned long long *d_out, int sm_count, cudaStream_t st)
{
    TrafficArgs a{};
    a.bases = bases; a.read_off = read_off; a.n_reads = n_reads; a.k = k; a.n_hash = n_hash; a.kind = kind; a.span = span;
    a.entry_bytes = entry_bytes; a.ptr = ptr; a.out = d_out;
    if (cudaMemsetAsync(d_out, 0, 3 * sizeof(unsigned long long), st) != cudaSuccess) return -1;
    traffic_kernel<<<sm_count * 8, 256, 0, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace rb
