// ibf_microbench.cu -- random-sector gather microbenchmark.
//
// Establishes the measured ceiling for the tile kernel's access pattern (SURVEY.md section 8d:
// "a random-32B-gather microbenchmark must be added to establish the random-sector ceilings"):
// every thread issues independent, uniformly random, row-aligned loads of ROWB bytes over a
// buffer, 8 in flight per thread, and XOR-folds them into a sink so nothing is optimised away.
#include "../../include/rb_ibf.h"
#include "ibf_common.cuh"

namespace rb {

__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

template <int ROWB>
__global__ void __launch_bounds__(256) gather_kernel(const uint8_t *__restrict__ buf, uint64_t n_rows, uint64_t magic,
                                                     uint64_t probes_per_thread, uint64_t *sink)
{
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t acc = 0;
    for (uint64_t it = 0; it < probes_per_thread; it += 8) {
        uint64_t v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            uint64_t row = fast_mod(mix64(tid * 0x100000001B3ULL + it + u), n_rows, magic);
            const uint8_t *p = buf + row * ROWB;
            if constexpr (ROWB == 8) v[u] = __ldg(reinterpret_cast<const uint64_t *>(p));
            else if constexpr (ROWB == 16) { ulonglong2 t = __ldg(reinterpret_cast<const ulonglong2 *>(p)); v[u] = t.x ^ t.y; }
            else {
                uint64_t a = 0;
#pragma unroll
                for (int i = 0; i < ROWB / 16; ++i) {
                    ulonglong2 t = __ldg(reinterpret_cast<const ulonglong2 *>(p) + i);
                    a ^= t.x ^ t.y;
                }
                v[u] = a;
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) acc ^= v[u];
    }
    if (acc == 0x123456789ABCDEFULL) sink[0] = acc;   // practically never; keeps the loads live
}


// Cooperative variant: G = ROWB / V adjacent lanes read ONE row with a single load instruction of V bytes
// per lane (V = 16: LDG.128, V = 32: LDG.256), so the L1 sees one request per row that names all of its
// sectors, instead of ROWB/16 separate requests from one thread as in gather_kernel.
template <int V>
__device__ __forceinline__ uint64_t load_fold(const uint8_t *p)
{
    if constexpr (V == 32) {
        uint64_t a, b, c, d;
        asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
        return a ^ b ^ c ^ d;
    } else {
        ulonglong2 t = __ldg(reinterpret_cast<const ulonglong2 *>(p));
        return t.x ^ t.y;
    }
}

template <int ROWB, int V>
__global__ void __launch_bounds__(256) gather_coop_kernel(const uint8_t *__restrict__ buf, uint64_t n_rows, uint64_t magic,
                                                          uint64_t probes_per_group, uint64_t *sink)
{
    constexpr int G = ROWB / V;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t gid = tid / G;
    const uint32_t sub = (uint32_t)(tid % G);
    uint64_t acc = 0;
    for (uint64_t it = 0; it < probes_per_group; it += 8) {
        uint64_t v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            uint64_t row = fast_mod(mix64(gid * 0x100000001B3ULL + it + u), n_rows, magic);
            v[u] = load_fold<V>(buf + row * ROWB + sub * V);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) acc ^= v[u];
    }
    if (acc == 0x123456789ABCDEFULL) sink[0] = acc;
}

// Synthetic reference generator (test / benchmark aid): base i of the stream `seed` is a pure function of (seed, i), so
// the host regenerates any window without holding the sequence (readbouncer_b200/synth.py: hash_bases).
//   h(b) = mix64(seed * 0xD1342543DE82EF95 + b), b = i >> 5;  code = (h >> 2 (i & 31)) & 3;  base = "ACGT"[code]
__global__ void __launch_bounds__(256) synth_bases_kernel(uint8_t *__restrict__ out, uint64_t n, uint64_t seed, uint64_t start)
{
    const uint64_t b_first = start >> 5, b_last = (start + n + 31) >> 5;          // hash blocks [b_first, b_last)
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t b = b_first + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < b_last; b += stride) {
        const uint64_t h = mix64(seed * 0xD1342543DE82EF95ULL + b);
        const uint64_t p0 = b << 5;
        uint32_t wds[8];
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            uint32_t x = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) x |= ((0x54474341u >> (8 * (uint32_t)((h >> (2 * (4 * w + j))) & 3u))) & 0xFFu) << (8 * j);
            wds[w] = x;
        }
        uint8_t *dst = out + (p0 - start);                                       // may lie before `out` for the first block
        if (p0 >= start && p0 + 32 <= start + n && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
            reinterpret_cast<uint4 *>(dst)[0] = make_uint4(wds[0], wds[1], wds[2], wds[3]);
            reinterpret_cast<uint4 *>(dst)[1] = make_uint4(wds[4], wds[5], wds[6], wds[7]);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const uint64_t p = p0 + j;
                if (p >= start && p < start + n) out[p - start] = (uint8_t)(wds[j >> 2] >> (8 * (j & 3)));
            }
        }
    }
}

}  // namespace rb

extern "C" RB_API int rb_synth_bases_dev(uint8_t *d_out, uint64_t n, uint64_t seed, uint64_t start, rb_stream stream)
{
    if (n == 0) return RB_OK;
    if (!d_out) return RB_ERR_INVALID_ARG;
    const uint64_t blocks = ((n + 31) / 32 + 1 + 255) / 256;
    rb::synth_bases_kernel<<<(unsigned)(blocks < 148u * 32u ? blocks : 148u * 32u), 256, 0, (cudaStream_t)stream>>>(d_out, n, seed, start);
    return cudaGetLastError() == cudaSuccess ? RB_OK : RB_ERR_CUDA;
}

extern "C" RB_API int rb_microbench_gather(const void *d_buf, uint64_t n_rows, uint32_t row_bytes,
                                           uint64_t probes_per_thread, uint32_t n_blocks, uint64_t *d_sink,
                                           rb_stream stream)
{
    if (!d_buf || !d_sink || n_rows == 0 || n_blocks == 0) return RB_ERR_INVALID_ARG;
    const uint64_t magic = rb::mod_magic(n_rows);
    cudaStream_t st = (cudaStream_t)stream;
    probes_per_thread = (probes_per_thread + 7) / 8 * 8;
    const uint8_t *b = static_cast<const uint8_t *>(d_buf);
    if (row_bytes == 8) rb::gather_kernel<8><<<n_blocks, 256, 0, st>>>(b, n_rows, magic, probes_per_thread, d_sink);
    else if (row_bytes == 16) rb::gather_kernel<16><<<n_blocks, 256, 0, st>>>(b, n_rows, magic, probes_per_thread, d_sink);
    else if (row_bytes == 32) rb::gather_kernel<32><<<n_blocks, 256, 0, st>>>(b, n_rows, magic, probes_per_thread, d_sink);
    else if (row_bytes == 64) rb::gather_kernel<64><<<n_blocks, 256, 0, st>>>(b, n_rows, magic, probes_per_thread, d_sink);
    else if (row_bytes == 128) rb::gather_kernel<128><<<n_blocks, 256, 0, st>>>(b, n_rows, magic, probes_per_thread, d_sink);
    else return RB_ERR_INVALID_ARG;
    return cudaGetLastError() == cudaSuccess ? RB_OK : RB_ERR_CUDA;
}

// rows of row_bytes read by row_bytes / lane_bytes adjacent lanes, one load instruction per row
extern "C" RB_API int rb_microbench_gather_coop(const void *d_buf, uint64_t n_rows, uint32_t row_bytes, uint32_t lane_bytes,
                                                uint64_t probes_per_group, uint32_t n_blocks, uint64_t *d_sink,
                                                rb_stream stream)
{
    if (!d_buf || !d_sink || n_rows == 0 || n_blocks == 0) return RB_ERR_INVALID_ARG;
    const uint64_t magic = rb::mod_magic(n_rows);
    cudaStream_t st = (cudaStream_t)stream;
    probes_per_group = (probes_per_group + 7) / 8 * 8;
    const uint8_t *b = static_cast<const uint8_t *>(d_buf);
#define RB_COOP(R, V)                                                                                              \
    if (row_bytes == R && lane_bytes == V) {                                                                       \
        rb::gather_coop_kernel<R, V><<<n_blocks, 256, 0, st>>>(b, n_rows, magic, probes_per_group, d_sink);        \
        return cudaGetLastError() == cudaSuccess ? RB_OK : RB_ERR_CUDA;                                            \
    }
    RB_COOP(16, 16) RB_COOP(32, 16) RB_COOP(32, 32) RB_COOP(64, 16) RB_COOP(64, 32) RB_COOP(128, 16) RB_COOP(128, 32)
    RB_COOP(256, 16) RB_COOP(256, 32) RB_COOP(512, 16)
#undef RB_COOP
    return RB_ERR_INVALID_ARG;
}
