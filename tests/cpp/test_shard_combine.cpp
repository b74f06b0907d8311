// What a C++ host does with a bin-sharded filter, and nothing but the C ABI:
//   (a) the whole filter on device 0                                   -> rb_ibf_count_batch
//   (b) n_dev x 2 bin shards spread over the devices, ONE process       -> rb_ibf_count_batch_sharded (keys folded over NVLink)
//   (c) one shard per device, one NCCL rank per device (needs >= 2 GPUs) -> rb_ibf_count_batch_dev + rb_keys_combine_nccl + decode
// All three must return the same max_count / hit / argmax_bin for every read and threshold.
//   test_shard_combine [n_reads]
#include "rb_ibf.h"

#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define CHECK(cond)                                                                                                      \
    do {                                                                                                                 \
        if (!(cond)) { std::fprintf(stderr, "CHECK failed %s:%d: %s (%s)\n", __FILE__, __LINE__, #cond, rb_last_error()); std::exit(1); } \
    } while (0)

static uint64_t state = 88172645463325252ULL;
static uint32_t rnd() { state ^= state << 13; state ^= state >> 7; state ^= state << 17; return (uint32_t)(state >> 20); }

int main(int argc, char **argv)
{
    const uint64_t n_reads = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 3000;
    int n_dev = rb_device_count();
    CHECK(n_dev >= 1);
    if (n_dev > 4) n_dev = 4;
    // reference: 1 400 fragments of 2 000 bases -> 1 401 bins (22 row words), k = 13
    const uint32_t k = 13;
    const uint64_t frag = 2000, seqlen = 1400 * frag + 7;
    std::string ref(seqlen, 'A');
    for (auto &c : ref) c = "ACGT"[rnd() & 3];
    const uint64_t n_frags = rb_fragment_schedule(seqlen, frag, k, nullptr, nullptr, 0);
    std::vector<uint64_t> fb(n_frags), fe(n_frags), fbin(n_frags);
    rb_fragment_schedule(seqlen, frag, k, fb.data(), fe.data(), n_frags);
    for (uint64_t i = 0; i < n_frags; ++i) fbin[i] = i;
    const uint64_t n_bins = seqlen / frag + 1;
    CHECK(n_bins == n_frags);
    const uint64_t n_bits = rb_ibf_size_bits(frag, k, 3, 0.01, n_bins);
    // reads: 250-base windows of the reference with ~8 % substitutions, or random
    std::string bases;
    std::vector<uint64_t> off{0};
    for (uint64_t r = 0; r < n_reads; ++r) {
        const uint64_t len = 100 + rnd() % 400;
        if (rnd() & 1) {
            const uint64_t p = rnd() % (seqlen - len);
            for (uint64_t i = 0; i < len; ++i) bases += (rnd() % 100 < 8) ? "ACGT"[rnd() & 3] : ref[p + i];
        } else
            for (uint64_t i = 0; i < len; ++i) bases += "ACGT"[rnd() & 3];
        off.push_back(bases.size());
    }
    std::vector<uint16_t> luts(2 * 65536);
    CHECK(rb_threshold_lut(0.1, 0.95, k, luts.data()) == RB_OK);
    CHECK(rb_threshold_lut(0.08, 0.95, k, luts.data() + 65536) == RB_OK);
    const uint64_t nk = 2 * n_reads;
    int st = 0;

    // (a) whole filter
    rb_ibf *whole = rb_ibf_create(n_bins, 3, k, n_bits, 0, &st);
    CHECK(whole && st == RB_OK);
    CHECK(rb_ibf_insert_batch(whole, ref.data(), ref.size(), fb.data(), fe.data(), fbin.data(), n_frags, nullptr) == RB_OK);
    std::vector<uint16_t> mx_a(nk), mx_b(nk), mx_c(nk);
    std::vector<uint8_t> hit_a(nk), hit_b(nk), hit_c(nk), fl_a(n_reads), fl_b(n_reads);
    std::vector<uint32_t> am_a(nk), am_b(nk), am_c(nk);
    CHECK(rb_ibf_count_batch(whole, bases.data(), off.data(), n_reads, luts.data(), 2, nullptr, nullptr, mx_a.data(), hit_a.data(),
                             am_a.data(), fl_a.data(), nullptr) == RB_OK);
    uint64_t hits = 0;
    for (uint64_t i = 0; i < n_reads; ++i) hits += hit_a[i];
    CHECK(hits > n_reads / 4 && hits < n_reads);

    // (b) 2 shards per device, one process
    const int n_shards = 2 * n_dev;
    std::vector<rb_ibf *> shards(n_shards);
    for (int s = 0; s < n_shards; ++s) {
        shards[s] = rb_ibf_create_shard(n_bins, 3, k, n_bits, s % n_dev, s, n_shards, &st);
        CHECK(shards[s] && st == RB_OK);
        CHECK(rb_ibf_insert_batch(shards[s], ref.data(), ref.size(), fb.data(), fe.data(), fbin.data(), n_frags, nullptr) == RB_OK);
    }
    for (int round = 0; round < 2; ++round) {                        // round 1: with the shards' postings tables built
        CHECK(rb_ibf_count_batch_sharded(shards.data(), n_shards, bases.data(), off.data(), n_reads, luts.data(), 2, mx_b.data(),
                                         hit_b.data(), am_b.data(), fl_b.data()) == RB_OK);
        CHECK(mx_a == mx_b && hit_a == hit_b && am_a == am_b && fl_a == fl_b);
        if (round == 0) CHECK(rb_ibf_enable_kmer_tables(shards.data(), n_shards, 0, nullptr) == RB_OK);
    }
    std::printf("sharded call OK: %d shards on %d device(s), %llu reads, %llu hits\n", n_shards, n_dev, (unsigned long long)n_reads,
                (unsigned long long)hits);

    // (c) one shard per device + NCCL (single process, one communicator per device, grouped calls)
    if (n_dev >= 2) {
        std::vector<ncclComm_t> comms(n_dev);
        std::vector<int> devs(n_dev);
        for (int d = 0; d < n_dev; ++d) devs[d] = d;
        CHECK(ncclCommInitAll(comms.data(), n_dev, devs.data()) == ncclSuccess);
        std::vector<rb_ibf *> sh(n_dev);
        std::vector<cudaStream_t> streams(n_dev);
        std::vector<uint8_t *> d_bases(n_dev);
        std::vector<uint64_t *> d_off(n_dev), d_keys(n_dev);
        std::vector<uint16_t *> d_lut(n_dev);
        for (int d = 0; d < n_dev; ++d) {
            sh[d] = rb_ibf_create_shard(n_bins, 3, k, n_bits, d, d, n_dev, &st);
            CHECK(sh[d] && st == RB_OK);
            CHECK(rb_ibf_insert_batch(sh[d], ref.data(), ref.size(), fb.data(), fe.data(), fbin.data(), n_frags, nullptr) == RB_OK);
            CHECK(cudaSetDevice(d) == cudaSuccess);
            CHECK(cudaStreamCreate(&streams[d]) == cudaSuccess);
            CHECK(cudaMalloc(&d_bases[d], bases.size() + 32) == cudaSuccess && cudaMalloc(&d_off[d], off.size() * 8) == cudaSuccess);
            CHECK(cudaMalloc(&d_keys[d], nk * 8) == cudaSuccess && cudaMalloc(&d_lut[d], luts.size() * 2) == cudaSuccess);
            CHECK(cudaMemcpy(d_bases[d], bases.data(), bases.size(), cudaMemcpyHostToDevice) == cudaSuccess);
            CHECK(cudaMemcpy(d_off[d], off.data(), off.size() * 8, cudaMemcpyHostToDevice) == cudaSuccess);
            CHECK(cudaMemcpy(d_lut[d], luts.data(), luts.size() * 2, cudaMemcpyHostToDevice) == cudaSuccess);
            CHECK(rb_ibf_count_batch_dev(sh[d], d_bases[d], d_off[d], n_reads, 0, d_lut[d], 2, d_keys[d], nullptr, nullptr, nullptr,
                                         streams[d]) == RB_OK);
        }
        CHECK(ncclGroupStart() == ncclSuccess);
        for (int d = 0; d < n_dev; ++d) {
            CHECK(cudaSetDevice(d) == cudaSuccess);
            CHECK(rb_keys_combine_nccl(comms[d], d_keys[d], nk, streams[d]) == RB_OK);
        }
        CHECK(ncclGroupEnd() == ncclSuccess);
        for (int d = 0; d < n_dev; ++d) {                            // every rank ends up with the whole filter's keys
            CHECK(cudaSetDevice(d) == cudaSuccess);
            std::vector<uint64_t> keys(nk);
            CHECK(cudaStreamSynchronize(streams[d]) == cudaSuccess);
            CHECK(cudaMemcpy(keys.data(), d_keys[d], nk * 8, cudaMemcpyDeviceToHost) == cudaSuccess);
            for (uint64_t i = 0; i < nk; ++i) {
                mx_c[i] = RB_KEY_MAX_COUNT(keys[i]); hit_c[i] = RB_KEY_HIT(keys[i]); am_c[i] = RB_KEY_ARGMAX_BIN(keys[i]);
            }
            CHECK(mx_a == mx_c && hit_a == hit_c && am_a == am_c);
        }
        for (int d = 0; d < n_dev; ++d) { rb_ibf_free(sh[d]); ncclCommDestroy(comms[d]); }
        std::printf("nccl combine OK: %d ranks\n", n_dev);
    } else
        std::printf("nccl combine skipped: 1 device\n");
    for (rb_ibf *s : shards) rb_ibf_free(s);
    rb_ibf_free(whole);
    std::puts("shard combine OK");
    return 0;
}
