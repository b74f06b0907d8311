// rb_live_bench -- the usage="target" shape (src/main/adaptive_sampling.hpp:214-356) without a sequencer: micro-batches of
// basecalled 250-base chunks go through rblive::LiveClassifier (check_unblock against depletion + target filters, both
// thresholds of the retry in one pass per filter), and the time from "batch handed over" to "decisions back" is measured.
//   rb_live_bench <n_target_filters> <n_deplete_filters> <genome_len> <batches> <batch_size...>
// Filters are built from synthetic genomes (fragment_size 100 000, k = 13); chunks: half from the genomes with 8 % errors.
// Prints one JSON line per batch size.
#include "rb_live.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>

static std::string random_genome(size_t n, uint64_t seed)
{
    std::mt19937_64 rng(seed);
    std::string s(n, 'A');
    for (size_t i = 0; i < n; i += 32) {
        uint64_t r = rng();
        for (size_t j = i; j < std::min(n, i + 32); ++j, r >>= 2) s[j] = "ACGT"[r & 3];
    }
    return s;
}

int main(int argc, char **argv)
{
    using namespace interleave;
    if (argc < 6) { std::fprintf(stderr, "usage: rb_live_bench n_target n_deplete genome_len batches batch_size...\n"); return 2; }
    const int n_tgt = std::atoi(argv[1]), n_dep = std::atoi(argv[2]);
    const size_t glen = std::strtoull(argv[3], nullptr, 10);
    const int batches = std::atoi(argv[4]);
    std::vector<std::string> genomes;
    std::vector<IBFMeta> tgt, dep;
    try {
        for (int g = 0; g < n_tgt + n_dep; ++g) {
            genomes.push_back(random_genome(glen, 1000 + g));
            const std::string &s = genomes.back();
            const uint64_t n_frags = rb_fragment_schedule(s.size(), 100000, 13, nullptr, nullptr, 0);
            std::vector<uint64_t> fb(n_frags), fe(n_frags), fbin(n_frags);
            rb_fragment_schedule(s.size(), 100000, 13, fb.data(), fe.data(), n_frags);
            for (uint64_t i = 0; i < n_frags; ++i) fbin[i] = i;
            const uint64_t bins = s.size() / 100000 + 1;
            int st = 0;
            rb_ibf *h = rb_ibf_create(bins, 3, 13, rb_ibf_size_bits(100000, 13, 3, 0.01, bins), 0, &st);
            if (!h) throw_status(st, "create");
            st = rb_ibf_insert_batch(h, s.data(), s.size(), fb.data(), fe.data(), fbin.data(), std::min<uint64_t>(n_frags, bins), nullptr);
            if (st != RB_OK) throw_status(st, "insert");
            IBFMeta m;
            m.filter = TIbf(h);
            m.name = "g" + std::to_string(g);
            (g < n_tgt ? tgt : dep).push_back(std::move(m));
        }
        ClassifyConfig conf;
        conf.error_rate = 0.1;
        conf.significance = 0.95;
        const auto t_setup = std::chrono::steady_clock::now();
        rblive::LiveClassifier live(dep, tgt, conf);            // plans and builds the k-mer tables of all filters
        const double setup_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_setup).count();
        std::mt19937_64 rng(7);
        uint64_t next_id = 0;
        for (int a = 5; a < argc; ++a) {
            const size_t bs = std::strtoull(argv[a], nullptr, 10);
            std::vector<double> lat;
            size_t decisions = 0, unblocked = 0, stopped = 0;
            for (int b = 0; b < batches + 3; ++b) {
                std::vector<rblive::LiveRead> batch(bs);
                for (size_t i = 0; i < bs; ++i) {
                    rblive::LiveRead &r = batch[i];
                    r.id = "read" + std::to_string(next_id++);
                    r.channelNr = (uint32_t)(i % 3000);
                    r.readNr = (uint32_t)next_id;
                    if (rng() & 1) {
                        const std::string &g = genomes[rng() % genomes.size()];
                        const size_t p = rng() % (g.size() - 250);
                        r.sequence = g.substr(p, 250);
                        for (char &c : r.sequence) if (rng() % 100 < 8) c = "ACGT"[rng() & 3];
                    } else {
                        r.sequence.resize(250);
                        for (char &c : r.sequence) c = "ACGT"[rng() & 3];
                    }
                }
                const auto t0 = std::chrono::steady_clock::now();
                std::vector<rblive::LiveDecision> d = live.classify_batch(std::move(batch));
                const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
                if (b >= 3) {
                    lat.push_back(us);
                    decisions += d.size();
                    for (const auto &x : d) { unblocked += x.action == rblive::kUnblock; stopped += x.action == rblive::kStopReceiving; }
                }
            }
            std::sort(lat.begin(), lat.end());
            const double med = lat[lat.size() / 2], p90 = lat[(size_t)(lat.size() * 0.9)];
            std::printf("{\"batch\": %zu, \"target_filters\": %d, \"deplete_filters\": %d, \"median_us\": %.1f, \"p90_us\": %.1f, "
                        "\"chunks_per_s_at_median\": %.0f, \"decisions\": %zu, \"unblock\": %zu, \"stop_receiving\": %zu, "
                        "\"pending_reads\": %zu, \"table_setup_ms\": %.1f}\n",
                        bs, n_tgt, n_dep, med, p90, bs / med * 1e6, decisions, unblocked, stopped, live.pending(), setup_ms);
            std::fflush(stdout);
        }
    } catch (const std::exception &e) {
        std::fprintf(stderr, "[Error] %s\n", e.what());
        return 1;
    }
    return 0;
}
