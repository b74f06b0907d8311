// Exercises include/rb_interleave.hpp the way ReadBouncer's own gtests exercise src/IBF
// (src/test/libIBFTests/read.hpp, createfilter.hpp): same inputs, same expected values.
//   test_shim host                       -- host-only checks (no GPU needed)
//   test_shim nogpu <test.ibf>           -- expects a loud failure without a CUDA device
//   test_shim gpu <test.ibf> <test1.ibf> <test.fasta> <tmpdir> <read354>
#include "rb_interleave.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

#define CHECK(cond)                                                                        \
    do {                                                                                   \
        if (!(cond)) { std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #cond); std::exit(1); } \
    } while (0)

template <class E, class F> static bool throws(F f)
{
    try { f(); } catch (const E &) { return true; } catch (...) { return false; }
    return false;
}

static int host_checks()
{
    using namespace interleave;
    TInterval ci = calculateCI(0.1, 13, 35, 0.95);                 // read.hpp:156-157
    CHECK(ci.first == 5 && ci.second == 30);
    CHECK(rb_ibf_size_bits(100000, 13, 3, 0.01, 2) == 79121216);   // createfilter.hpp:148
    IBFConfig cfg;                                                 // ibfconfigtest.hpp:32-59
    CHECK(IBFConfig::MBinBits == 8388608 && cfg.overlap_length == 1500 && cfg.kmer_size == 13 && cfg.hash_functions == 3);
    CHECK(cfg.threads == 2 && cfg.n_refs == 400 && cfg.n_batches == 500000 && cfg.max_fp == 0.01);
    CHECK(cfg.validate() && cfg.threads_build == 1);
    cfg.threads = 5; cfg.validate(); CHECK(cfg.threads_build == 4);
    const std::vector<uint16_t> &lut = threshold_lut(0.1, 0.95, 13);
    CHECK(lut[35] == 65529 && lut[250] == 18 && lut[354] == 36);
    std::string in = "AAAAAAAACCCCCCCCCGAGAGAGGAGAGAGGAGAGAGAGAGCCCCAAAAGAGAGGAGATTTTANNNNNNNNTATATTATA";
    std::string out(in.size(), '\0');
    out.resize(rb_cut_out_nnns(in.data(), in.size(), out.data()));
    CHECK(out == "AAAAAAAACCCCCCCCCGAGAGAGGAGAGAGGAGAGAGAGAGCCCCAAAAGAGAGGAGATTTTATATATTAT");   // createfilter.hpp:135
    std::vector<TIbf> none;
    std::vector<IBFMeta> none_meta;
    ClassifyConfig cc;
    Read r("r", "ACGTACGTACGTACGTACGT");
    CHECK(throws<NullFilterException>([&] { r.classify(none, cc); }));                       // read.hpp:188
    CHECK(throws<NullFilterException>([&] { r.classify(none_meta, cc); }));                  // read.hpp:208
    CHECK(throws<NullFilterException>([&] { r.classify(none_meta, none_meta, cc); }));       // read.hpp:241
    IBF ibf;
    IBFConfig empty;
    CHECK(throws<MissingIBFFileException>([&] { ibf.load_filter(empty); }));
    CHECK(throws<MissingReferenceFilesException>([&] { ibf.create_filter(empty); }));
    std::puts("host OK");
    return 0;
}

int main(int argc, char **argv)
{
    using namespace interleave;
    if (argc < 2) return 2;
    if (!std::strcmp(argv[1], "host")) return host_checks();
    if (!std::strcmp(argv[1], "nogpu")) {
        IBF ibf;
        IBFConfig cfg;
        cfg.input_filter_file = argv[2];
        bool failed = false;
        try { ibf.load_filter(cfg); } catch (const IBFException &e) { failed = std::strstr(e.what(), "no CUDA device") != nullptr; }
        CHECK(failed);       // no silent CPU fallback
        std::puts("nogpu OK");
        return 0;
    }
    CHECK(argc >= 7);
    const std::string ibf0 = argv[2], ibf1 = argv[3], fasta = argv[4], tmp = argv[5], read354 = argv[6];
    ClassifyConfig config;
    config.error_rate = 0.1; config.significance = 0.95;
    std::vector<IBFMeta> filters;
    std::vector<TIbf> IBFs;
    for (const std::string &file : {ibf0, ibf1}) {
        IBF f;
        IBFConfig c;
        c.input_filter_file = file;
        FilterStats stats = f.load_filter(c);
        CHECK(c.kmer_size == 13);
        IBFMeta m; m.filter = f.getFilter(); m.name = file;
        CHECK(stats.totalBinsFile == m.filter.noOfBins);
        filters.push_back(m);
        IBFs.push_back(f.getFilter());
    }
    CHECK(IBFs[0].noOfBins == 2 && IBFs[1].noOfBins == 4 && IBFs[0].kmerSize == 13);      // read.hpp:199
    Read read("read354", read354);
    CHECK(read.getReadLength() == 354);                                                      // read.hpp:200
    CHECK(read.classify(IBFs, config) == true);                                              // read.hpp:202
    CHECK(read.count_matches(filters[0], config) == 282);                                    // read.hpp:221-229
    CHECK(read.count_matches(filters[1], config) == 182);
    CHECK(read.classify(filters, config) == 0);                                              // read.hpp:231
    std::vector<IBFMeta> v1{filters[0]}, v2{filters[1]};
    std::pair<int, int> p = read.classify(v1, v2, config);
    CHECK(p.first == 282 && p.second == 182);                                                // read.hpp:250
    Read shorty("s", "ACGT");
    CHECK(throws<ShortReadException>([&] { shorty.classify(IBFs, config); }));
    CHECK(throws<ShortReadException>([&] { shorty.classify(filters, config); }));
    p = shorty.classify(v1, v2, config);
    CHECK(p.first == 0 && p.second == 0);
    Read r35("r35", "AAAAAAACCCCCCCCCGAGAGAGGAGAGAGGAGAG");
    CHECK(r35.count_matches(filters[0], config) == 0);       // thr -7 wraps to 65529 (quirk Q6)
    std::vector<IBFMeta> nof;
    CHECK(check_unblock(read, config, v1, nof) == 1);
    CHECK(check_unblock(read, config, nof, v1) == 2);
    CHECK(check_unblock(read, config, v1, v2) == 0);
    CHECK(check_unblock(r35, config, v1, nof) == 0 && check_unblock(r35, config, nof, v1) == 1);
    // batch decisions equal the per-read decisions
    std::string cat = read354 + r35.sequence + "ACGT";
    uint64_t off[4] = {0, 354, 354 + 35, 354 + 35 + 4};
    std::vector<uint8_t> d = check_unblock_batch(cat.data(), off, 3, config, v1, v2);
    CHECK(d[0] == 0 && d[1] == 0 && d[2] == 0);
    d = check_unblock_batch(cat.data(), off, 3, config, v1, nof);
    CHECK(d[0] == 1 && d[1] == 0 && d[2] == 255);
    d = check_unblock_batch(cat.data(), off, 3, config, nof, v1);
    CHECK(d[0] == 2 && d[1] == 1 && d[2] == 255);

    // create_filter: production path inserts each sequence once (bins = 1 for test.fasta)
    IBF builder;
    IBFConfig bc;
    bc.reference_files.push_back(fasta);
    bc.output_filter_file = tmp + "/shim_built.ibf";
    bc.kmer_size = 13; bc.fragment_length = 100000;
    FilterStats bs = builder.create_filter(bc);
    CHECK(bs.totalBinsBinId == 1 && bs.sumSeqLen == 72 && bs.invalidSeqs == 0 && bs.totalSeqsFile == 1);
    CHECK(bc.filter_size_bits == 79121216);
    IBF reload;
    IBFConfig rc;
    rc.input_filter_file = bc.output_filter_file;
    CHECK(reload.load_filter(rc).totalBinsFile == 1);
    IBFMeta built; built.filter = reload.getFilter();
    CHECK(read.count_matches(built, config) == 282);
    // update_filter (IBFBuild.cpp:223-321): append the same reference as a new bin of the stored filter
    IBF updater;
    IBFConfig uc;
    uc.update_filter_file = bc.output_filter_file;
    uc.reference_files.push_back(fasta);
    uc.fragment_length = 100000;
    FilterStats us = updater.update_filter(uc);
    CHECK(us.totalBinsFile == 1 && us.newBins == 1 && us.totalBinsBinId == 2 && uc.kmer_size == 13);
    IBF reload2;
    CHECK(reload2.load_filter(rc).totalBinsFile == 2);
    IBFMeta grown; grown.filter = reload2.getFilter();
    CHECK(read.count_matches(grown, config) == 282);
    {
        const uint64_t off1[2] = {0, read354.size()};
        std::vector<uint16_t> lut(threshold_lut(0.1, 0.95, 13));
        std::vector<uint16_t> cf(2), cr(2);
        CHECK(rb_ibf_count_batch(grown.filter.get(), read354.data(), off1, 1, lut.data(), 1, cf.data(), cr.data(), nullptr, nullptr, nullptr, nullptr, nullptr) == RB_OK);
        CHECK(cf[0] == 282 && cf[1] == 282 && cr[0] == 0 && cr[1] == 0);      // both bins now hold the sequence
    }
    IBFConfig bad;
    bad.input_filter_file = fasta;
    CHECK(throws<ParseIBFFileException>([&] { reload.load_filter(bad); }));     // configReader.cpp:210-224 sniffing
    std::puts("gpu OK");
    return 0;
}
