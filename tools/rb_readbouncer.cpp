// rb_readbouncer -- `ReadBouncer --config x.toml` for the GPU-backed usages ("build", "classify").
// Same single CLI flag as the reference (src/main/parser.hpp:13-39); see include/rb_drivers.hpp.
#include "rb_drivers.hpp"

#include <cstring>

int main(int argc, char **argv)
{
    std::string cfg;
    bool print_only = false;
    for (int i = 1; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--config") && i + 1 < argc) cfg = argv[++i];
        else if (!std::strcmp(argv[i], "--print-config")) print_only = true;
    }
    if (cfg.empty()) { std::cerr << "usage: rb_readbouncer --config <file.toml> [--print-config]" << std::endl; return 2; }
    try {
        rbdrv::ConfigReader c = rbdrv::ConfigReader::from_toml(cfg);
        if (print_only) {
            const rbdrv::IBF_Params &p = c.IBF_Parsed;
            std::cout << "usage=" << c.usage << " output_directory=" << c.output_dir.string() << " kmer_size=" << p.size_k
                      << " fragment_size=" << p.fragment_size << " threads=" << p.threads << " exp_seq_error_rate=" << p.error_rate
                      << " chunk_length=" << p.chunk_length << " max_chunks=" << p.max_chunks << " targets=" << p.target_files.size()
                      << " depletes=" << p.deplete_files.size() << " reads=" << p.read_files.size() << std::endl;
            return 0;
        }
        rbdrv::ClassificationResults r;
        int rc = rbdrv::run_program(c, &r);
        if (rc == 0 && c.usage == "classify")
            std::cout << "RESULT found=" << r.found << " failed=" << r.failed << " too_short=" << r.too_short
                      << " reads=" << r.readCounter << " avg_classify_s=" << r.avgClassifyduration << " read_file_s=" << r.readFileSeconds
                      << " table_setup_s=" << r.tableSetupSeconds << std::endl;
        return rc;
    } catch (const std::exception &e) {
        std::cerr << "[Error] " << e.what() << std::endl;
        return 1;
    }
}
