// rb_drivers.hpp -- batched C++ drivers for `usage = "build"` and `usage = "classify"` on top of
// rb_interleave.hpp.  Header-only.
//
// Mirrors, with the reference's names and TOML keys:
//   ConfigReader::readIBF defaults           src/config/configReader.cpp:232-345
//   run_program / buildIBF / getIBF          src/main/main.cpp:274-406, src/main/ibfbuild.hpp:21-182
//   classify_reads, classify_deplete_target, fragment_start/end, ClassificationResults
//                                            src/main/classify.hpp:58-380
// What changes: instead of one read at a time (one std::async per filter per read,
// src/IBF/IBFClassify.cpp:259), chunk i of ALL still-unclassified reads is classified in one GPU pass
// per filter; outputs are written afterwards in file order, so files and counters equal the serial loop's.
#pragma once

#include "rb_interleave.hpp"

#include <chrono>
#include <cstdio>
#include <filesystem>
#include <iostream>
#include <sstream>

namespace rbdrv {

// ---- the [IBF] table of the TOML config (src/config/configReader.hpp:55-66, defaults configReader.cpp:238-243) ----
struct IBF_Params {
    int size_k = 13;
    int fragment_size = 100000;
    int threads = 1;
    double error_rate = 0.1;
    int chunk_length = 250;
    int max_chunks = 5;
    std::vector<std::filesystem::path> target_files, deplete_files, read_files;
    int device = 0;                          // new optional key: [IBF] device
};

struct ConfigReader {
    std::string usage = "classify";
    std::filesystem::path output_dir = "RB_out", log_dir = "RB_out/logs";
    IBF_Params IBF_Parsed;

    class ConfigReaderException : public std::runtime_error { using std::runtime_error::runtime_error; };

    // ConfigReader::filterException (src/config/configReader.cpp:210-224): true iff the file loads as an IBF
    static bool filterException(const std::filesystem::path &file)
    {
        // the reference attempts seqan::retrieve; here the header/metadata validation of rb_ibf_load is
        // replicated on the host so that sniffing does not allocate device memory
        std::ifstream in(file, std::ios::binary);
        if (!in) return false;
        uint64_t bit_len = 0;
        in.read(reinterpret_cast<char *>(&bit_len), 8);
        std::error_code ec;
        const uint64_t fsize = std::filesystem::file_size(file, ec);
        if (!in || ec || bit_len < 320 || (bit_len % 64) || fsize != 8 + bit_len / 8) return false;
        uint64_t tail[4];
        in.seekg((std::streamoff)(8 + (bit_len - 256) / 8));
        in.read(reinterpret_cast<char *>(tail), 32);
        return in && tail[0] > 0 && tail[1] > 0 && tail[1] <= 16 && tail[2] > 0 && tail[2] <= 32 && tail[3] == tail[2];
    }

    // Minimal TOML subset: [tables], key = int | float | "str" | 'str' | [ "a", 'b', ... ] (arrays may span lines), # comments
    static ConfigReader from_toml(const std::string &path)
    {
        std::ifstream in(path);
        if (!in) throw ConfigReaderException("cannot open config file " + path);
        ConfigReader c;
        std::string line, table, pending_key, pending_val;
        auto strip = [](std::string s) {
            bool q1 = false, q2 = false;
            for (size_t i = 0; i < s.size(); ++i) {
                if (s[i] == '\'' && !q2) q1 = !q1;
                else if (s[i] == '"' && !q1) q2 = !q2;
                else if (s[i] == '#' && !q1 && !q2) { s.erase(i); break; }
            }
            size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
            return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
        };
        auto unquote = [](std::string v) {
            if (v.size() >= 2 && (v.front() == '"' || v.front() == '\'') && v.back() == v.front()) return v.substr(1, v.size() - 2);
            return v;
        };
        auto list = [&](const std::string &v) {
            std::vector<std::filesystem::path> out;
            std::string cur; char q = 0;
            for (char ch : v) {
                if (q) { if (ch == q) { out.emplace_back(cur); cur.clear(); q = 0; } else cur += ch; }
                else if (ch == '"' || ch == '\'') q = ch;
            }
            return out;
        };
        auto assign = [&](const std::string &key, const std::string &val) {
            try {
                if (table.empty()) {
                    if (key == "usage") c.usage = unquote(val);
                    else if (key == "output_directory") c.output_dir = unquote(val);
                    else if (key == "log_directory") c.log_dir = unquote(val);
                } else if (table == "IBF") {
                    IBF_Params &p = c.IBF_Parsed;
                    if (key == "kmer_size") p.size_k = std::stoi(val);
                    else if (key == "fragment_size") p.fragment_size = std::stoi(val);
                    else if (key == "threads") p.threads = std::stoi(val);
                    else if (key == "exp_seq_error_rate") p.error_rate = std::stod(val);
                    else if (key == "chunk_length") p.chunk_length = std::stoi(val);
                    else if (key == "max_chunks") p.max_chunks = std::stoi(val);
                    else if (key == "device") p.device = std::stoi(val);
                    else if (key == "target_files") p.target_files = list(val);
                    else if (key == "deplete_files") p.deplete_files = list(val);
                    else if (key == "read_files") p.read_files = list(val);
                }   // [MinKNOW] and [Basecaller] stay with the host application
            } catch (const std::exception &) {
                throw ConfigReaderException("bad value for " + key + ": " + val);
            }
        };
        while (std::getline(in, line)) {
            line = strip(line);
            if (line.empty()) continue;
            if (!pending_key.empty()) {                          // continuation of a multi-line array
                pending_val += " " + line;
                if (line.find(']') != std::string::npos) { assign(pending_key, pending_val); pending_key.clear(); }
                continue;
            }
            if (line.front() == '[' && line.back() == ']' && line.find('=') == std::string::npos) { table = strip(line.substr(1, line.size() - 2)); continue; }
            size_t eq = line.find('=');
            if (eq == std::string::npos) continue;
            std::string key = strip(line.substr(0, eq)), val = strip(line.substr(eq + 1));
            if (!val.empty() && val.front() == '[' && val.find(']') == std::string::npos) { pending_key = key; pending_val = val; continue; }
            assign(key, val);
        }
        IBF_Params &p = c.IBF_Parsed;
        if (c.usage != "test" && p.deplete_files.size() + p.target_files.size() == 0)
            throw ConfigReaderException("[Error] At least one target or deplete file has to be specified!");
        for (auto &f : p.target_files) if (!std::filesystem::exists(f)) throw ConfigReaderException("[Error] The following target file does not exist: " + f.string());
        for (auto &f : p.deplete_files) if (!std::filesystem::exists(f)) throw ConfigReaderException("[Error] The following deplete file does not exist: " + f.string());
        for (auto &f : p.read_files) if (!std::filesystem::exists(f)) throw ConfigReaderException("[Error] The following read file does not exist: " + f.string());
        if (c.usage == "classify" && p.read_files.empty()) throw ConfigReaderException("[Error] read_files missing for usage classify");
        return c;
    }
};

// ---- build (src/main/ibfbuild.hpp:21-59, print_build_stats src/IBF/IBFBuild.cpp:558-571) ---------------------------------
inline interleave::TIbf buildIBF(const ConfigReader &config_reader, const std::string &reference_file,
                                 const std::string &bloom_filter_output_path, interleave::FilterStats *stats_out = nullptr)
{
    interleave::IBFConfig config{};
    config.reference_files.emplace_back(reference_file);
    config.output_filter_file = bloom_filter_output_path;
    config.kmer_size = (uint16_t)config_reader.IBF_Parsed.size_k;
    config.threads_build = (uint16_t)config_reader.IBF_Parsed.threads;
    config.fragment_length = (uint64_t)config_reader.IBF_Parsed.fragment_size;
    config.device = config_reader.IBF_Parsed.device;
    interleave::IBF filter{};
    interleave::FilterStats stats = filter.create_filter(config);
    const uint64_t valid = stats.totalSeqsFile - stats.invalidSeqs;
    std::cerr << "IBF-build processed " << valid << " sequences (" << stats.sumSeqLen / 1000000.0 << " Mbp)" << std::endl;
    if (stats.invalidSeqs > 0) std::cerr << " - " << stats.invalidSeqs << " invalid sequences were skipped" << std::endl;
    std::cerr << " - " << valid << " sequences in " << stats.totalBinsFile + stats.newBins << " bins were written to the IBF" << std::endl;
    if (stats_out) *stats_out = stats;
    return filter.getFilter();
}

// getIBF (src/main/ibfbuild.hpp:69-182): load .ibf files, build FASTA files into <output_dir>/<stem>.ibf
inline std::vector<interleave::IBFMeta> getIBF(const ConfigReader &config, bool depleteFilter, bool targetFilter)
{
    std::vector<interleave::IBFMeta> out;
    const auto &files = depleteFilter ? config.IBF_Parsed.deplete_files : config.IBF_Parsed.target_files;
    if (!depleteFilter && !targetFilter) return out;
    for (const std::filesystem::path &file : files) {
        interleave::IBFMeta filter{};
        filter.name = file.stem().string();
        if (ConfigReader::filterException(file)) {
            interleave::IBF tf{};
            interleave::IBFConfig cfg{};
            cfg.input_filter_file = file.string();
            cfg.device = config.IBF_Parsed.device;
            interleave::FilterStats stats = tf.load_filter(cfg);
            filter.filter = tf.getFilter();
            std::cerr << stats.totalBinsFile << " bins were loaded from the IBF" << std::endl;
        } else {
            std::filesystem::path o = config.output_dir / file.filename();
            o.replace_extension("ibf");
            filter.filter = buildIBF(config, file.string(), o.string());
        }
        out.emplace_back(std::move(filter));
    }
    return out;
}

// usage = "build" (src/main/main.cpp:286-343)
inline int run_build(const ConfigReader &config)
{
    std::filesystem::create_directories(config.output_dir);
    for (const auto *files : {&config.IBF_Parsed.target_files, &config.IBF_Parsed.deplete_files})
        for (const std::filesystem::path &file : *files) {
            if (!ConfigReader::filterException(file)) {
                std::cout << "The file: " << file.filename() << " is a fasta file, start building ibf ......." << '\n';
                std::filesystem::path o = config.output_dir / file.filename();
                o.replace_extension("ibf");
                buildIBF(config, file.string(), o.string());
                std::cout << '\n';
            } else {
                std::cout << "[INFO] The following file is an IBF file: " << file.string() << '\n';
            }
        }
    return 0;
}

// ---- classify (src/main/classify.hpp) -----------------------------------------------------------------------------------------
inline uint64_t fragment_start(int chunk_length, uint8_t i) { return (uint64_t)i * chunk_length; }          // classify.hpp:115
inline uint64_t fragment_end(int chunk_length, uint8_t i) { return (uint64_t)(i + 1) * chunk_length; }      // classify.hpp:121

struct ClassificationResults {          // classify.hpp:127-134
    uint64_t found = 0;
    uint16_t failed = 0;
    uint64_t too_short = 0;
    uint64_t readCounter = 0;
    std::vector<int> assignment;         // per read: target index, -2 depleted-mode hit, -1 unclassified, -3 failed, -4 too short
    double avgClassifyduration = 0;      // seconds per read: chunk loop + decisions + output records (classify.hpp:252-303, :362)
    double readFileSeconds = 0;          // parsing the read file (outside the reference's per-read timer)
    double tableSetupSeconds = 0;        // the joint k-mer table plan (one-off, before the first read)
};

inline std::string to_dna5_string(const std::string &s)
{
    std::string o(s.size(), 'N');
    for (size_t i = 0; i < s.size(); ++i) {
        switch (s[i] & 0xDF) { case 'A': o[i] = 'A'; break; case 'C': o[i] = 'C'; break; case 'G': o[i] = 'G'; break;
                               case 'T': case 'U': o[i] = 'T'; break; default: break; }
    }
    return o;
}

// classify_reads (classify.hpp:142-380), batched by chunk index.  Returns the results of the LAST read file,
// like the reference's ClassificationResults_ global.
inline ClassificationResults classify_reads(const ConfigReader &config, std::vector<interleave::IBFMeta> DepletionFilters,
                                            std::vector<interleave::IBFMeta> TargetFilters)
{
    using namespace interleave;
    const bool deplete = !DepletionFilters.empty(), target = !TargetFilters.empty();
    if (!deplete && !target) throw NullFilterException("[Error] No depletion or target filters have been provided for classification!");
    ClassifyConfig Conf{};
    Conf.significance = 0.95;                                  // classify.hpp:173
    Conf.error_rate = config.IBF_Parsed.error_rate;
    const int cl = config.IBF_Parsed.chunk_length;
    using clk = std::chrono::steady_clock;
    auto secs = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    const clk::time_point t_setup = clk::now();
    enable_kmer_tables(DepletionFilters, TargetFilters);      // one plan for all filters, before the first read
    const double table_setup_s = secs(t_setup, clk::now());
    for (const std::vector<IBFMeta> *v : {&TargetFilters, &DepletionFilters})
        for (const IBFMeta &f : *v) {
            rb_ibf_info_t info{};
            if (f.filter && rb_ibf_info(f.filter.get(), &info) == RB_OK)
                std::cerr << "k-mer table: " << f.name << " bins=" << info.n_bins << " kind=" << info.kmer_table_kind << " span="
                          << info.kmer_table_span << " bytes=" << info.kmer_table_bytes << std::endl;
        }
    ClassificationResults res;
    std::filesystem::create_directories(config.output_dir);

    for (const std::filesystem::path &read_file : config.IBF_Parsed.read_files) {
        res = ClassificationResults();
        res.tableSetupSeconds = table_setup_s;
        for (IBFMeta &f : TargetFilters) f.classified = 0;
        const clk::time_point t_read = clk::now();
        std::vector<SeqRecord> reads = read_sequence_file(read_file.string());
        const clk::time_point t_classify = clk::now();
        res.readFileSeconds = secs(t_read, t_classify);
        const size_t n = reads.size();
        res.readCounter = n;
        res.assignment.assign(n, -1);
        std::vector<size_t> active;
        for (size_t r = 0; r < n; ++r) {
            if (reads[r].seq.size() < (size_t)cl) { res.too_short++; res.assignment[r] = -4; }      // classify.hpp:247-250
            else active.push_back(r);
        }
        std::cout << '\n' << "Classification results of: " << read_file.string() << '\n' << '\n';

        for (int i = 0; i < config.IBF_Parsed.max_chunks && !active.empty(); ++i) {
            // chunk i of every still-unclassified read (classify.hpp:262-271)
            std::string bases;
            std::vector<uint64_t> off{0};
            std::vector<size_t> owner;
            for (size_t r : active) {
                const std::string &seq = reads[r].seq;
                const uint64_t b = fragment_start(cl, (uint8_t)i);
                if (b >= seq.size()) continue;                  // the reference's infix is undefined here (quirk Q9): no more chunks
                const uint64_t e = std::min<uint64_t>(fragment_end(cl, (uint8_t)i), seq.size());
                bases.append(seq, b, e - b);
                off.push_back(bases.size());
                owner.push_back(r);
            }
            const uint64_t m = owner.size();
            if (m == 0) break;
            // per-filter summaries of the whole batch: max_count at error_rate (and at error_rate-0.02 when both sets are given)
            // all filters of both sets (count_matches_batch_all)
            std::vector<const TIbf *> all;
            for (IBFMeta &f : TargetFilters) all.push_back(&f.filter);
            for (IBFMeta &f : DepletionFilters) all.push_back(&f.filter);
            const std::vector<BatchCounts> counts = count_matches_batch_all(all, bases.data(), off.data(), m, Conf, deplete && target);
            auto run = [&](size_t first, size_t count, std::vector<std::vector<uint16_t>> &cnt, std::vector<std::vector<uint16_t>> &cnt_s,
                           std::vector<uint8_t> &flag) {
                flag.assign(m, 0);
                for (size_t fi = first; fi < first + count; ++fi) {
                    const BatchCounts &c = counts[fi];
                    cnt.emplace_back(c.max_count.begin(), c.max_count.begin() + m);
                    if (deplete && target) cnt_s.emplace_back(c.max_count.begin() + m, c.max_count.begin() + 2 * m);
                    for (uint64_t j = 0; j < m; ++j) flag[j] = std::max(flag[j], c.read_flag[j]);
                }
            };
            std::vector<std::vector<uint16_t>> tc, tcs, dc, dcs;
            std::vector<uint8_t> tflag, dflag;
            if (target) run(0, TargetFilters.size(), tc, tcs, tflag);
            if (deplete) run(TargetFilters.size(), DepletionFilters.size(), dc, dcs, dflag);
            auto best_of = [&](const std::vector<std::vector<uint16_t>> &c, uint64_t j, int &idx) {
                uint64_t best = 0; idx = -1;
                for (size_t f = 0; f < c.size(); ++f) if (c[f][j] > best) { best = c[f][j]; idx = (int)f; }   // strictly greater, lowest index
                return best;
            };
            std::vector<size_t> still;
            for (uint64_t j = 0; j < m; ++j) {
                const size_t r = owner[j];
                int ti = -1, di = -1, tmp = -1;
                bool classified = false;
                if (deplete && target) {                        // classify_deplete_target, classify.hpp:58-111
                    if (tflag[j] >= 2 || dflag[j] >= 2) { res.failed++; res.assignment[r] = -3; continue; }   // chunk > 65 535 bases
                    const uint64_t t0 = best_of(tc, j, ti), d0 = best_of(dc, j, di);
                    if (t0 > 0) {
                        if (d0 > 0) {
                            const uint64_t t1 = best_of(tcs, j, tmp), d1 = best_of(dcs, j, tmp);
                            classified = t1 > 0 && d1 == 0;     // best target index still taken at the original error rate
                        } else classified = true;
                    }
                    if (classified) res.assignment[r] = ti;
                } else if (deplete) {                           // classify.hpp:280-281
                    if (dflag[j]) { res.failed++; res.assignment[r] = -3; continue; }      // ShortReadException -> failed (classify.hpp:306-316)
                    classified = best_of(dc, j, di) > 0;
                    if (classified) res.assignment[r] = -2;
                } else {                                        // classify.hpp:283-292
                    if (tflag[j]) { res.failed++; res.assignment[r] = -3; continue; }
                    classified = best_of(tc, j, ti) > 0;
                    if (classified) res.assignment[r] = ti;
                }
                if (classified) {
                    res.found++;
                    if (res.assignment[r] >= 0) TargetFilters[res.assignment[r]].classified += 1;
                } else still.push_back(r);
            }
            active.swap(still);
        }

        // outputs in file order: <output_dir>/<target name>.fasta and unclassified.fasta (classify.hpp:196-218,299-300)
        std::vector<std::ofstream> targetFastas;
        for (IBFMeta &f : TargetFilters) targetFastas.emplace_back(config.output_dir / (f.name + ".fasta"));
        std::ofstream unclassified(config.output_dir / "unclassified.fasta");
        std::vector<std::string> tbuf(TargetFilters.size());
        std::string ubuf;
        for (size_t r = 0; r < n; ++r) {
            const int a = res.assignment[r];
            if (a >= 0) {                                       // same bytes as `<< ">" << id << endl << seq << endl`, one write
                std::string &b = tbuf[a];
                b.append(1, '>').append(reads[r].id).append(1, '\n').append(reads[r].seq).append(1, '\n');
                if (b.size() > (8u << 20)) { targetFastas[a].write(b.data(), (std::streamsize)b.size()); b.clear(); }
            } else if (a == -1) {                               // seqan::writeRecord(out, id, (Dna5String) seq): 70 columns
                ubuf.append(1, '>').append(reads[r].id).append(1, '\n');
                const std::string &q = reads[r].seq;
                for (size_t p = 0; p < q.size(); p += 70) {
                    const size_t e = std::min(q.size(), p + 70), o = ubuf.size();
                    ubuf.append(q, p, e - p);
                    for (size_t i = o; i < ubuf.size(); ++i) {  // the Dna5String cast: A C G T (U = T), everything else N
                        const char c = ubuf[i] & 0xDF;
                        ubuf[i] = (c == 'A' || c == 'C' || c == 'G' || c == 'T') ? c : (c == 'U' ? 'T' : 'N');
                    }
                    ubuf.append(1, '\n');
                }
                if (ubuf.size() > (8u << 20)) { unclassified.write(ubuf.data(), (std::streamsize)ubuf.size()); ubuf.clear(); }
            }
        }
        for (size_t t = 0; t < tbuf.size(); ++t) { targetFastas[t].write(tbuf[t].data(), (std::streamsize)tbuf[t].size()); targetFastas[t].flush(); }
        unclassified.write(ubuf.data(), (std::streamsize)ubuf.size());
        unclassified.flush();
        // the reference's per-read timer covers the chunk loop, the decision and the output record of a read; here the same work
        // is done for all reads of the file at once
        res.avgClassifyduration = res.readCounter ? secs(t_classify, clk::now()) / (double)res.readCounter : 0.0;
        std::cout << "------------------------------- Final Results -------------------------------" << std::endl;
        std::cout << "Number of classified reads                         :   " << res.found << std::endl;
        std::cout << "Number of of too short reads (len < " << cl << ")           :   " << res.too_short << std::endl;
        std::cout << "Number of all reads                                :   " << res.readCounter << std::endl;
        for (IBFMeta &f : TargetFilters)
            std::cout << f.name << "\t : " << f.classified << "\t\t" << ((float)f.classified) / ((float)res.readCounter) << std::endl;
        std::cout << "Average Processing Time Read Classification        :   " << res.avgClassifyduration << std::endl;
        std::cout << "-----------------------------------------------------------------------------------" << std::endl;
    }
    return res;
}

// run_program (src/main/main.cpp:274-406) for the two GPU-backed usages
inline int run_program(const ConfigReader &config, ClassificationResults *results = nullptr)
{
    if (config.usage == "build") return run_build(config);
    if (config.usage == "classify") {
        std::filesystem::create_directories(config.output_dir);
        ClassificationResults r = classify_reads(config, getIBF(config, true, false), getIBF(config, false, true));
        if (results) *results = r;
        return 0;
    }
    std::cerr << "usage \"" << config.usage << "\" needs the MinKNOW client and basecallers of the host application" << std::endl;
    return 2;
}

}  // namespace rbdrv
