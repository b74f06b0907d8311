// rb_interleave.hpp -- C++17 host-side mirror of ReadBouncer's src/IBF interface on top of the
// C ABI in rb_ibf.h.  Header-only; link with -lrb_ibf.
//
// It keeps the reference's names, argument meaning and error behaviour so that the drivers in
// src/main (classify.hpp, adaptive_sampling.hpp, ibfbuild.hpp) compile against it unchanged in
// spirit: interleave::IBF (create_filter / load_filter / getFilter), interleave::IBFMeta,
// interleave::Read::classify x3, interleave::calculateCI, the exception tree of
// src/IBF/IBFExceptions.hpp, ClassifyConfig / IBFConfig of src/IBF/IBFConfig.hpp, and
// check_unblock of src/main/adaptive_sampling.hpp:35-113.  What changes underneath: the filter
// lives in B200 HBM (TIbf is a shared handle, not a multi-GB value type) and every classify call
// is one GPU batch.  The batch entry points (Read::classify_batch, check_unblock_batch) are the
// ones a throughput-minded caller should use; the per-read overloads exist for drop-in parity.
#pragma once

#include "rb_ibf.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <exception>
#include <fstream>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

namespace interleave {

// ---- exceptions (src/IBF/IBFExceptions.hpp:16-372) ---------------------------------------------------
class IBFException : public std::exception {
public:
    explicit IBFException(std::string msg = "") : error_message(std::move(msg)) {}
    const char *what() const noexcept override { return error_message.c_str(); }
private:
    std::string error_message;
};
#define RB_DEFINE_EXC(name, base) \
    class name : public base { public: explicit name(std::string msg = "") : base(std::move(msg)) {} }
RB_DEFINE_EXC(IBFBuildException, IBFException);
RB_DEFINE_EXC(IBFClassifyException, IBFException);
RB_DEFINE_EXC(ShortReadException, IBFClassifyException);
RB_DEFINE_EXC(CountKmerException, IBFClassifyException);
RB_DEFINE_EXC(InvalidConfigException, IBFBuildException);
RB_DEFINE_EXC(NullFilterException, IBFBuildException);
RB_DEFINE_EXC(InsertSequenceException, IBFBuildException);
RB_DEFINE_EXC(StoreFilterException, IBFBuildException);
RB_DEFINE_EXC(FileParserException, IBFBuildException);
RB_DEFINE_EXC(MissingReferenceFilesException, FileParserException);
RB_DEFINE_EXC(MissingIBFFileException, FileParserException);
RB_DEFINE_EXC(ParseIBFFileException, FileParserException);
#undef RB_DEFINE_EXC

// status code -> the reference's exception
[[noreturn]] inline void throw_status(int status, const std::string &context = "")
{
    const std::string msg = (context.empty() ? "" : context + ": ") + rb_last_error();
    switch (status) {
    case RB_ERR_NULL_FILTER: throw NullFilterException(msg);
    case RB_ERR_SHORT_READ: throw ShortReadException(msg);
    case RB_ERR_COUNT_KMER: throw CountKmerException(msg);
    case RB_ERR_PARSE_IBF_FILE: throw ParseIBFFileException(msg);
    case RB_ERR_MISSING_IBF_FILE: throw MissingIBFFileException(msg);
    case RB_ERR_STORE_FILTER: throw StoreFilterException(msg);
    case RB_ERR_INSERT_SEQUENCE: throw InsertSequenceException(msg);
    case RB_ERR_INVALID_CONFIG: throw InvalidConfigException(msg);
    default: throw IBFException(std::string(rb_status_string(status)) + ": " + msg);
    }
}

// ---- configuration (src/IBF/IBFConfig.hpp:23-145) ---------------------------------------------------------
class ClassifyConfig {
public:
    double significance = 0.95;
    double error_rate = 0.1;
    uint16_t max_error = 0;
    uint16_t strata_filter = 0;
};

class IBFConfig {
public:
    static constexpr uint32_t MBinBits = 8388608;
    std::vector<std::string> reference_files;
    std::string directory_reference_files = "";
    std::string extension = "";
    std::string output_filter_file = "";
    std::string input_filter_file = "";
    std::string update_filter_file = "";
    bool update_complete = false;
    uint64_t filter_size = 0;
    uint64_t filter_size_bits = 0;
    uint64_t fragment_length = 0;
    uint16_t overlap_length = 1500;
    uint16_t kmer_size = 13;
    uint16_t hash_functions = 3;
    uint16_t threads = 2;
    uint32_t n_refs = 400;
    uint32_t n_batches = 500000;
    double max_fp = 0.01;
    bool verbose = false;
    bool quiet = false;
    uint16_t threads_build = 1;
    int device = 0;   // new: CUDA device that holds the filter

    bool validate()
    {
        threads_build = threads <= 2 ? 1 : threads - 1;
        if (n_batches < 1) n_batches = 1;
        if (n_refs < 1) n_refs = 1;
        if (!update_filter_file.empty()) {
            kmer_size = 0; hash_functions = 0; filter_size = 0; filter_size_bits = 0;
        } else if (filter_size_bits != 0) {
            filter_size = filter_size_bits / MBinBits;
        } else if (filter_size != 0) {
            filter_size_bits = filter_size * MBinBits;
        }
        return true;
    }
};

struct FilterStats {
    uint64_t sumSeqLen = 0;
    uint64_t totalSeqsBinId = 0;
    uint32_t totalBinsBinId = 0;
    uint64_t totalSeqsFile = 0;
    uint32_t totalBinsFile = 0;
    uint64_t invalidSeqs = 0;
    uint32_t newBins = 0;
};

// ---- TIbf: shared handle on a device-resident filter -------------------------------------------------------
// Stands in for seqan::BinningDirectory<InterleavedBloomFilter, ...> (src/IBF/IBF.hpp:92-94); keeps the
// two public fields the reference reads (filter.noOfBins, filter.kmerSize, src/IBF/IBFClassify.cpp:27,102).
class TIbf {
public:
    TIbf() = default;
    explicit TIbf(rb_ibf *h) : handle_(h, rb_ibf_free)
    {
        rb_ibf_info_t info{};
        if (rb_ibf_info(h, &info) == RB_OK) {
            noOfBins = info.n_bins; kmerSize = (uint16_t)info.kmer_size; noOfHashFunc = (uint16_t)info.n_hash;
            noOfBits = info.n_bits;
        }
    }
    // re-read the metadata of the same handle (after resizeBins)
    void refresh()
    {
        rb_ibf_info_t info{};
        if (handle_ && rb_ibf_info(handle_.get(), &info) == RB_OK) {
            noOfBins = info.n_bins; kmerSize = (uint16_t)info.kmer_size; noOfHashFunc = (uint16_t)info.n_hash;
            noOfBits = info.n_bits;
        }
    }
    rb_ibf *get() const { return handle_.get(); }
    explicit operator bool() const { return (bool)handle_; }
    uint64_t noOfBins = 0;
    uint64_t noOfBits = 0;
    uint16_t kmerSize = 0;
    uint16_t noOfHashFunc = 0;
private:
    std::shared_ptr<rb_ibf> handle_;
};

struct IBFMeta {
    TIbf filter;
    std::string name;
    uint64_t classified = 0;
};

// One joint k-mer table plan for every filter a caller classifies against (rb_ibf_enable_kmer_tables): called once by
// classify_reads and LiveClassifier before the first read, so that no classify call stalls on a multi-GB table build
// and the first filter does not take the HBM the others need.
inline void enable_kmer_tables(const std::vector<IBFMeta> &a, const std::vector<IBFMeta> &b = {}, uint64_t total_bytes = 0)
{
    std::vector<rb_ibf *> hs;
    for (const std::vector<IBFMeta> *v : {&a, &b})
        for (const IBFMeta &f : *v)
            if (f.filter) hs.push_back(f.filter.get());
    if (hs.empty()) return;
    int st = rb_ibf_enable_kmer_tables(hs.data(), (uint32_t)hs.size(), total_bytes, nullptr);
    if (st != RB_OK) throw_status(st, "enable_kmer_tables");
}

typedef std::pair<uint16_t, uint16_t> TInterval;

// calculateCI (src/IBF/IBF.hpp:320-338)
inline TInterval calculateCI(const double r, const uint8_t kmer_size, const uint32_t readlen, const double confidence)
{
    uint16_t lo = 0, hi = 0;
    int st = rb_calculate_ci(r, kmer_size, readlen, confidence, &lo, &hi);
    if (st != RB_OK) throw_status(st, "calculateCI");
    return TInterval{lo, hi};
}

// ---- sequence files (host I/O; minimal FASTA/FASTQ reader standing in for seqan::SeqFileIn) ---------------
struct SeqRecord {
    std::string id;
    std::string seq;
};

inline std::vector<SeqRecord> read_sequence_file(const std::string &path)
{
    std::ifstream in(path, std::ios::binary);
    if (!in) throw FileParserException("Unable to open the file: " + path);
    std::vector<SeqRecord> recs;
    std::string line;
    auto chomp = [](std::string &s) { while (!s.empty() && (s.back() == '\r' || s.back() == '\n')) s.pop_back(); };
    bool have = (bool)std::getline(in, line);
    while (have) {
        chomp(line);
        if (line.empty()) { have = (bool)std::getline(in, line); continue; }
        if (line[0] == '>') {
            SeqRecord r; r.id = line.substr(1);
            while ((have = (bool)std::getline(in, line))) {
                chomp(line);
                if (!line.empty() && line[0] == '>') break;
                r.seq += line;
            }
            recs.push_back(std::move(r));
        } else if (line[0] == '@') {
            SeqRecord r; r.id = line.substr(1);
            if (!std::getline(in, r.seq)) throw FileParserException("truncated FASTQ record in " + path);
            chomp(r.seq);
            std::string plus, qual;
            if (!std::getline(in, plus) || !std::getline(in, qual)) throw FileParserException("truncated FASTQ record in " + path);
            recs.push_back(std::move(r));
            have = (bool)std::getline(in, line);
        } else {
            throw FileParserException("ERROR: Problems parsing the file: " + path);
        }
    }
    return recs;
}

// ---- interleave::IBF (src/IBF/IBF.hpp:103-159, src/IBF/IBFBuild.cpp) -----------------------------------------
class IBF {
public:
    // create_filter, src/IBF/IBFBuild.cpp:421-521: parse references, cut N runs, reserve len/F+1 bins per
    // sequence, size the filter, insert every fragment with consecutive bin ids, store the file.
    FilterStats create_filter(IBFConfig &config)
    {
        if (!config.validate()) throw InvalidConfigException("Config not valid!");
        FilterStats stats;
        FragmentPlan plan = parse_ref_seqs(config, stats, 0);
        if (stats.totalBinsBinId == 0) throw NullFilterException("Could not instantiate IBF Filter");
        config.filter_size_bits = rb_ibf_size_bits(config.fragment_length, config.kmer_size, config.hash_functions,
                                                   config.max_fp, stats.totalBinsBinId);
        int st = RB_OK;
        rb_ibf *h = rb_ibf_create(stats.totalBinsBinId, config.hash_functions, config.kmer_size, config.filter_size_bits,
                                  config.device, &st);
        if (!h) {
            if (st == RB_ERR_INVALID_CONFIG) throw NullFilterException("Could not instantiate IBF Filter");
            throw_status(st, "create_filter");
        }
        filter = TIbf(h);
        stats.totalBinsFile = (uint32_t)filter.noOfBins;
        add_sequences_to_filter(plan);
        if (!config.output_filter_file.empty()) {
            st = rb_ibf_store(h, config.output_filter_file.c_str());
            if (st != RB_OK) throw_status(st, "Could not store IBF to " + config.output_filter_file);
        }
        return stats;
    }

    // update_filter, src/IBF/IBFBuild.cpp:223-321: load update_filter_file, append the bins of the new
    // references (resizeBins), insert their fragments from bin id = old bin count, store back.
    FilterStats update_filter(IBFConfig &config)
    {
        if (!config.validate()) throw InvalidConfigException("Config not valid!");
        if (config.update_filter_file.empty())
            throw MissingIBFFileException("Error: Either update_filter_file or input_filter_file have to be specified.");
        FilterStats stats = load_filter(config);                 // sets config.kmer_size from the file
        FragmentPlan plan = parse_ref_seqs(config, stats, stats.totalBinsFile);
        const uint32_t number_new_bins = stats.totalBinsBinId + stats.totalBinsFile;
        if (number_new_bins > stats.totalBinsFile) {
            int st = rb_ibf_resize_bins(filter.get(), number_new_bins, nullptr);
            if (st != RB_OK) throw_status(st, "resizeBins");
            filter.refresh();                                    // noOfBins / noOfBits changed
            stats.newBins = stats.totalBinsBinId;
            stats.totalBinsBinId = number_new_bins;
        }
        add_sequences_to_filter(plan);
        int st = rb_ibf_store(filter.get(), config.update_filter_file.c_str());
        if (st != RB_OK) throw_status(st, "Could not store IBF to " + config.update_filter_file);
        return stats;
    }

    // load_filter, src/IBF/IBFBuild.cpp:329-396
    FilterStats load_filter(IBFConfig &config)
    {
        const std::string &path = !config.update_filter_file.empty() ? config.update_filter_file : config.input_filter_file;
        if (path.empty())
            throw MissingIBFFileException("Error: Either update_filter_file or input_filter_file have to be specified.");
        int st = RB_OK;
        rb_ibf *h = rb_ibf_load(path.c_str(), config.device, &st);
        if (!h) {
            // the reference turns every seqan::retrieve failure into ParseIBFFileException (IBFBuild.cpp:345-369)
            if (st == RB_ERR_MISSING_IBF_FILE || st == RB_ERR_PARSE_IBF_FILE)
                throw ParseIBFFileException("Error parsing IBF input file " + path + ": " + rb_last_error());
            throw_status(st, "load_filter");
        }
        filter = TIbf(h);
        FilterStats stats;
        stats.totalBinsFile = (uint32_t)filter.noOfBins;
        config.kmer_size = filter.kmerSize;
        return stats;
    }

    TIbf getFilter() { return filter; }

private:
    struct FragmentPlan {
        std::string bases;
        std::vector<uint64_t> begin, end, bin;
    };

    // parse_ref_seqs + the fragment loop of add_sequences_to_filter (src/IBF/IBFBuild.cpp:16-104,165-202)
    FragmentPlan parse_ref_seqs(IBFConfig &config, FilterStats &stats, uint64_t first_bin)
    {
        if (config.reference_files.empty()) throw MissingReferenceFilesException("There were no reference files specified!");
        if (config.fragment_length < config.kmer_size) throw InvalidConfigException("fragment_length must be >= kmer_size");
        FragmentPlan plan;
        uint64_t binid = first_bin;
        for (const std::string &file : config.reference_files) {
            for (SeqRecord &rec : read_sequence_file(file)) {
                stats.totalSeqsFile += 1;
                if (rec.seq.size() < config.kmer_size) { stats.invalidSeqs += 1; continue; }   // IBFBuild.cpp:70-74
                std::string cut(rec.seq.size(), '\0');
                cut.resize(rb_cut_out_nnns(rec.seq.data(), rec.seq.size(), cut.data()));
                stats.totalBinsBinId += (uint32_t)(cut.size() / config.fragment_length + 1);          // IBFBuild.cpp:90
                stats.sumSeqLen += cut.size();
                const uint64_t n = rb_fragment_schedule(cut.size(), config.fragment_length, config.kmer_size, nullptr, nullptr, 0);
                std::vector<uint64_t> b(n), e(n);
                rb_fragment_schedule(cut.size(), config.fragment_length, config.kmer_size, b.data(), e.data(), n);
                for (uint64_t i = 0; i < n; ++i) {
                    plan.begin.push_back(plan.bases.size() + b[i]);
                    plan.end.push_back(plan.bases.size() + e[i]);
                    plan.bin.push_back(binid++);
                }
                plan.bases += cut;
            }
        }
        return plan;
    }

    void add_sequences_to_filter(const FragmentPlan &plan)
    {
        int st = rb_ibf_insert_batch(filter.get(), plan.bases.data(), plan.bases.size(), plan.begin.data(), plan.end.data(),
                                     plan.bin.data(), plan.begin.size(), nullptr);
        if (st != RB_OK) throw_status(st, "Error inserting the sequences to the IBF");
    }

    TIbf filter{};
};

// ---- threshold tables, cached per (error_rate, significance, k) ---------------------------------------------------
// raw: no range check on the rate -- the retry of check_unblock / classify_deplete_target at error_rate - 0.02 is never
// validated by the reference (adaptive_sampling.hpp:55-59), so a configured rate <= 0.02 must not make a batch throw
inline const std::vector<uint16_t> &threshold_lut(double error_rate, double significance, uint32_t k, bool raw = false)
{
    static std::mutex mu;
    static std::map<std::tuple<double, double, uint32_t, bool>, std::vector<uint16_t>> cache;
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_tuple(error_rate, significance, k, raw);
    auto it = cache.find(key);
    if (it == cache.end()) {
        std::vector<uint16_t> lut(65536);
        int st = raw ? rb_threshold_lut_raw(error_rate, significance, k, lut.data())
                     : rb_threshold_lut(error_rate, significance, k, lut.data());
        if (st != RB_OK) throw_status(st, "threshold");
        it = cache.emplace(key, std::move(lut)).first;
    }
    return it->second;
}

// Per-read summaries of one filter for a whole batch (one GPU pass; two thresholds when retry != 0).
struct BatchCounts {
    std::vector<uint16_t> max_count;     // [n_lut][n]
    std::vector<uint8_t> hit;            // [n_lut][n]
    std::vector<uint32_t> argmax_bin;    // [n_lut][n]
    std::vector<uint8_t> read_flag;      // [n]
    uint32_t n_lut = 1;
    uint64_t n = 0;
};

inline BatchCounts count_matches_batch(const TIbf &filter, const char *bases, const uint64_t *read_off, uint64_t n_reads,
                                       const ClassifyConfig &config, bool with_retry_threshold = false)
{
    if (!filter) throw NullFilterException("No IBF provided to classify the read!");
    BatchCounts out;
    out.n = n_reads;
    out.n_lut = with_retry_threshold ? 2 : 1;
    std::vector<uint16_t> luts(threshold_lut(config.error_rate, config.significance, filter.kmerSize));
    if (with_retry_threshold) {
        const std::vector<uint16_t> &l2 = threshold_lut(config.error_rate - 0.02, config.significance, filter.kmerSize, true);
        luts.insert(luts.end(), l2.begin(), l2.end());
    }
    out.max_count.resize(out.n_lut * n_reads);
    out.hit.resize(out.n_lut * n_reads);
    out.argmax_bin.resize(out.n_lut * n_reads);
    out.read_flag.resize(n_reads);
    int st = rb_ibf_count_batch(filter.get(), bases, read_off, n_reads, luts.data(), out.n_lut, nullptr, nullptr,
                                out.max_count.data(), out.hit.data(), out.argmax_bin.data(), out.read_flag.data(), nullptr);
    if (st != RB_OK) throw_status(st, "Error counting kmers in IBF bins");
    return out;
}

// The same for several filters, one after the other.  (One host thread per filter -- the reference spawns one std::async per
// filter per read, src/IBF/IBFClassify.cpp:259 -- was measured and is slower here: 623 vs 445 us for a micro-batch of 64 chunks
// against 4 filters, 4.09 vs 2.76 ms for 4 096; thread start-up and the contention of four packers cost more than the calls'
// synchronisation they would hide, profiles/r2_k_live_microbatch_latency.jsonl.)
inline std::vector<BatchCounts> count_matches_batch_all(const std::vector<const TIbf *> &filters, const char *bases,
                                                        const uint64_t *read_off, uint64_t n_reads, const ClassifyConfig &config,
                                                        bool with_retry_threshold = false)
{
    std::vector<BatchCounts> out(filters.size());
    for (size_t i = 0; i < filters.size(); ++i)
        out[i] = count_matches_batch(*filters[i], bases, read_off, n_reads, config, with_retry_threshold);
    return out;
}

// ---- interleave::Read (src/IBF/IBF.hpp:169-226, src/IBF/IBFClassify.cpp) --------------------------------------------
class Read {
public:
    std::string sequence{};
    std::string id{};

    Read() {}
    Read(const std::string &id_, const std::string &seq) : sequence(seq), id(id_) {}

    inline uint32_t getReadLength() const { return (uint32_t)sequence.size(); }

    // count_matches, src/IBF/IBFClassify.cpp:138-171
    uint64_t count_matches(const IBFMeta &filter, const ClassifyConfig &config) const
    {
        const uint64_t off[2] = {0, sequence.size()};
        BatchCounts c = count_matches_batch(filter.filter, sequence.data(), off, 1, config);
        return c.max_count[0];
    }

    // classify(std::vector<TIbf>&), src/IBF/IBFClassify.cpp:181-226 (+ find_matches :81-128)
    bool classify(std::vector<TIbf> &filters, ClassifyConfig &config)
    {
        if (filters.empty()) throw NullFilterException("No IBF provided to classify the read!");
        if (getReadLength() < filters[0].kmerSize) throw ShortReadException("Read " + id + " shorter than kmer size");
        const uint64_t off[2] = {0, sequence.size()};
        for (TIbf &f : filters) {
            BatchCounts c = count_matches_batch(f, sequence.data(), off, 1, config);
            if (c.hit[0]) return true;
        }
        return false;
    }

    // classify(std::vector<IBFMeta>&), src/IBF/IBFClassify.cpp:239-297: index of the filter with the strictly
    // greatest count_matches, -1 if all are 0
    int classify(std::vector<IBFMeta> &filters, ClassifyConfig &config)
    {
        if (filters.empty()) throw NullFilterException("No IBF provided to classify the read!");
        if (sequence.size() < filters[0].filter.kmerSize) throw ShortReadException("Read " + id + " shorter than kmer size");
        uint64_t best = 0;
        int best_index = -1;
        for (size_t i = 0; i < filters.size(); ++i) {
            uint64_t c = count_matches(filters[i], config);
            if (c > best) { best = c; best_index = (int)i; }
        }
        return best_index;
    }

    // classify(filt1, filt2), src/IBF/IBFClassify.cpp:299-365: filters with k > read length are skipped
    std::pair<int, int> classify(std::vector<IBFMeta> &filt1, std::vector<IBFMeta> &filt2, ClassifyConfig &config)
    {
        if (filt1.empty() || filt2.empty()) throw NullFilterException("No IBF provided to classify the read!");
        uint64_t a = 0, b = 0;
        for (IBFMeta &f : filt1)
            if (sequence.size() >= f.filter.kmerSize) a = std::max(a, count_matches(f, config));
        for (IBFMeta &f : filt2)
            if (sequence.size() >= f.filter.kmerSize) b = std::max(b, count_matches(f, config));
        return std::make_pair((int)a, (int)b);
    }
};

typedef std::vector<Read> TReads;

}  // namespace interleave

// ---- check_unblock (src/main/adaptive_sampling.hpp:35-113): 0 keep / 1 unblock / 2 stop_further_data ------------------
inline uint8_t check_unblock(interleave::Read &read, interleave::ClassifyConfig &conf,
                             std::vector<interleave::IBFMeta> &DepletionFilters,
                             std::vector<interleave::IBFMeta> &TargetFilters)
{
    const bool withTarget = !TargetFilters.empty(), withDepletion = !DepletionFilters.empty();
    if (withDepletion && withTarget) {
        std::pair<uint64_t, uint64_t> p = read.classify(DepletionFilters, TargetFilters, conf);
        if (p.first > 0) {
            if (p.second > 0) {
                interleave::ClassifyConfig strict = conf;       // the reference mutates conf in place (race, quirk Q7)
                strict.error_rate -= 0.02;
                p = read.classify(DepletionFilters, TargetFilters, strict);
                return (p.first > 0 && p.second == 0) ? 1 : 0;
            }
            return 1;
        }
        return p.second > 0 ? 2 : 0;
    }
    if (withDepletion) return read.classify(DepletionFilters, conf) > -1 ? 1 : 0;
    return read.classify(TargetFilters, conf) < 0 ? 1 : 2;
}

// Batch form of check_unblock: one GPU pass per filter evaluates both thresholds (error_rate and
// error_rate - 0.02), then the decision table above is applied per read on the host.
// Reads shorter than k follow the reference: the pair overload skips the filter (count 0); the single-list
// overloads throw ShortReadException per read, reported here as decision 255.  Reads longer than 65 535 bases
// (read_flag 2) are decision 255 in every mode.
inline std::vector<uint8_t> check_unblock_batch(const char *bases, const uint64_t *read_off, uint64_t n_reads,
                                                const interleave::ClassifyConfig &conf,
                                                std::vector<interleave::IBFMeta> &DepletionFilters,
                                                std::vector<interleave::IBFMeta> &TargetFilters)
{
    using namespace interleave;
    const bool withTarget = !TargetFilters.empty(), withDepletion = !DepletionFilters.empty();
    if (!withTarget && !withDepletion) throw NullFilterException("No IBF provided to classify the read!");
    const bool both = withTarget && withDepletion;
    // all filters of both sets, then the best count per set
    std::vector<const TIbf *> all;
    for (IBFMeta &f : DepletionFilters) all.push_back(&f.filter);
    for (IBFMeta &f : TargetFilters) all.push_back(&f.filter);
    const std::vector<BatchCounts> counts = count_matches_batch_all(all, bases, read_off, n_reads, conf, both);
    auto best_of = [&](size_t first, size_t count, std::vector<uint16_t> &best, std::vector<uint16_t> &best_strict,
                       std::vector<uint8_t> &flag) {
        best.assign(n_reads, 0); best_strict.assign(n_reads, 0); flag.assign(n_reads, 0);
        for (size_t fi = first; fi < first + count; ++fi) {
            const BatchCounts &c = counts[fi];
            for (uint64_t i = 0; i < n_reads; ++i) {
                best[i] = std::max(best[i], c.max_count[i]);
                if (both) best_strict[i] = std::max(best_strict[i], c.max_count[n_reads + i]);
                flag[i] = std::max(flag[i], c.read_flag[i]);
            }
        }
    };
    std::vector<uint16_t> dep, dep_s, tgt, tgt_s;
    std::vector<uint8_t> fd, ft;
    std::vector<uint8_t> out(n_reads, 0);
    if (withDepletion) best_of(0, DepletionFilters.size(), dep, dep_s, fd);
    if (withTarget) best_of(DepletionFilters.size(), TargetFilters.size(), tgt, tgt_s, ft);
    for (uint64_t i = 0; i < n_reads; ++i) {
        if (both) {
            // reads shorter than k are skipped per filter by the pair overload (count 0); a read the engine cannot
            // classify at all (longer than 65 535 bases: the reference's uint16 readlen) is reported, not kept silently
            if (fd[i] >= 2 || ft[i] >= 2) { out[i] = 255; continue; }
            if (dep[i] > 0) {
                if (tgt[i] > 0) out[i] = (dep_s[i] > 0 && tgt_s[i] == 0) ? 1 : 0;
                else out[i] = 1;
            } else out[i] = tgt[i] > 0 ? 2 : 0;
        } else if (withDepletion) {
            out[i] = fd[i] ? 255 : (dep[i] > 0 ? 1 : 0);
        } else {
            out[i] = ft[i] ? 255 : (tgt[i] > 0 ? 2 : 1);
        }
    }
    return out;
}
