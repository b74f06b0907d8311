// rb_live.hpp -- deadline-batched replacement for classify_live_reads.
//
// The reference's live loop (src/main/adaptive_sampling.hpp:214-356) pops ONE basecalled chunk at a
// time from classification_queue, calls check_unblock, and keeps unclassified chunks in `once_seen`
// so that the next chunk of the same read is classified again on the concatenation, giving up (read
// is kept = stop_receiving) once more than 1500 bases have been seen.  Here the same state machine runs
// on micro-batches: everything that is waiting in the queue (up to max_batch chunks, or until
// deadline_ms after the first chunk) is classified in one GPU pass per filter, the chunks that stay
// unclassified but have history are concatenated and classified in a second pass, then the reference's
// rules are applied per read in arrival order.  MinKNOW / basecaller threads are untouched: the adaptor
// only needs `empty()`, `pop()` and `push()` on the two queues (util/SafeQueue.hpp:14-119).
#pragma once

#include "rb_interleave.hpp"

#include <atomic>
#include <chrono>
#include <thread>
#include <unordered_map>
#include <unordered_set>

namespace rblive {

struct LiveRead {            // the fields of interfaces::ONTRead the classifier needs (interfaces/ont_read.hpp:24-61)
    uint32_t channelNr = 0;
    uint32_t readNr = 0;
    std::string id;
    std::string sequence;    // basecalled chunk
};

enum Action : uint8_t { kKeepGoing = 0, kUnblock = 1, kStopReceiving = 2 };

struct LiveDecision {
    LiveRead read;           // the chunk that triggered the decision
    Action action;
    std::string seen;        // everything classified for this read (earlier chunks + this one) when concatenated
    bool gave_up = false;    // kept because more than 1500 unclassified bases were seen (adaptive_sampling.hpp:315)
};

class LiveClassifier {
public:
    LiveClassifier(std::vector<interleave::IBFMeta> depletion, std::vector<interleave::IBFMeta> target,
                   interleave::ClassifyConfig conf, uint32_t give_up_length = 1500)
        : dep_(std::move(depletion)), tgt_(std::move(target)), conf_(conf), give_up_(give_up_length)
    {
        if (dep_.empty() && tgt_.empty()) throw interleave::NullFilterException("No IBF provided to classify the read!");
        interleave::enable_kmer_tables(dep_, tgt_);       // up front: a micro-batch must never wait for a table build
    }

    size_t pending() const { return once_seen_.size(); }

    // One micro-batch.  Chunks whose read id already occurs earlier in the same batch are returned in
    // `deferred` (they must be classified after their predecessor; feed them into the next batch first).
    std::vector<LiveDecision> classify_batch(std::vector<LiveRead> chunks, std::vector<LiveRead> *deferred = nullptr)
    {
        std::vector<LiveRead> batch;
        std::unordered_set<std::string> ids;
        for (LiveRead &c : chunks) {
            if (ids.insert(c.id).second) batch.push_back(std::move(c));
            else if (deferred) deferred->push_back(std::move(c));
            else throw interleave::IBFClassifyException("duplicate read id in one micro-batch: " + c.id);
        }
        std::vector<LiveDecision> out;
        if (batch.empty()) return out;
        std::vector<uint8_t> d1 = decide([&](size_t i) -> const std::string & { return batch[i].sequence; }, batch.size());
        // second pass: unclassified chunks with history are classified again on the concatenation
        std::vector<size_t> again;
        std::vector<std::string> concat;
        for (size_t i = 0; i < batch.size(); ++i) {
            if (d1[i] != kKeepGoing) continue;
            auto it = once_seen_.find(batch[i].id);
            if (it != once_seen_.end()) { again.push_back(i); concat.push_back(it->second.first + batch[i].sequence); }
        }
        std::vector<uint8_t> d2;
        if (!again.empty()) d2 = decide([&](size_t j) -> const std::string & { return concat[j]; }, again.size());
        size_t a = 0;
        for (size_t i = 0; i < batch.size(); ++i) {
            LiveRead &r = batch[i];
            uint8_t d = d1[i];
            if (d == 255) continue;                       // classify threw for this read (too short): logged and dropped by the reference
            auto it = once_seen_.find(r.id);
            if (d == kUnblock) {                          // adaptive_sampling.hpp:241-262
                std::string seen = it != once_seen_.end() ? it->second.first + r.sequence : r.sequence;
                if (it != once_seen_.end()) once_seen_.erase(it);
                out.push_back({std::move(r), kUnblock, std::move(seen), false});
            } else if (d == kStopReceiving) {             // :263-273
                if (it != once_seen_.end()) once_seen_.erase(it);
                std::string seen = r.sequence;
                out.push_back({std::move(r), kStopReceiving, std::move(seen), false});
            } else if (it != once_seen_.end()) {          // :276-332
                std::string &cat = concat[a];
                const uint8_t dd = d2[a];
                ++a;
                if (dd == kUnblock || dd == kStopReceiving) {
                    once_seen_.erase(it);
                    out.push_back({std::move(r), (Action)dd, std::move(cat), false});
                } else if (dd == 255) {
                    // exception path of the reference: nothing stored, nothing sent
                } else if (cat.size() > give_up_) {
                    once_seen_.erase(it);
                    out.push_back({std::move(r), kStopReceiving, std::move(cat), true});
                } else {
                    it->second.first = std::move(cat);
                    it->second.second += 1;
                }
            } else {
                once_seen_.emplace(r.id, std::make_pair(std::move(r.sequence), (uint8_t)1));   // :333-337
            }
        }
        return out;
    }

    // Deadline-batched loop between two SafeQueue-like queues of Item; `get`/`make` adapt Item <-> LiveRead.
    // Runs until `finished` is set and the input queue is drained.
    template <class InQueue, class OutQueue, class Get, class Make>
    void run(InQueue &in, OutQueue &out, std::atomic<bool> &finished, Get get, Make make, size_t max_batch = 4096,
             double deadline_ms = 10.0)
    {
        using clock = std::chrono::steady_clock;
        std::vector<LiveRead> carry;
        while (true) {
            std::vector<LiveRead> batch = std::move(carry);
            carry.clear();
            clock::time_point first{};
            bool have = !batch.empty();
            if (have) first = clock::now();
            while (batch.size() < max_batch) {
                if (!in.empty()) {
                    batch.push_back(get(in.pop()));
                    if (!have) { have = true; first = clock::now(); }
                } else {
                    if (finished.load() || (have && std::chrono::duration<double, std::milli>(clock::now() - first).count() >= deadline_ms)) break;
                    std::this_thread::sleep_for(std::chrono::microseconds(50));
                }
            }
            if (!batch.empty())
                for (LiveDecision &d : classify_batch(std::move(batch), &carry)) out.push(make(std::move(d)));
            if (finished.load() && in.empty() && carry.empty()) break;
        }
    }

private:
    template <class SeqAt>
    std::vector<uint8_t> decide(SeqAt seq_at, size_t n)
    {
        std::string bases;
        std::vector<uint64_t> off{0};
        for (size_t i = 0; i < n; ++i) { bases += seq_at(i); off.push_back(bases.size()); }
        return check_unblock_batch(bases.data(), off.data(), n, conf_, dep_, tgt_);
    }

    std::vector<interleave::IBFMeta> dep_, tgt_;
    interleave::ClassifyConfig conf_;
    uint32_t give_up_;
    std::unordered_map<std::string, std::pair<std::string, uint8_t>> once_seen_;
};

}  // namespace rblive
