// Drives rblive::LiveClassifier with a recorded chunk stream and prints its decisions.
//   test_live <stream.tsv> <error_rate> <n_dep> <dep.ibf>... <n_tgt> <tgt.ibf>...
// stream.tsv: one line per chunk "batch<TAB>read_id<TAB>sequence"; chunks of one batch are classified together.
#include "rb_live.hpp"

#include <cstdlib>
#include <iostream>
#include <sstream>

static std::vector<interleave::IBFMeta> load(int &i, char **argv)
{
    std::vector<interleave::IBFMeta> v;
    int n = std::atoi(argv[i++]);
    for (int j = 0; j < n; ++j) {
        interleave::IBF f;
        interleave::IBFConfig c;
        c.input_filter_file = argv[i++];
        f.load_filter(c);
        interleave::IBFMeta m;
        m.filter = f.getFilter();
        m.name = c.input_filter_file;
        v.push_back(m);
    }
    return v;
}

int main(int argc, char **argv)
{
    if (argc < 5) return 2;
    int i = 3;
    interleave::ClassifyConfig conf;
    conf.error_rate = std::atof(argv[2]);
    conf.significance = 0.95;
    std::vector<interleave::IBFMeta> dep = load(i, argv), tgt = load(i, argv);
    rblive::LiveClassifier lc(dep, tgt, conf);
    std::ifstream in(argv[1]);
    std::string line;
    long cur = -1;
    std::vector<rblive::LiveRead> batch, deferred;
    auto flush = [&]() {
        while (!batch.empty()) {
            deferred.clear();
            for (rblive::LiveDecision &d : lc.classify_batch(std::move(batch), &deferred))
                std::cout << d.read.id << "\t" << (int)d.action << "\t" << (int)d.gave_up << "\t" << d.seen.size() << "\n";
            batch = std::move(deferred);
        }
    };
    while (std::getline(in, line)) {
        std::istringstream ss(line);
        std::string b, id, seq;
        std::getline(ss, b, '\t'); std::getline(ss, id, '\t'); std::getline(ss, seq, '\t');
        long bi = std::atol(b.c_str());
        if (bi != cur) { flush(); cur = bi; }
        rblive::LiveRead r;
        r.id = id; r.sequence = seq;
        batch.push_back(std::move(r));
    }
    flush();
    std::cout << "PENDING\t" << lc.pending() << "\n";
    return 0;
}
