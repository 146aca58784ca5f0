// TEST INFRASTRUCTURE (oracle/): runs the UNMODIFIED reference input path — CompleteSample, parseBreakpoints,
// the three site filters and getSubSample (/root/reference/src/Sample.cpp, driven as main.cpp:107-283 does) —
// on a set of FASTA files and prints, for every eligible locus pair, the Sample it would hand to runSampler.
// Used to pin the C++ host driver's own input path (coalescenator_b200/host/fasta_input.cpp).
//
// usage: coalref_input <fasta-list> <time-list> <viral-list> <loci-size | coordinate-list> <min-dist> <max-dist>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "Utils.hpp"
#include "Sample.hpp"

static std::string bits(const std::vector<bool>& v) { std::string s; for (size_t i = 0; i < v.size(); i++) s.push_back(v[i] ? '1' : '0'); return s; }

int main(int argc, char** argv) {
    if (argc < 7) { fprintf(stderr, "usage: coalref_input fasta times virals loci min max\n"); return 2; }
    std::streambuf* keep = std::cout.rdbuf(std::cerr.rdbuf());      // the reference chats on stdout
    CompleteSample cs(argv[1], argv[2], argv[3]);
    std::string loci_arg = argv[4];
    if (loci_arg.find_first_of("[,") != std::string::npos) cs.parseBreakpoints(loci_arg); else cs.parseBreakpoints((uint)atoi(argv[4]));
    cs.filterAmbiguousSites(); cs.filterIndels(); cs.filterNonBiAllelic();
    std::cout.rdbuf(keep);
    const uint min_dist = atoi(argv[5]), max_dist = atoi(argv[6]);
    std::vector<std::pair<uint, uint> > loci = cs.getLociCoordinates(), orig = cs.getOriginalLociCoordinates();
    printf("loci %zu\n", loci.size());
    for (size_t i = 0; i < loci.size(); i++) printf("locus %u %u %u %u\n", loci[i].first, loci[i].second, orig[i].first, orig[i].second);
    for (size_t i = 0; i < loci.size(); i++) for (size_t j = i + 1; j < loci.size(); j++) {
        uint dist = orig[j].first - orig[i].second - 1;
        if (dist < min_dist || dist > max_dist) continue;
        std::pair<Sample, std::vector<uint> > sub = cs.getSubSample(loci[i], loci[j]);
        Sample& s = sub.first;
        std::vector<AncestralTypeMaps> maps = s.getTypeMaps();
        std::vector<double> times = s.getTimes(), viral = s.getViralLoads();
        printf("pair %u %u %u %u dist %u LA %u LB %u T %zu counts %u %u %u\n", orig[i].first, orig[i].second, orig[j].first, orig[j].second, dist,
               s.getSeqLens().first, s.getSeqLens().second, times.size(), sub.second[0], sub.second[1], sub.second[2]);
        for (size_t t = 0; t < times.size(); t++) {
            std::vector<std::string> lines;
            for (auto& kv : maps[t].A) lines.push_back("A " + bits(kv.first) + " " + std::to_string(kv.second));
            for (auto& kv : maps[t].B) lines.push_back("B " + bits(kv.first) + " " + std::to_string(kv.second));
            for (auto& kv : maps[t].C) lines.push_back("C " + bits(kv.first) + " " + std::to_string(kv.second));
            std::sort(lines.begin(), lines.end());
            printf("time %.17g viral %.17g types %zu\n", times[t], viral[t], lines.size());
            for (auto& l : lines) printf("%s\n", l.c_str());
        }
    }
    return 0;
}
