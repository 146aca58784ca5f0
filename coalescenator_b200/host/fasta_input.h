// Host-side input path of the C++ driver: serially sampled FASTA files -> per-locus-pair problems in the
// packed layout of include/coalgpu.h. Mirrors the semantics of the reference's CompleteSample
// (/root/reference/src/Sample.cpp:81-692) — allele coding, missing / indel / ambiguous flags, the three site
// filters with their coordinate remapping, locus tiling and sub-sample extraction — written from scratch on
// flat arrays (no per-pair rescans of vector<uint> keys: sequences are kept as byte rows + multiplicities).
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

namespace coalhost {

struct Locus { uint32_t first, second; };   // 1-based inclusive site coordinates

struct PairProblem {
    uint32_t LA = 0, LB = 0;
    std::vector<double> times, viral;
    std::vector<int32_t> type_offsets;      // [T+1]
    std::vector<uint8_t> type_gamma;        // 0 = A only, 1 = B only, 2 = both
    std::vector<uint64_t> type_words;       // locus A site s = bit s, locus B site s = bit LA+s
    std::vector<uint32_t> type_counts;
    uint32_t nA = 0, nB = 0, nC = 0;        // sequence counts by ancestry (main.cpp:279)
};

class Alignment {
  public:
    // Sample.cpp:81-296. Lists are comma separated; samples are sorted by time (stable).
    Alignment(const std::string& fasta_list, const std::string& time_list, const std::string& viral_list);

    void tile_loci(uint32_t loci_size);                 // Sample.cpp:386-411
    void parse_coordinates(const std::string& spec);    // Sample.cpp:299-383
    // Sample.cpp:426-453; each returns the number of sites left
    size_t filter_ambiguous();
    size_t filter_indels();
    size_t filter_non_biallelic();

    const std::vector<Locus>& loci() const { return loci_; }                    // coordinates after filtering
    const std::vector<Locus>& original_loci() const { return original_loci_; }  // same loci, input coordinates
    size_t n_sites() const { return n_sites_; }
    size_t n_reads() const { return n_reads_; }
    size_t n_times() const { return times_.size(); }
    const std::vector<double>& times() const { return times_; }
    const std::vector<double>& viral() const { return viral_; }
    std::vector<std::string> messages;     // what the reference prints while parsing / filtering

    // Sample.cpp:587-692. Throws std::runtime_error when no sequence covers either locus.
    PairProblem sub_sample(const Locus& a, const Locus& b) const;

  private:
    size_t n_sites_ = 0, n_reads_ = 0;
    std::vector<double> times_, viral_;
    // per sampling time: distinct coded sequences (row-major bytes, n_sites_ per row) + multiplicities
    std::vector<std::vector<uint8_t>> rows_;
    std::vector<std::vector<uint32_t>> mult_;
    std::vector<uint8_t> has_missing_, has_indel_, has_ambiguous_;
    std::vector<uint32_t> poly_count_;
    std::vector<Locus> loci_, original_loci_;
    size_t apply_filter(const std::vector<uint8_t>& keep);   // genericFilterFunction, Sample.cpp:456-584
};

std::vector<std::string> split(const std::string& s, const std::string& any_of);

}  // namespace coalhost
