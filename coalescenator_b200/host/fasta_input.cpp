#include "fasta_input.h"

#include <algorithm>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <unordered_map>

namespace coalhost {

std::vector<std::string> split(const std::string& s, const std::string& any_of) {
    std::vector<std::string> out;
    std::string cur;
    for (char c : s) {
        if (any_of.find(c) != std::string::npos) { out.push_back(cur); cur.clear(); } else cur.push_back(c);
    }
    out.push_back(cur);
    return out;
}

namespace {
// codes as in Sample.cpp:208-222: 0..3 alleles in first-seen order per site, 4 indel, 5 missing, 6 ambiguous
constexpr uint8_t kIndel = 4, kMissing = 5, kAmbiguous = 6;

struct RowHash {
    size_t operator()(const std::string& r) const { return std::hash<std::string>()(r); }
};
}  // namespace

Alignment::Alignment(const std::string& fasta_list, const std::string& time_list, const std::string& viral_list) {
    std::vector<std::string> files = split(fasta_list, ","), ts = split(time_list, ","), vs = split(viral_list, ",");
    if (ts.size() != files.size() || vs.size() != files.size()) throw std::runtime_error("fasta, time and viral lists differ in length");
    // sort by sampling time, stable (Sample.cpp:106-139)
    std::vector<size_t> order(files.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = i;
    std::vector<double> tval(files.size()), vval(files.size());
    for (size_t i = 0; i < files.size(); i++) { tval[i] = std::atof(ts[i].c_str()); vval[i] = std::atof(vs[i].c_str()); }
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return tval[a] < tval[b]; });

    std::vector<std::map<char, uint8_t>> alleles;   // per site, first-seen order over the whole input
    size_t unique_total = 0;
    for (size_t oi = 0; oi < order.size(); oi++) {
        const size_t f = order[oi];
        times_.push_back(tval[f]); viral_.push_back(vval[f]);
        std::ifstream in(files[f].c_str());
        if (!in.is_open()) throw std::runtime_error("cannot open " + files[f]);
        {
            std::ostringstream m;
            m << "Parsing: " << files[f] << " taken at time " << tval[f] << " and with a measured concentration of " << vval[f] << ".";
            messages.push_back(m.str());
        }
        std::unordered_map<std::string, size_t, RowHash> index;   // coded row -> position in rows_/mult_
        std::vector<uint8_t> rows; std::vector<uint32_t> mult;
        std::string line;
        while (std::getline(in, line)) {
            if (!line.empty() && line.back() == '\r') line.pop_back();
            if (line.empty() || line[0] == '>') continue;
            if (n_reads_ == 0) {
                n_sites_ = line.size();
                alleles.resize(n_sites_);
                has_missing_.assign(n_sites_, 0); has_indel_.assign(n_sites_, 0); has_ambiguous_.assign(n_sites_, 0);
            } else if (line.size() != n_sites_) {
                throw std::runtime_error("sequences differ in length (one sequence per line expected, Sample.cpp:179)");
            }
            n_reads_++;
            std::string coded(n_sites_, '\0');
            std::vector<uint8_t> indel_here(n_sites_, 0);
            bool left = true;
            for (size_t j = 0; j < n_sites_; j++) {
                const char ch = line[j];
                if (ch == '-' && left) { has_missing_[j] = 1; coded[j] = (char)kMissing; continue; }
                left = false;
                if (ch == '-') { indel_here[j] = 1; coded[j] = (char)kIndel; }
                else if (ch == 'N') { has_ambiguous_[j] = 1; coded[j] = (char)kAmbiguous; }
                else {
                    auto it = alleles[j].find(ch);
                    if (it == alleles[j].end()) it = alleles[j].insert({ch, (uint8_t)alleles[j].size()}).first;
                    coded[j] = (char)it->second;
                }
            }
            for (size_t j = n_sites_; j-- > 0;) {     // right-hand run of '-' is missing data, not an indel
                if (line[j] != '-') break;
                has_missing_[j] = 1; indel_here[j] = 0; coded[j] = (char)kMissing;
            }
            for (size_t j = 0; j < n_sites_; j++) if (indel_here[j]) has_indel_[j] = 1;
            auto it = index.find(coded);
            if (it == index.end()) {
                index.insert({coded, mult.size()});
                rows.insert(rows.end(), coded.begin(), coded.end());
                mult.push_back(1);
            } else {
                mult[it->second]++;
            }
        }
        unique_total += mult.size();
        rows_.push_back(std::move(rows)); mult_.push_back(std::move(mult));
    }
    if (n_reads_ == 0) throw std::runtime_error("no sequences read");
    poly_count_.resize(n_sites_);
    for (size_t j = 0; j < n_sites_; j++) poly_count_[j] = (uint32_t)alleles[j].size() + (has_indel_[j] ? 1u : 0u);   // Sample.cpp:278-288
    std::ostringstream m;
    m << n_reads_ << " read parsed in total for " << times_.size() << " serially sampled samples.";
    messages.push_back(m.str());
}

void Alignment::tile_loci(uint32_t loci_size) {
    if (loci_size == 0 || loci_size > n_sites_) throw std::runtime_error("loci-size out of range");
    loci_.clear();
    uint32_t left = 1, right = loci_size;
    while (right <= n_sites_) { loci_.push_back({left, right}); left = right + 1; right = left + loci_size - 1; }
    if (loci_.back().second < n_sites_) loci_.push_back({loci_.back().second + 1, (uint32_t)n_sites_});
    original_loci_ = loci_;
}

void Alignment::parse_coordinates(const std::string& spec) {
    loci_.clear();
    if (spec.find('[') != std::string::npos) {
        for (std::string tok : split(spec, ",")) {
            if (tok.size() < 3 || tok.front() != '[' || tok.back() != ']') throw std::runtime_error("bad coordinate list");
            tok = tok.substr(1, tok.size() - 2);
            std::vector<std::string> lr = split(tok, "-");
            if (lr.size() != 2) throw std::runtime_error("bad coordinate list");
            Locus l{(uint32_t)std::atof(lr[0].c_str()), (uint32_t)std::atof(lr[1].c_str())};
            if (l.first > l.second || (!loci_.empty() && loci_.back().second >= l.first)) throw std::runtime_error("loci must be ascending and disjoint");
            loci_.push_back(l);
        }
    } else {
        std::vector<uint32_t> bp;
        for (const std::string& tok : split(spec, ",")) {
            bp.push_back((uint32_t)std::atof(tok.c_str()));
            if (bp.back() <= 1 || bp.back() > n_sites_) throw std::runtime_error("breakpoint out of range");
            if (bp.size() == 1) loci_.push_back({1, bp[0] - 1}); else loci_.push_back({bp[bp.size() - 2], bp.back() - 1});
        }
        loci_.push_back({bp.back(), (uint32_t)n_sites_});
    }
    if (loci_.empty() || loci_.front().first == 0 || loci_.back().second > n_sites_) throw std::runtime_error("loci outside the alignment");
    original_loci_ = loci_;
}

size_t Alignment::apply_filter(const std::vector<uint8_t>& keep) {
    // cumulative count of kept sites (Sample.cpp:472-508)
    std::vector<uint32_t> perm(n_sites_);
    uint32_t kept = 0;
    for (size_t k = 0; k < n_sites_; k++) { if (keep[k]) kept++; perm[k] = kept; }
    // rebuild the distinct rows of every sampling time, merging rows that became equal (Sample.cpp:476-520)
    for (size_t t = 0; t < rows_.size(); t++) {
        std::unordered_map<std::string, size_t, RowHash> index;
        std::vector<uint8_t> rows; std::vector<uint32_t> mult;
        const size_t n = mult_[t].size();
        std::string coded(kept, '\0');
        for (size_t r = 0; r < n; r++) {
            const uint8_t* src = rows_[t].data() + r * n_sites_;
            size_t o = 0;
            for (size_t k = 0; k < n_sites_; k++) if (keep[k]) coded[o++] = (char)src[k];
            auto it = index.find(coded);
            if (it == index.end()) { index.insert({coded, mult.size()}); rows.insert(rows.end(), coded.begin(), coded.end()); mult.push_back(mult_[t][r]); }
            else mult[it->second] += mult_[t][r];
        }
        rows_[t] = std::move(rows); mult_[t] = std::move(mult);
    }
    // remap the loci (Sample.cpp:522-557): a locus that lost all its sites disappears
    std::vector<Locus> nl, no;
    for (size_t i = 0; i < loci_.size(); i++) {
        Locus l = loci_[i];
        const uint32_t f = keep[l.first - 1] ? perm[l.first - 1] : perm[l.first - 1] + 1;
        const uint32_t s = perm[l.second - 1];
        if (f > s) {
            std::ostringstream m;
            m << "WARNING: Removed loci (" << original_loci_[i].first << ", " << original_loci_[i].second << ") due to a lack of relevant sites.";
            messages.push_back(m.str());
            continue;
        }
        nl.push_back({f, s}); no.push_back(original_loci_[i]);
    }
    loci_ = nl; original_loci_ = no;
    auto squeeze8 = [&](std::vector<uint8_t>& v) { std::vector<uint8_t> o; for (size_t k = 0; k < n_sites_; k++) if (keep[k]) o.push_back(v[k]); v.swap(o); };
    squeeze8(has_missing_); squeeze8(has_indel_); squeeze8(has_ambiguous_);
    { std::vector<uint32_t> o; for (size_t k = 0; k < n_sites_; k++) if (keep[k]) o.push_back(poly_count_[k]); poly_count_.swap(o); }
    n_sites_ = kept;
    return kept;
}

size_t Alignment::filter_ambiguous() {
    std::vector<uint8_t> keep(n_sites_);
    for (size_t k = 0; k < n_sites_; k++) keep[k] = !has_ambiguous_[k];
    return apply_filter(keep);
}
size_t Alignment::filter_indels() {
    std::vector<uint8_t> keep(n_sites_);
    for (size_t k = 0; k < n_sites_; k++) keep[k] = !has_indel_[k];
    return apply_filter(keep);
}
size_t Alignment::filter_non_biallelic() {
    std::vector<uint8_t> keep(n_sites_);
    for (size_t k = 0; k < n_sites_; k++) keep[k] = (poly_count_[k] == 2);
    return apply_filter(keep);
}

PairProblem Alignment::sub_sample(const Locus& a, const Locus& b) const {
    PairProblem p;
    p.LA = a.second - a.first + 1; p.LB = b.second - b.first + 1;
    if (p.LA + p.LB > 64) throw std::runtime_error("locus pair has more than 64 segregating sites");
    p.type_offsets.push_back(0);
    for (size_t t = 0; t < rows_.size(); t++) {
        const size_t start = p.type_words.size();
        const size_t n = mult_[t].size();
        for (size_t r = 0; r < n; r++) {
            const uint8_t* row = rows_[t].data() + r * n_sites_;
            bool okA = true, okB = true;
            for (uint32_t j = a.first - 1; j < a.second; j++) if (row[j] == kMissing) { okA = false; break; }
            for (uint32_t j = b.first - 1; j < b.second; j++) if (row[j] == kMissing) { okB = false; break; }
            if (!okA && !okB) continue;
            uint64_t w = 0;
            if (okA) for (uint32_t s = 0; s < p.LA; s++) if (row[a.first - 1 + s]) w |= 1ull << s;
            if (okB) for (uint32_t s = 0; s < p.LB; s++) if (row[b.first - 1 + s]) w |= 1ull << (p.LA + s);
            const uint8_t g = (okA && okB) ? 2 : (okA ? 0 : 1);
            bool hit = false;
            for (size_t i = start; i < p.type_words.size(); i++)
                if (p.type_gamma[i] == g && p.type_words[i] == w) { p.type_counts[i] += mult_[t][r]; hit = true; break; }
            if (!hit) { p.type_gamma.push_back(g); p.type_words.push_back(w); p.type_counts.push_back(mult_[t][r]); }
            (g == 2 ? p.nC : g == 0 ? p.nA : p.nB) += mult_[t][r];
        }
        if (p.type_words.size() > start) {      // sampling times without a usable sequence are dropped (Sample.cpp:669-674)
            p.type_offsets.push_back((int32_t)p.type_words.size());
            p.times.push_back(times_[t]); p.viral.push_back(viral_[t]);
        }
    }
    if (p.times.empty()) throw std::runtime_error("All sequences in both loci have missing (increase --loci-size or specify another region)");
    return p;
}

}  // namespace coalhost
