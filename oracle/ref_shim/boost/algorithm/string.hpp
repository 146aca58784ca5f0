// Test infrastructure (oracle/): boost::split / boost::is_any_of stand-ins for the input
// parser (Sample.cpp:88-90,307-348). Empty tokens are kept, as Boost's default does.
#pragma once
#include <string>
#include <vector>
namespace boost {
struct any_of_pred { std::string set; bool operator()(char c) const { return set.find(c) != std::string::npos; } };
inline any_of_pred is_any_of(const std::string& s) { return any_of_pred{s}; }
template <class Pred>
inline void split(std::vector<std::string>& out, const std::string& in, Pred p) {
    out.clear();
    std::string cur;
    for (char c : in) {
        if (p(c)) { out.push_back(cur); cur.clear(); } else { cur.push_back(c); }
    }
    out.push_back(cur);
}
}
