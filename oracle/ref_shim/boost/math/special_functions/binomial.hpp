// Test infrastructure (oracle/): boost::math::binomial_coefficient<double> stand-in
// (used at ConditionalSamplingHMM.cpp:79, ProposalEngine.cpp:73-82). Boost.Math is not
// installed and the reference does not pin a version (README.md:9). Boost returns the
// quotient of factorials rounded to the nearest integer (ceil(x - 0.5)); here the product
// form is evaluated in long double (exact while the value fits 64 bits) and rounded the same way.
#pragma once
#include <cmath>
namespace boost { namespace math {
template <class T>
inline T binomial_coefficient(unsigned n, unsigned k) {
    if (k > n) return std::nan("");
    if (k == 0 || k == n) return static_cast<T>(1);
    if (k == 1 || k == n - 1) return static_cast<T>(n);
    unsigned kk = (k < n - k) ? k : n - k;
    long double r = 1.0L;
    for (unsigned i = 1; i <= kk; i++) {
        r = r * static_cast<long double>(n - kk + i) / static_cast<long double>(i);
    }
    return static_cast<T>(std::ceil(static_cast<double>(r) - 0.5));
}
}}
