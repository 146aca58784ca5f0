// Test infrastructure (oracle/): boost::math::lgamma stand-in (ProposalEngine.cpp:95-113),
// arguments there are small positive integers + 1.
#pragma once
#include <cmath>
namespace boost { namespace math {
inline double lgamma(double x) { int sign; return ::lgamma_r(x, &sign); }
}}
