// Stand-in for <boost/math/tools/minima.hpp> -- TEST INFRASTRUCTURE ONLY: brent_find_minima(f, min, max, bits) as called
// at /root/reference/ACFcalculator.cpp:784 and :810, answered by this repo's restatement in csrc/host/gle.cpp (see
// roots.hpp in this directory for what that does and does not prove).
#pragma once
#include <functional>
#include <utility>

#include "gle.hpp"

namespace boost {
namespace math {
namespace tools {

template <class F, class T>
std::pair<T, T> brent_find_minima(F f, T min, T max, int bits)
{
    return acfhost::minimise_brent(std::function<double(double)>(f), min, max, bits);
}

}  // namespace tools
}  // namespace math
}  // namespace boost
