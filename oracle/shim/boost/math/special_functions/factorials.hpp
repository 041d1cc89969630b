// TEST INFRASTRUCTURE ONLY (oracle/): boost::math::factorial / double_factorial
// stand-ins (exact products in double, as boost's lookup tables are).
#ifndef DFTB200_ORACLE_SHIM_BOOST_FACTORIALS
#define DFTB200_ORACLE_SHIM_BOOST_FACTORIALS
#include <fstream>
#include <iostream>
#include <sstream>
#include <vector>
namespace boost {
namespace math {
template <typename T>
inline T factorial(unsigned n) {
    T r = 1;
    for (unsigned i = 2; i <= n; i++) r *= (T)i;
    return r;
}
template <typename T>
inline T double_factorial(unsigned n) {
    T r = 1;
    for (unsigned i = n; i > 1; i -= 2) r *= (T)i;
    return r;
}
}  // namespace math
}  // namespace boost
#endif
