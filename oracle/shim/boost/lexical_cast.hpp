// TEST INFRASTRUCTURE ONLY (oracle/): strict boost::lexical_cast stand-in.
#ifndef DFTB200_ORACLE_SHIM_BOOST_LEXICAL_CAST
#define DFTB200_ORACLE_SHIM_BOOST_LEXICAL_CAST
#include <sstream>
#include <stdexcept>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
namespace boost {
struct bad_lexical_cast : public std::runtime_error {
    bad_lexical_cast() : std::runtime_error("bad lexical cast") {}
};
template <typename T>
T lexical_cast(const std::string& s) {
    if (s.empty()) throw bad_lexical_cast();
    if (isspace((unsigned char)s.front()) || isspace((unsigned char)s.back())) throw bad_lexical_cast();
    std::istringstream is(s);
    T v;
    is >> v;
    if (is.fail() || !is.eof()) {
        // allow eofbit not yet set when nothing remains
        if (is.fail() || is.peek() != std::char_traits<char>::eof()) throw bad_lexical_cast();
    }
    return v;
}
}  // namespace boost
#endif
