// TEST INFRASTRUCTURE ONLY (oracle/): boost::regex → std::regex aliases.
#ifndef DFTB200_ORACLE_SHIM_BOOST_REGEX
#define DFTB200_ORACLE_SHIM_BOOST_REGEX
#include <regex>
namespace boost {
using std::regex;
using std::regex_match;
using std::smatch;
}  // namespace boost
#endif
