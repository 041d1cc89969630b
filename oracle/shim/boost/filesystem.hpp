// TEST INFRASTRUCTURE ONLY (oracle/): boost::filesystem::exists stand-in.
#ifndef DFTB200_ORACLE_SHIM_BOOST_FILESYSTEM
#define DFTB200_ORACLE_SHIM_BOOST_FILESYSTEM
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include <sys/stat.h>
namespace boost {
namespace filesystem {
inline bool exists(const std::string& p) {
    struct stat st;
    return ::stat(p.c_str(), &st) == 0;
}
}  // namespace filesystem
}  // namespace boost
#endif
