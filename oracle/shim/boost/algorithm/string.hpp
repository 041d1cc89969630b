// TEST INFRASTRUCTURE ONLY (oracle/): boost::split / is_any_of / trim stand-ins.
// token_compress_on merges adjacent separators but (like boost) still yields an
// empty first token when the input starts with a separator.
#ifndef DFTB200_ORACLE_SHIM_BOOST_ALGO_STRING
#define DFTB200_ORACLE_SHIM_BOOST_ALGO_STRING
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include <vector>
namespace boost {
enum token_compress_mode_type { token_compress_on, token_compress_off };
struct is_any_of {
    std::string set;
    explicit is_any_of(const std::string& s) : set(s) {}
    bool operator()(char c) const { return set.find(c) != std::string::npos; }
};
template <typename Pred>
std::vector<std::string>& split(std::vector<std::string>& out, const std::string& in, Pred pred,
                                token_compress_mode_type mode = token_compress_off) {
    out.clear();
    std::string cur;
    size_t i = 0;
    while (i < in.size()) {
        if (pred(in[i])) {
            out.push_back(cur);
            cur.clear();
            i++;
            if (mode == token_compress_on)
                while (i < in.size() && pred(in[i])) i++;
        } else {
            cur += in[i++];
        }
    }
    out.push_back(cur);
    return out;
}
inline void trim(std::string& s) {
    size_t b = 0, e = s.size();
    while (b < e && isspace((unsigned char)s[b])) b++;
    while (e > b && isspace((unsigned char)s[e - 1])) e--;
    s = s.substr(b, e - b);
}
}  // namespace boost
#endif
