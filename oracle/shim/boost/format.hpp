// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for the boost::format subset the
// dftcxx reference uses (%i %3i %f %9.7f %4.2f %12.6f, fed through operator%).
// Like boost, a directive only sets stream width/precision/fixed flags and the
// argument is then streamed with operator<<, so an integer fed to %f prints as an integer.
#ifndef DFTB200_ORACLE_SHIM_BOOST_FORMAT
#define DFTB200_ORACLE_SHIM_BOOST_FORMAT
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
namespace boost {
class format {
    struct Piece {
        std::string lit;
        bool directive;
        int width, prec;
        char conv;
    };
    std::vector<Piece> pieces;
    std::vector<std::string> args;
    size_t next_directive(size_t from) const {
        for (size_t i = from; i < pieces.size(); i++)
            if (pieces[i].directive) return i;
        return pieces.size();
    }
    size_t cur;

public:
    explicit format(const std::string& f) : cur(0) {
        std::string lit;
        for (size_t i = 0; i < f.size(); i++) {
            if (f[i] != '%') {
                lit += f[i];
                continue;
            }
            if (i + 1 < f.size() && f[i + 1] == '%') {
                lit += '%';
                i++;
                continue;
            }
            if (!lit.empty()) pieces.push_back(Piece{lit, false, 0, -1, 0});
            lit.clear();
            int width = 0, prec = -1;
            i++;
            while (i < f.size() && isdigit((unsigned char)f[i])) width = width * 10 + (f[i++] - '0');
            if (i < f.size() && f[i] == '.') {
                prec = 0;
                i++;
                while (i < f.size() && isdigit((unsigned char)f[i])) prec = prec * 10 + (f[i++] - '0');
            }
            char conv = i < f.size() ? f[i] : 's';
            pieces.push_back(Piece{"", true, width, prec, conv});
        }
        if (!lit.empty()) pieces.push_back(Piece{lit, false, 0, -1, 0});
        cur = next_directive(0);
    }
    template <typename T>
    format& operator%(const T& v) {
        if (cur < pieces.size()) {
            std::ostringstream os;
            const Piece& p = pieces[cur];
            if (p.width > 0) os << std::setw(p.width);
            if (p.conv == 'f') os << std::fixed << std::setprecision(p.prec >= 0 ? p.prec : 6);
            os << v;
            pieces[cur].lit = os.str();
            cur = next_directive(cur + 1);
        }
        return *this;
    }
    std::string str() const {
        std::string s;
        for (const auto& p : pieces) s += p.lit;
        return s;
    }
};
inline std::ostream& operator<<(std::ostream& os, const format& f) { return os << f.str(); }
}  // namespace boost
#endif
