#include <fstream>
#include <sstream>
#include <stdexcept>
#include <vector>

#include "settings.hpp"

#include <sys/stat.h>

namespace dftcxx {

std::vector<std::string> split_compress(const std::string& line, const std::string& seps) {
    std::vector<std::string> out;
    std::string cur;
    size_t i = 0;
    while (i < line.size()) {
        if (seps.find(line[i]) != std::string::npos) {
            out.push_back(cur);
            cur.clear();
            while (i < line.size() && seps.find(line[i]) != std::string::npos) i++;
        } else {
            cur += line[i++];
        }
    }
    out.push_back(cur);
    return out;
}

std::string trimmed(const std::string& s) {
    size_t b = 0, e = s.size();
    while (b < e && isspace((unsigned char)s[b])) b++;
    while (e > b && isspace((unsigned char)s[e - 1])) e--;
    return s.substr(b, e - b);
}

bool parse_uint(const std::string& s, unsigned int& out) {
    if (s.empty()) return false;
    unsigned long v = 0;
    for (char c : s) {
        if (c < '0' || c > '9') return false;
        v = v * 10 + (unsigned long)(c - '0');
        if (v > 0xffffffffUL) return false;
    }
    out = (unsigned int)v;
    return true;
}

double parse_double(const std::string& s) {
    if (s.empty() || isspace((unsigned char)s.front()) || isspace((unsigned char)s.back()))
        throw std::runtime_error("bad lexical cast: '" + s + "'");
    std::istringstream is(s);
    double v;
    is >> v;
    if (is.fail() || is.peek() != std::char_traits<char>::eof()) throw std::runtime_error("bad lexical cast: '" + s + "'");
    return v;
}

Settings::Settings(const std::string& filename) {
    struct stat st;
    if (::stat(filename.c_str(), &st) != 0) throw std::runtime_error("Cannot open " + filename + "!");
    std::ifstream in(filename);
    parse(in);
    set_default_settings();
}

Settings Settings::from_text(const std::string& text) {
    Settings s;
    std::istringstream in(text);
    s.parse(in);
    s.set_default_settings();
    return s;
}

void Settings::parse(std::istream& in) {
    std::string line;
    while (std::getline(in, line)) {
        const std::vector<std::string> pieces = split_compress(line, "=");
        if (pieces.size() == 2) key_values.emplace(trimmed(pieces[0]), trimmed(pieces[1]));
        // the block ends at a line that is exactly "system:" followed by optional white space
        if (line.compare(0, 7, "system:") == 0 && trimmed(line.substr(7)).empty()) break;
    }
}

const std::string& Settings::get_value(const std::string& key) const {
    auto it = key_values.find(key);
    if (it == key_values.end()) throw std::logic_error("Could not find " + key);
    return it->second;
}

void Settings::set_default_settings() {
    const std::string he = has("hartree_evaluation") ? get_value("hartree_evaluation") : "becke_grid";
    hartree_evaluation = he == "two_electron_integrals" ? (unsigned)TWO_ELECTRON_INTEGRALS : (unsigned)BECKE_GRID;

    const std::string grid = has("grid") ? get_value("grid") : "medium";
    if (grid == "coarse")
        set_grid_fineness(GRID_COARSE);
    else if (grid == "fine")
        set_grid_fineness(GRID_FINE);
    else if (grid == "ultrafine")
        set_grid_fineness(GRID_ULTRAFINE);
    else
        set_grid_fineness(GRID_MEDIUM);

    unsigned int v;
    if (has("radial_points") && parse_uint(get_value("radial_points"), v)) radial_points = v;
    if (has("lebedev_order") && parse_uint(get_value("lebedev_order"), v)) lebedev_order = v;
    if (has("lmax") && parse_uint(get_value("lmax"), v)) lmax = v;
    if (has("gpus") && parse_uint(get_value("gpus"), v) && v >= 1) gpus = v;
    integrals_device = !(has("integrals") && get_value("integrals") == "host");
    if (has("density_dump_points") && parse_uint(get_value("density_dump_points"), v) && v >= 2) dump_points = v;
    if (has("density_dump_size")) {
        const double sz = parse_double(get_value("density_dump_size"));
        if (!(sz > 0.0)) throw std::runtime_error("density_dump_size must be positive");
        dump_size = sz;
    }
    const std::string scf = has("scf") ? get_value("scf") : "device";
    const std::string fock = has("fock") ? get_value("fock") : "fused";
    scf_mode = scf == "host" ? (fock == "separate" ? (unsigned)SCF_HOST_SEPARATE : (unsigned)SCF_HOST_FUSED) : (unsigned)SCF_DEVICE;
}

void Settings::set_grid_fineness(unsigned int fineness) {
    // radial points / Lebedev order index (50, 110, 146, 194 points) / lmax
    static const unsigned int presets[4][3] = {{10, 4, 5}, {15, 7, 8}, {20, 8, 10}, {30, 10, 11}};
    const unsigned int* p = presets[fineness < 4 ? fineness : GRID_MEDIUM];
    radial_points = p[0];
    lebedev_order = p[1];
    lmax = p[2];
}

}  // namespace dftcxx
