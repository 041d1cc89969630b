#include "molecule.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <stdexcept>

#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

namespace dftcxx {

static bool file_exists(const std::string& p) {
    struct stat st;
    return ::stat(p.c_str(), &st) == 0;
}

// ---- GTO / CGF ---------------------------------------------------------------------------------------------
static double odd_double_factorial(unsigned int k) {  // (2k-1)!! with (−1)!! = 1
    double r = 1.0;
    for (unsigned int i = 2 * k; i > 1; i -= 2) r *= (double)(i - 1);
    return r;
}

GTO::GTO(double c_, const vec3& position_, double alpha_, unsigned int l_, unsigned int m_, unsigned int n_)
    : c(c_), alpha(alpha_), l(l_), m(m_), n(n_), position(position_) {
    // <GTO|GTO> = 1.  The reference evaluates this with pi truncated to 3.14159265359 (src/cgf.cpp:102-114);
    // the truncation is kept because every amplitude and integral carries it.
    const double pi_ref = 3.14159265359;
    const unsigned int L = l + m + n;
    const double nom = std::pow(2.0, 2.0 * L + 3.0 / 2.0) * std::pow(alpha, L + 3.0 / 2.0);
    const double denom = odd_double_factorial(l) * odd_double_factorial(m) * odd_double_factorial(n) * std::pow(pi_ref, 3.0 / 2.0);
    norm = std::sqrt(nom / denom);
}

double GTO::get_amp(const vec3& r) const {
    const double dx = r[0] - position[0], dy = r[1] - position[1], dz = r[2] - position[2];
    const double r2 = dx * dx + dy * dy + dz * dz;
    return norm * std::pow(dx, (double)l) * std::pow(dy, (double)m) * std::pow(dz, (double)n) * std::exp(-alpha * r2);
}

double CGF::get_amp(const vec3& p) const {
    double sum = 0.0;
    for (const GTO& g : gtos) sum += g.get_coefficient() * g.get_amp(p);
    return sum;
}

void CGF::add_gto(unsigned int type, double alpha, double c, const vec3& /*origin*/) {
    static const unsigned int lmn[NUM_GTO][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {2, 0, 0},
                                                 {1, 1, 0}, {1, 0, 1}, {0, 2, 0}, {0, 1, 1}, {0, 0, 2}};
    if (type >= NUM_GTO) {
        std::cerr << "Undefined orbital type. Exiting..." << std::endl;
        std::exit(-1);
    }
    gtos.emplace_back(c, r, alpha, lmn[type][0], lmn[type][1], lmn[type][2]);
}

void CGF::set_position(const vec3& pos) {
    r = pos;
    for (GTO& g : gtos) g.set_position(pos);
}

// ---- Molecule ----------------------------------------------------------------------------------------------
Molecule::Molecule(const std::string& filename, const std::shared_ptr<Settings>& settings_, bool verbose) : settings(settings_) {
    read_molecule_from_file(filename, verbose);
}

unsigned int Molecule::atom_number_from_string(const std::string& el) {
    static const char* names[] = {"H", "He", "Li", "Be", "B", "C", "N", "O", "F", "Ne", "Na", "Mg", "Al", "Si", "P", "S", "Cl", "Ar"};
    for (unsigned int i = 0; i < 18; i++)
        if (el == names[i]) return i + 1;
    throw std::runtime_error("Unknown element: " + el);
}

void Molecule::read_molecule_from_file(const std::string& filename, bool verbose) {
    const double angstrom_to_bohr = 1.889725989;
    const bool unit_angstrom = settings->has("units") && settings->get_value("units") == "angstrom";
    const std::string basis_set = "basis/" + settings->get_value("basis") + ".dat";
    if (verbose) {
        std::cout << "           Reading input file           " << std::endl;
        std::cout << "========================================" << std::endl;
        std::cout << "Reading file:\t\t" << filename << std::endl << std::endl;
        std::cout << "System name: " << settings->get_value("name") << std::endl;
        std::cout << "Basis set: " << basis_set << std::endl << std::endl;
    }
    std::ifstream in(filename);
    std::string line;
    while (std::getline(in, line))
        if (line.compare(0, 7, "system:") == 0) break;
    std::getline(in, line);
    unsigned int nratoms = 0;
    if (!parse_uint(line, nratoms)) throw std::runtime_error("bad lexical cast: number of atoms '" + line + "'");
    if (verbose) std::cout << "Atoms in system:\t" << nratoms << std::endl;
    for (unsigned int i = 0; i < nratoms; i++) {
        std::getline(in, line);
        const std::vector<std::string> p = split_compress(line, " \t");
        if (p.size() < 4) throw std::runtime_error("malformed atom line: '" + line + "'");
        const unsigned int atnr = atom_number_from_string(p[0]);
        double x = parse_double(p[1]), y = parse_double(p[2]), z = parse_double(p[3]);
        if (unit_angstrom) {
            x *= angstrom_to_bohr;
            y *= angstrom_to_bohr;
            z *= angstrom_to_bohr;
        }
        add_atom(Atom(atnr, vec3{x, y, z}));
    }
    set_basis_set(basis_set);
    if (verbose) {
        std::cout << "========================================" << std::endl;
        for (unsigned int i = 0; i < nratoms; i++) {
            char buf[128];
            std::snprintf(buf, sizeof buf, "%u  %12.6f  %12.6f  %12.6f", atoms[i]->get_charge(), atoms[i]->get_position()[0],
                          atoms[i]->get_position()[1], atoms[i]->get_position()[2]);
            std::cout << buf << std::endl;
        }
        std::cout << "========================================" << std::endl;
        std::cout << "Total number of GTOs: " << get_nr_gtos() << std::endl << std::endl;
    }
}

// The reference only looks at "../<basis_set>" (it must be run from its build directory, src/molecule.cpp:147-152).
// That location is tried first; $DFTCXX_BASIS_DIR and the data directory shipped next to the executable follow.
std::string Molecule::locate_basis_file(const std::string& basis_set) {
    const std::string ref_path = "../" + basis_set;
    if (file_exists(ref_path)) return ref_path;
    const std::string leaf = basis_set.substr(basis_set.find('/') + 1);
    if (const char* env = std::getenv("DFTCXX_BASIS_DIR")) {
        const std::string p = std::string(env) + "/" + leaf;
        if (file_exists(p)) return p;
    }
    std::vector<std::string> homes;  // directory of the executable, and of the shared object this code lives in
    char exe[4096];
    const ssize_t n = ::readlink("/proc/self/exe", exe, sizeof exe - 1);
    if (n > 0) homes.emplace_back(exe, (size_t)n);
    Dl_info info;
    if (dladdr((const void*)&file_exists, &info) && info.dli_fname) homes.emplace_back(info.dli_fname);
    for (std::string dir : homes) {
        dir = dir.substr(0, dir.rfind('/'));
        for (const char* rel : {"/../data/basis/", "/../../data/basis/", "/data/basis/"}) {
            const std::string p = dir + rel + leaf;
            if (file_exists(p)) return p;
        }
    }
    std::cerr << "Please make sure you are running dftcxx from the build directory..." << std::endl;
    throw std::runtime_error("Cannot open ../" + basis_set + "!");
}

void Molecule::set_basis_set(const std::string& basis_set) {
    const vec3 origin{0, 0, 0};
    unsigned int highest_atom = 0;
    for (const auto& a : atoms) highest_atom = std::max(highest_atom, a->get_charge());
    std::ifstream in(locate_basis_file(basis_set));
    std::string line;
    while (std::getline(in, line)) {
        if (line.empty() || line[0] == '#') continue;
        std::vector<std::string> p = split_compress(line, " \t");
        unsigned int atnr = 0, nshell = 0;
        if (p.size() < 2 || !parse_uint(p[0], atnr) || !parse_uint(p[1], nshell)) throw std::runtime_error("malformed basis header: '" + line + "'");
        std::vector<CGF> shell_cgfs;
        for (unsigned int s = 0; s < nshell; s++) {
            std::getline(in, line);
            p = split_compress(line, " \t");
            unsigned int nprim = 0;
            if (p.size() < 2 || p[0].empty() || !parse_uint(p[1], nprim)) throw std::runtime_error("malformed shell line: '" + line + "'");
            const char type = p[0][0];
            const unsigned int first = type == 'S' ? (unsigned)CGF::GTO_S : type == 'P' ? (unsigned)CGF::GTO_PX : (unsigned)CGF::GTO_DX2;
            const unsigned int count = type == 'S' ? 1 : type == 'P' ? 3 : type == 'D' ? 6 : 0;
            const size_t base = shell_cgfs.size();
            shell_cgfs.resize(base + count);
            for (unsigned int g = 0; g < nprim; g++) {
                std::getline(in, line);
                p = split_compress(line, " \t");
                if (p.size() < 3) throw std::runtime_error("malformed primitive line: '" + line + "'");
                const double exponent = parse_double(p[1]), coefficient = parse_double(p[2]);
                for (unsigned int k = 0; k < count; k++) shell_cgfs[base + k].add_gto(first + k, exponent, coefficient, origin);
            }
        }
        for (unsigned int i = 0; i < atoms.size(); i++)
            if (atoms[i]->get_charge() == atnr)
                for (CGF& c : shell_cgfs) {
                    c.set_position(atoms[i]->get_position());
                    add_cgf(i, c);
                }
        if (atnr == highest_atom) break;
    }
}

unsigned int Molecule::get_nr_elec() const {
    unsigned int n = 0;
    for (const auto& a : atoms) n += a->get_charge();
    return n;
}

unsigned int Molecule::get_nr_gtos() const {
    unsigned int n = 0;
    for (const CGF& c : cgfs) n += c.size();
    return n;
}

}  // namespace dftcxx
