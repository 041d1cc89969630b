// Atoms and the contracted Cartesian Gaussian basis: the reference's Atom / GTO / CGF / Molecule surface
// (src/molecule.h:111-175, src/cgf.h:41-314) on plain arrays.  Basis functions are appended element by element
// in basis-file order and, per element, atom by atom (src/molecule.cpp:222-235) — every matrix in the program
// uses that ordering.
#pragma once
#include <array>
#include <memory>
#include <string>
#include <vector>

#include "settings.hpp"

namespace dftcxx {

typedef std::array<double, 3> vec3;

class GTO {  // primitive Cartesian Gaussian N x^l y^m z^n exp(-alpha r^2)
public:
    GTO(double c, const vec3& position, double alpha, unsigned int l, unsigned int m, unsigned int n);
    double get_coefficient() const { return c; }
    double get_alpha() const { return alpha; }
    unsigned int get_l() const { return l; }
    unsigned int get_m() const { return m; }
    unsigned int get_n() const { return n; }
    double get_norm() const { return norm; }
    const vec3& get_position() const { return position; }
    double get_amp(const vec3& r) const;
    void set_position(const vec3& p) { position = p; }

private:
    double c, alpha;
    unsigned int l, m, n;
    vec3 position;
    double norm;
};

class CGF {  // contracted Gaussian function
public:
    enum { GTO_S, GTO_PX, GTO_PY, GTO_PZ, GTO_DX2, GTO_DXY, GTO_DXZ, GTO_DY2, GTO_DYZ, GTO_DZ2, NUM_GTO };
    CGF() : r{0, 0, 0} {}
    explicit CGF(const vec3& r_) : r(r_) {}
    unsigned int size() const { return (unsigned int)gtos.size(); }
    double get_norm_gto(unsigned int i) const { return gtos[i].get_norm(); }
    double get_coefficient_gto(unsigned int i) const { return gtos[i].get_coefficient(); }
    const GTO& get_gto(unsigned int i) const { return gtos[i]; }
    const vec3& get_position() const { return r; }
    double get_amp(const vec3& p) const;
    void add_gto(unsigned int type, double alpha, double c, const vec3& origin);
    void set_position(const vec3& pos);

private:
    std::vector<GTO> gtos;
    vec3 r;
};

class Atom {
public:
    Atom() : atnr(0), position{0, 0, 0} {}
    Atom(unsigned int atnr_, const vec3& p) : atnr(atnr_), position(p) {}
    const vec3& get_position() const { return position; }
    unsigned int get_charge() const { return atnr; }

private:
    unsigned int atnr;
    vec3 position;
};

class Molecule {
public:
    Molecule(const std::string& filename, const std::shared_ptr<Settings>& settings, bool verbose = true);
    unsigned int get_nr_atoms() const { return (unsigned int)atoms.size(); }
    unsigned int get_nr_bfs() const { return (unsigned int)cgfs.size(); }
    const CGF& get_cgf(unsigned int i) const { return cgfs[i]; }
    const std::vector<CGF>* get_cgfs() const { return &cgfs; }
    const std::shared_ptr<Atom>& get_atom(unsigned int i) const { return atoms[i]; }
    const vec3& get_atomic_position(unsigned int i) const { return atoms[i]->get_position(); }
    unsigned int get_atomic_charge(unsigned int i) const { return atoms[i]->get_charge(); }
    unsigned int get_nr_elec() const;
    unsigned int get_nr_gtos() const;
    void add_atom(const Atom& a) { atoms.emplace_back(std::make_shared<Atom>(a)); }
    void add_cgf(unsigned int /*atid*/, const CGF& cgf) { cgfs.push_back(cgf); }
    void set_basis_set(const std::string& basis_set);  // "basis/<name>.dat"

private:
    void read_molecule_from_file(const std::string& filename, bool verbose);
    static unsigned int atom_number_from_string(const std::string& el);
    static std::string locate_basis_file(const std::string& basis_set);

    std::vector<std::shared_ptr<Atom>> atoms;
    std::vector<CGF> cgfs;
    std::shared_ptr<Settings> settings;
};

}  // namespace dftcxx
