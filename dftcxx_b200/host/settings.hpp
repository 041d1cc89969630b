// Input settings of a dftcxx run: the `key = value` lines that precede `system:` in a .in file, plus the grid
// presets.  Same keys, defaults and precedence as the reference (src/settings.cpp:39-187): first occurrence of a
// key wins, `grid` in {coarse, medium (default), fine, ultrafine}, optional overrides radial_points /
// lebedev_order / lmax (ignored when not an unsigned integer), hartree_evaluation defaults to becke_grid.
#pragma once
#include <string>
#include <unordered_map>
#include <vector>

namespace dftcxx {

class Settings {
public:
    enum { BECKE_GRID, TWO_ELECTRON_INTEGRALS };
    enum { GRID_COARSE, GRID_MEDIUM, GRID_FINE, GRID_ULTRAFINE };

    explicit Settings(const std::string& filename);
    static Settings from_text(const std::string& text);

    const std::string& get_value(const std::string& key) const;  // throws std::logic_error when absent
    bool has(const std::string& key) const { return key_values.count(key) != 0; }
    unsigned int get_hartree_evaluation_method() const { return hartree_evaluation; }
    unsigned int get_radial_points() const { return radial_points; }
    unsigned int get_lebedev_order() const { return lebedev_order; }
    unsigned int get_lmax() const { return lmax; }
    // B200 engine keys (absent from stock inputs, which therefore run unchanged): `gpus = N` devices of this box driven
    // from the one process (default 1), `scf = device|host` where the SCF algebra runs (default device), `fock =
    // fused|separate` whether the host path asks the grid for F_grid = 2J + XC in one contraction or for J and XC
    unsigned int get_gpus() const { return gpus; }
    enum { SCF_DEVICE, SCF_HOST_FUSED, SCF_HOST_SEPARATE };
    unsigned int get_scf_mode() const { return scf_mode; }
    // `integrals = device|host`: where S, T, V of DFT::construct_matrices are evaluated (default device)
    bool get_integrals_on_device() const { return integrals_device; }
    // density dump of DFT::finalize (src/dft.cpp:489-504): `density_dump = <file>` switches it on; the two arguments of
    // RectangularGrid::build_grid default to the reference's own 5.0 / 15 (`density_dump_size`, `density_dump_points`)
    double get_density_dump_size() const { return dump_size; }
    unsigned int get_density_dump_points() const { return dump_points; }

private:
    Settings() {}
    void parse(std::istream& in);
    void set_default_settings();
    void set_grid_fineness(unsigned int fineness);

    std::unordered_map<std::string, std::string> key_values;
    unsigned int radial_points = 15, lebedev_order = 7, lmax = 8;
    unsigned int hartree_evaluation = BECKE_GRID;
    unsigned int gpus = 1, scf_mode = SCF_DEVICE;
    bool integrals_device = true;
    double dump_size = 5.0;
    unsigned int dump_points = 15;
};

// shared text helpers (boost::split(token_compress_on) / trim / strict lexical_cast semantics)
std::vector<std::string> split_compress(const std::string& line, const std::string& seps);
std::string trimmed(const std::string& s);
bool parse_uint(const std::string& s, unsigned int& out);
double parse_double(const std::string& s);  // throws std::runtime_error on junk

}  // namespace dftcxx
