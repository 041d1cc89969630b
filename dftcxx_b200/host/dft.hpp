// Closed-shell LDA Kohn-Sham SCF driver: the reference's DFT class (src/dft.h, src/dft.cpp) with the grid work
// delegated to the GPU MolecularGrid.  The canonical orthogonalisation stays on the host; the one-electron integrals come
// from the device unless `integrals = host`; the eigen-solve, density mixing and energy expression run on the device by
// default (scf = device) or on the host exactly as in the reference (scf = host).
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "integrals.hpp"
#include "linalg.hpp"
#include "molecule.hpp"
#include "moleculargrid.hpp"
#include "settings.hpp"

namespace dftcxx {

struct ScfRecord {  // one line of the SCF table
    double et, exc, e_one, e_j, nelec_grid, ms;
    double ms_algebra = 0.0, ms_grid = 0.0;  // device times of the two halves of the iteration (scf = device)
    int purification_steps = 0;
};

class DFT {
public:
    // ngpus / scf_mode < 0: taken from the input file's `gpus` / `scf` / `fock` keys (Settings)
    explicit DFT(const std::string& filename, int device = 0, bool verbose = true, int ngpus = -1, int scf_mode = -1);
    ~DFT();
    void scf(unsigned int max_iterations = 100, double threshold = 1e-4);

    const std::vector<ScfRecord>& history() const { return records; }
    const Mat& overlap_matrix() const { return S; }
    const Mat& core_hamiltonian() const { return H; }
    const Mat& kinetic_matrix() const { return T; }
    const Mat& nuclear_matrix() const { return V; }
    Mat density_matrix() const { return scf_mode == Settings::SCF_DEVICE ? molgrid->scf_matrix(DFTGRID_SCF_P) : P; }
    const Mat& coulomb_matrix() const { return J; }
    const Mat& xc_matrix() const { return XC; }
    double nuclear_repulsion() const { return enuc; }
    double total_energy() const { return et; }
    unsigned int nbf() const { return mol->get_nr_bfs(); }
    // one pass of the loop body (src/dft.cpp:100-103); returns the total energy
    double scf_step();

private:
    void add_molecule();
    void construct_matrices();
    void calculate_nuclear_repulsion();
    void calculate_transformation_matrix();
    void calculate_density_matrix();
    void calculate_electronic_repulsion_matrix();
    void calculate_exchange_correlation_matrix();
    void calculate_energy();
    void finalize();  // density dump on a rectangular grid when the input asks for one (src/dft.cpp:489-504)

    std::shared_ptr<Settings> settings;
    std::shared_ptr<Molecule> mol;
    std::unique_ptr<MolecularGrid> molgrid;
    Integrator integrator;
    const std::vector<CGF>* cgfs = nullptr;
    Mat S, T, V, H, X, Xp, C, P, J, XC, Fg;  // Fg: the grid's Fock contribution 2J + XC (fused modes)
    int ngpus_opt, scf_mode_opt;
    unsigned int scf_mode = Settings::SCF_DEVICE;
    double nelec_grid = 0.0;
    bool pinned = false;
    unsigned int nelec = 0;
    double exc = 0.0, enuc = 0.0, et = 0.0, single_electron_energy = 0.0, electronic_repulsion = 0.0;
    bool is_first = true;
    bool verbose;
    int device;
    std::vector<ScfRecord> records;
};

}  // namespace dftcxx
