// MolecularGrid: the reference's grid class surface (src/moleculargrid.h:84-167) as a thin C++ view of the
// B200 grid engine behind the C ABI (include/dftgrid.h).  All numerical work happens in libdftgrid.so's CUDA
// kernels; a failing C-ABI call is rethrown as std::runtime_error, which is what the reference throws.
#pragma once
#include <memory>
#include <vector>

#include "../../include/dftgrid.h"
#include "linalg.hpp"
#include "molecule.hpp"

namespace dftcxx {

class MolecularGrid {
public:
    explicit MolecularGrid(const std::shared_ptr<Molecule>& mol, int device = 0, bool verbose = true, int ngpus = 1);
    ~MolecularGrid();
    MolecularGrid(const MolecularGrid&) = delete;
    MolecularGrid& operator=(const MolecularGrid&) = delete;

    void set_grid_parameters(unsigned int radial_points, unsigned int lebedev_order, unsigned int lmax);
    void create_grid();

    void set_density(const Mat& P);  // rho = 2 phi^T P phi on every grid point
    void correct_densities();        // rescale to the electron count (done on the device together with set_density)
    double calculate_density() const;
    Mat calculate_hartree_potential();                 // J
    Mat calculate_exchange_correlation(double& exc);   // XC matrix + E_xc (DFT::calculate_exchange_correlation_matrix)

    // B200 engine extensions (no reference counterpart): the grid's whole Fock contribution in one contraction, and the
    // device-resident SCF algebra (include/dftgrid.h: dftgrid_fock, dftgrid_scf_*)
    void fock(const Mat& P, bool include_xc, Mat& F, double& e_j, double& exc, double& nelec);
    void scf_init(const Mat& H, const Mat& X, unsigned int nocc, double alpha);
    void scf_step(bool include_xc, double out8[8]);
    Mat scf_matrix(int which) const;  // DFTGRID_SCF_P / DFTGRID_SCF_FGRID / ... (include/dftgrid.h)
    int gpus() const;
    // S, T, V of DFT::construct_matrices on the device (dftgrid_one_electron)
    void one_electron(Mat& S, Mat& T, Mat& V);
    // density and density gradient on the reference's RectangularGrid box (dftgrid_rectangular_density)
    void rectangular_density(double size, unsigned int dp, const Mat& P, double* pos, double* rho, double* grad);

    std::vector<double> get_weights() const;
    std::vector<double> get_densities() const;
    Mat get_amplitudes() const;  // basis functions x grid points, like the reference
    size_t get_grid_size() const;

private:
    void check(int rc) const;
    std::shared_ptr<Molecule> mol;
    dftgrid_t* handle = nullptr;
    unsigned int radial_points = 15, lebedev_order = 7, lmax = 8;
    int device, ngpus;
    bool verbose;
    bool density_pending = false;
};

}  // namespace dftcxx
