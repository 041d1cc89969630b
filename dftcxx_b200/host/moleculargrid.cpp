#include "moleculargrid.hpp"

#include <chrono>
#include <cstdio>
#include <iostream>
#include <stdexcept>

namespace dftcxx {

MolecularGrid::MolecularGrid(const std::shared_ptr<Molecule>& mol_, int device_, bool verbose_, int ngpus_)
    : mol(mol_), device(device_), ngpus(ngpus_ < 1 ? 1 : ngpus_), verbose(verbose_) {}

MolecularGrid::~MolecularGrid() {
    if (handle) dftgrid_destroy(handle);
}

void MolecularGrid::check(int rc) const {
    if (rc != 0) throw std::runtime_error(dftgrid_last_error());
}

void MolecularGrid::set_grid_parameters(unsigned int radial_points_, unsigned int lebedev_order_, unsigned int lmax_) {
    radial_points = radial_points_;
    lebedev_order = lebedev_order_;
    lmax = lmax_;
}

void MolecularGrid::create_grid() {
    const auto start = std::chrono::system_clock::now();
    if (verbose) {
        std::cout << "      Constructing molecular grid       " << std::endl;
        std::cout << "========================================" << std::endl;
        std::cout << "Number of radial points: " << radial_points << std::endl;
        std::cout << "Lebedev order: " << lebedev_order << std::endl;
        std::cout << "Lmax value: " << lmax << std::endl;
    }
    // flatten Molecule -> dftgrid_system
    const unsigned int na = mol->get_nr_atoms(), nb = mol->get_nr_bfs();
    std::vector<int> Z(na), bf_nprim(nb), lmn;
    std::vector<double> xyz(3 * na), bf_center(3 * nb), alpha, coeff, norm;
    for (unsigned int i = 0; i < na; i++) {
        Z[i] = (int)mol->get_atomic_charge(i);
        for (int d = 0; d < 3; d++) xyz[3 * i + d] = mol->get_atomic_position(i)[d];
    }
    for (unsigned int b = 0; b < nb; b++) {
        const CGF& c = mol->get_cgf(b);
        bf_nprim[b] = (int)c.size();
        for (int d = 0; d < 3; d++) bf_center[3 * b + d] = c.get_position()[d];
        for (unsigned int g = 0; g < c.size(); g++) {
            const GTO& gto = c.get_gto(g);
            alpha.push_back(gto.get_alpha());
            coeff.push_back(gto.get_coefficient());
            norm.push_back(gto.get_norm());
            lmn.push_back((int)gto.get_l());
            lmn.push_back((int)gto.get_m());
            lmn.push_back((int)gto.get_n());
        }
    }
    dftgrid_system sys{(int)na, Z.data(), xyz.data(), (int)nb, bf_nprim.data(), bf_center.data(), (int)alpha.size(),
                       alpha.data(), coeff.data(), norm.data(), lmn.data()};
    dftgrid_params prm{(int)radial_points, (int)lebedev_order, (int)lmax};
    if (handle) {
        dftgrid_destroy(handle);
        handle = nullptr;
    }
    if (ngpus > 1) {
        std::vector<int> devs(ngpus);
        for (int i = 0; i < ngpus; i++) devs[i] = device + i;
        check(dftgrid_create_multi(&handle, &sys, &prm, ngpus, devs.data()));  // one handle, ngpus devices, this process
    } else {
        check(dftgrid_create(&handle, &sys, &prm, device, 0, 1));
    }
    check(dftgrid_build(handle));
    const auto elapsed = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::system_clock::now() - start);
    if (verbose) {
        std::printf("Total time: %ld ms\n", (long)elapsed.count());
        std::cout << "========================================" << std::endl << std::endl;
    }
}

size_t MolecularGrid::get_grid_size() const { return handle ? (size_t)dftgrid_npoints(handle) : 0; }

void MolecularGrid::set_density(const Mat& P) {
    if (!handle) throw std::runtime_error("create_grid has not been called");
    check(dftgrid_set_density(handle, P.data()));  // P is symmetric: row-major here == column-major there
}

void MolecularGrid::correct_densities() {
    // folded into set_density on the device: the reference always calls the two back to back (src/dft.cpp:362-365)
}

double MolecularGrid::calculate_density() const {
    double n = 0.0;
    check(dftgrid_electron_count(handle, &n));
    return n;
}

Mat MolecularGrid::calculate_hartree_potential() {
    const unsigned int nb = mol->get_nr_bfs();
    Mat J(nb, nb);
    check(dftgrid_hartree_J(handle, J.data()));
    return J;
}

Mat MolecularGrid::calculate_exchange_correlation(double& exc) {
    const unsigned int nb = mol->get_nr_bfs();
    Mat XC(nb, nb);
    check(dftgrid_xc(handle, XC.data(), &exc));
    return XC;
}

void MolecularGrid::fock(const Mat& P, bool include_xc, Mat& F, double& e_j, double& exc, double& nelec) {
    if (!handle) throw std::runtime_error("create_grid has not been called");
    check(dftgrid_fock(handle, P.data(), include_xc ? 1 : 0, F.data(), &e_j, &exc, &nelec));
}

void MolecularGrid::scf_init(const Mat& H, const Mat& X, unsigned int nocc, double alpha) {
    if (!handle) throw std::runtime_error("create_grid has not been called");
    check(dftgrid_scf_init(handle, H.data(), X.data(), (int)nocc, alpha));
}

void MolecularGrid::scf_step(bool include_xc, double out8[8]) { check(dftgrid_scf_step(handle, include_xc ? 1 : 0, out8)); }

Mat MolecularGrid::scf_matrix(int which) const {
    const unsigned int nb = mol->get_nr_bfs();
    Mat M(nb, nb);
    check(dftgrid_scf_get_matrix(handle, which, M.data()));
    return M;
}

void MolecularGrid::one_electron(Mat& S, Mat& T, Mat& V) {
    const size_t n = mol->get_nr_bfs();
    S = Mat(n, n);
    T = Mat(n, n);
    V = Mat(n, n);
    check(dftgrid_one_electron(handle, S.data(), T.data(), V.data()));
}

void MolecularGrid::rectangular_density(double size, unsigned int dp, const Mat& P, double* pos, double* rho, double* grad) {
    check(dftgrid_rectangular_density(handle, size, (int)dp, P.data(), pos, rho, grad));
}

int MolecularGrid::gpus() const { return handle ? dftgrid_ngpus(handle) : ngpus; }

std::vector<double> MolecularGrid::get_weights() const {
    std::vector<double> w(get_grid_size());
    check(dftgrid_get_weights(handle, w.data()));
    return w;
}

std::vector<double> MolecularGrid::get_densities() const {
    std::vector<double> r(get_grid_size());
    check(dftgrid_get_densities(handle, r.data()));
    return r;
}

Mat MolecularGrid::get_amplitudes() const {
    const size_t np = get_grid_size(), nb = mol->get_nr_bfs();
    std::vector<double> phi(np * nb);
    check(dftgrid_get_amplitudes(handle, phi.data()));
    Mat A(nb, np);
    for (size_t p = 0; p < np; p++)
        for (size_t b = 0; b < nb; b++) A(b, p) = phi[p * nb + b];
    return A;
}

}  // namespace dftcxx
