// TEST INFRASTRUCTURE ONLY (oracle/).  Never linked into, or called by, the product.
//
// C-callable harness around the UNMODIFIED dftcxx reference classes.  The
// reference objects are compiled from the sources where they lie under
// /root/reference/src (see oracle/Makefile) against the header shims in
// oracle/shim/ (the image has no Eigen/Boost/TCLAP); this file only constructs
// the reference's own Settings / Molecule / MolecularGrid / DFT objects, calls
// their own methods and copies the results out as flat FP64 arrays.
//
// `#define private public` is used because the reference exposes no getters for
// AtomicGrid::{rho_lm,U_lm,V,V_fuzzy_cell,grid}, MolecularGrid::atomic_grids
// or DFT's matrices (src/atomicgrid.h:42-57, src/moleculargrid.h:54, src/dft.h:40-71).
// Access specifiers do not change the GCC object layout, so the separately
// compiled reference objects stay binary compatible.

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>
#include <regex>
#include <unistd.h>

#define private public
#include "dft.h"
#undef private

// The reference reads two members before ever writing them: Settings::hartree_evaluation when the input
// says (or defaults to) becke_grid (src/settings.cpp:102-108) and DFT::exc in construct_matrices
// (src/dft.cpp:226,446).  With the stock CLI the heap happens to be zero there (=> BECKE_GRID, the
// documented default); inside a long-lived host process it is not.  Zero-filling every allocation made
// by this library (linked -Bsymbolic, so only this library) makes that behaviour deterministic without
// touching the reference sources.
#include <cstdlib>
#include <new>
void* operator new(std::size_t n) {
    void* p = std::calloc(1, n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void* operator new[](std::size_t n) {
    void* p = std::calloc(1, n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void operator delete(void* p) noexcept { std::free(p); }
void operator delete[](void* p) noexcept { std::free(p); }
void operator delete(void* p, std::size_t) noexcept { std::free(p); }
void operator delete[](void* p, std::size_t) noexcept { std::free(p); }

namespace {
thread_local std::string g_err;

struct CoutSilencer {
    std::streambuf* old;
    std::ostringstream sink;
    bool on;
    explicit CoutSilencer(bool quiet) : old(nullptr), on(quiet) {
        if (on) old = std::cout.rdbuf(sink.rdbuf());
    }
    ~CoutSilencer() {
        if (on) std::cout.rdbuf(old);
    }
};

struct CwdGuard {
    char old[4096];
    bool ok;
    explicit CwdGuard(const char* dir) {
        ok = getcwd(old, sizeof old) != nullptr;
        if (dir && *dir && chdir(dir) != 0) throw std::runtime_error(std::string("cannot chdir to ") + dir);
    }
    ~CwdGuard() {
        if (ok && chdir(old) != 0) { /* nothing sensible to do */ }
    }
};

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
}  // namespace

struct refdft {
    std::shared_ptr<Settings> settings;
    std::shared_ptr<Molecule> mol;
    std::unique_ptr<DFT> dft;             // full mode
    std::unique_ptr<MolecularGrid> grid;  // grid-only mode
    bool quiet;
    double exc_gridonly;
    MolecularGrid* mg() { return dft ? dft->molgrid.get() : grid.get(); }
    Molecule* m() { return dft ? dft->mol.get() : mol.get(); }
};

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

// mode 0: full reference DFT object (integrals, core guess, first Hartree build: DFT::DFT, src/dft.cpp:30-76)
// mode 1: Settings + Molecule + MolecularGrid::create_grid only (src/dft.cpp:57-61)
// rundir must be a directory whose ../basis/ holds the .dat files (src/molecule.cpp:147-152).
refdft* ref_open(const char* infile, const char* rundir, int mode, int quiet) {
    try {
        CoutSilencer sil(quiet != 0);
        CwdGuard cwd(rundir);
        std::unique_ptr<refdft> h(new refdft());
        h->quiet = quiet != 0;
        h->exc_gridonly = 0.0;
        if (mode == 0) {
            h->dft.reset(new DFT(infile));
            h->settings = h->dft->settings;
            h->mol = h->dft->mol;
        } else {
            h->settings = std::make_shared<Settings>(infile);
            h->mol = std::make_shared<Molecule>(infile, h->settings);
            h->grid.reset(new MolecularGrid(h->mol));
            h->grid->set_grid_parameters(h->settings->get_radial_points(), h->settings->get_lebedev_order(),
                                         h->settings->get_lmax());
            h->grid->create_grid();
        }
        return h.release();
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}

void ref_close(refdft* h) { delete h; }

int ref_natoms(refdft* h) { return (int)h->m()->get_nr_atoms(); }
int ref_nbf(refdft* h) { return (int)h->m()->get_nr_bfs(); }
int ref_nprims(refdft* h) { return (int)h->m()->get_nr_gtos(); }
long ref_npoints(refdft* h) { return (long)h->mg()->grid_size; }
int ref_nrad(refdft* h) { return (int)h->mg()->radial_points; }
int ref_nang(refdft* h) { return (int)h->mg()->angular_points; }
int ref_lebedev_order(refdft* h) { return (int)h->mg()->lebedev_order; }
int ref_lmax(refdft* h) { return (int)h->mg()->lmax; }
int ref_nelec(refdft* h) { return (int)h->m()->get_nr_elec(); }

void ref_get_atoms(refdft* h, int* Z, double* xyz) {
    Molecule* m = h->m();
    for (unsigned i = 0; i < m->get_nr_atoms(); i++) {
        Z[i] = (int)m->get_atomic_charge(i);
        for (int c = 0; c < 3; c++) xyz[3 * i + c] = m->get_atomic_position(i)[c];
    }
}

// basis functions in the reference's own order (src/molecule.cpp:222-235)
void ref_get_basis(refdft* h, int* bf_nprim, double* bf_center, double* alpha, double* coeff, double* norm, int* lmn) {
    Molecule* m = h->m();
    size_t k = 0;
    for (unsigned b = 0; b < m->get_nr_bfs(); b++) {
        const CGF& c = m->get_cgf(b);
        bf_nprim[b] = (int)c.size();
        for (unsigned g = 0; g < c.size(); g++, k++) {
            const GTO& gto = c.get_gto(g);
            if (g == 0)
                for (int d = 0; d < 3; d++) bf_center[3 * b + d] = gto.get_position()[d];
            alpha[k] = gto.get_alpha();
            coeff[k] = gto.get_coefficient();
            norm[k] = gto.get_norm();
            lmn[3 * k + 0] = (int)gto.get_l();
            lmn[3 * k + 1] = (int)gto.get_m();
            lmn[3 * k + 2] = (int)gto.get_n();
        }
    }
}

// atom-major, radial-major, angular-minor point order (src/atomicgrid.cpp:50-82)
void ref_get_grid(refdft* h, double* xyz, double* w, double* wbecke) {
    size_t p = 0;
    for (auto& ag : h->mg()->atomic_grids)
        for (const GridPoint& gp : ag->grid) {
            if (xyz)
                for (int c = 0; c < 3; c++) xyz[3 * p + c] = gp.get_position()[c];
            if (w) w[p] = gp.get_weight();
            if (wbecke) wbecke[p] = gp.get_becke_weight();
            p++;
        }
}

// phi[p*nb + b]
void ref_get_amplitudes(refdft* h, double* phi) {
    const size_t nb = h->m()->get_nr_bfs();
    size_t p = 0;
    for (auto& ag : h->mg()->atomic_grids)
        for (const GridPoint& gp : ag->grid) {
            const VectorXd& a = gp.get_basis_func_amp();
            for (size_t b = 0; b < nb; b++) phi[p * nb + b] = a(b);
            p++;
        }
}

static MatrixXXd to_mat(const double* P, size_t nb) {
    MatrixXXd M(nb, nb);
    std::memcpy(M.data(), P, sizeof(double) * nb * nb);
    return M;
}

// DFT::calculate_density_matrix tail: set_density + correct_densities (src/dft.cpp:362-365)
void ref_set_density(refdft* h, const double* P) {
    const size_t nb = h->m()->get_nr_bfs();
    MatrixXXd M = to_mat(P, nb);
    if (h->dft) h->dft->P = M;
    h->mg()->set_density(M);
    h->mg()->correct_densities();
}

// set_density only, then report sum(w rho) before the rescale
double ref_set_density_raw(refdft* h, const double* P) {
    const size_t nb = h->m()->get_nr_bfs();
    h->mg()->set_density(to_mat(P, nb));
    return h->mg()->calculate_density();
}

void ref_get_densities(refdft* h, double* rho) {
    size_t p = 0;
    for (auto& ag : h->mg()->atomic_grids)
        for (const GridPoint& gp : ag->grid) rho[p++] = gp.get_density();
}

double ref_electron_count(refdft* h) { return h->mg()->calculate_density(); }

// MolecularGrid::calculate_hartree_potential (src/moleculargrid.cpp:336-389)
void ref_hartree(refdft* h, double* J) {
    MatrixXXd Jm = h->mg()->calculate_hartree_potential();
    if (h->dft) h->dft->J = Jm;
    if (J) std::memcpy(J, Jm.data(), sizeof(double) * Jm.size());
}

// rho_lm, U_lm: [atom][radial index i (as r_n, descending r)][lm]; V, Vfuzzy: [Npts]; q: per-atom sum(w rho)
void ref_get_hartree_intermediates(refdft* h, double* rho_lm, double* U_lm, double* V, double* Vfuzzy, double* q) {
    size_t p = 0, t = 0, a = 0;
    for (auto& ag : h->mg()->atomic_grids) {
        const size_t nr = ag->rho_lm.rows(), nlm = ag->rho_lm.cols();
        for (size_t i = 0; i < nr; i++)
            for (size_t n = 0; n < nlm; n++, t++) {
                if (rho_lm) rho_lm[t] = ag->rho_lm(i, n);
                if (U_lm) U_lm[t] = ag->U_lm(i, n);
            }
        for (size_t j = 0; j < ag->grid.size(); j++, p++) {
            if (V) V[p] = ag->V(j);
            if (Vfuzzy) Vfuzzy[p] = ag->V_fuzzy_cell(j);
        }
        if (q) q[a] = ag->get_density();
        a++;
    }
}

// evaluate atom k's lm spline exactly as the interpolation loop does (src/atomicgrid.cpp:498-500)
double ref_spline_value(refdft* h, int atom, int lm, double r) { return h->mg()->atomic_grids[atom]->get_sh_value(r, lm); }

// DFT::calculate_exchange_correlation_matrix (src/dft.cpp:394-433).  In grid-only mode there is no DFT
// object, so the same statements are issued here on the reference's own Functional and MolecularGrid
// getters (a transcription of the call sequence, not of the arithmetic, which stays in the reference).
void ref_xc(refdft* h, double* XC, double* exc) {
    const size_t nb = h->m()->get_nr_bfs();
    if (h->dft) {
        h->dft->calculate_exchange_correlation_matrix();
        if (XC) std::memcpy(XC, h->dft->XC.data(), sizeof(double) * nb * nb);
        if (exc) *exc = h->dft->exc;
        return;
    }
    MolecularGrid* g = h->mg();
    VectorXd densities = g->get_densities();
    VectorXd weights = g->get_weights();
    MatrixXXd amplitudes = g->get_amplitudes();
    VectorXd ex, vxa, vxb, ec, vca, vcb;
    VectorXd da = densities * 0.5, db = densities * 0.5;
    std::shared_ptr<Functional> f;  // the reference also calls through a never-constructed pointer (src/dft.h:46)
    f->xalpha_x_functional(da, db, ex, vxa, vxb);
    f->vwm_c_functional(da, db, ec, vca, vcb);
    h->exc_gridonly = weights.dot(ex + ec);
    VectorXd wva = weights.cwiseProduct((vxa + vxb + vca + vcb) * 0.5);
    MatrixXXd X = MatrixXXd::Zero(nb, nb);
    for (size_t i = 0; i < nb; i++) {
        VectorXd row = amplitudes.row(i);
        VectorXd wva_i = wva.cwiseProduct(row);
        for (size_t j = 0; j < nb; j++) X(i, j) = wva_i.dot(amplitudes.row(j));
    }
    if (XC) std::memcpy(XC, X.data(), sizeof(double) * nb * nb);
    if (exc) *exc = h->exc_gridonly;
}

// RectangularGrid::build_grid(size, dp) + set_density(P) (src/rectangulargrid.cpp:34-80): the data behind the density dump of
// DFT::finalize (src/dft.cpp:489-504).  pos [dp^3][3], rho [dp^3], grad [dp^3][3], in the reference's own point order.
int ref_rect_density(refdft* h, double size, int dp, const double* P, double* pos, double* rho, double* grad) {
    try {
        RectangularGrid rg(h->mol);
        rg.build_grid(size, (unsigned)dp);
        rg.set_density(to_mat(P, h->m()->get_nr_bfs()));
        for (size_t i = 0; i < rg.grid.size(); i++) {
            const vec3& r = rg.grid[i].get_position();
            const vec3& g = rg.grid[i].get_gradient();
            for (int c = 0; c < 3; c++) {
                pos[3 * i + c] = r[c];
                grad[3 * i + c] = g[c];
            }
            rho[i] = rg.grid[i].get_density();
        }
        return (int)rg.grid.size();
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

// the reference's own writer (RectangularGrid::write_gradient, src/rectangulargrid.cpp:82-95) for the same grid
int ref_rect_write(refdft* h, double size, int dp, const double* P, const char* filename) {
    try {
        RectangularGrid rg(h->mol);
        rg.build_grid(size, (unsigned)dp);
        rg.set_density(to_mat(P, h->m()->get_nr_bfs()));
        rg.write_gradient(filename);
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

// pointwise functional values straight from the reference's Functional (src/functionals.cpp:24-114)
void ref_functional(const double* rho, long n, double* ex, double* vx, double* ec, double* vc) {
    VectorXd d(n);
    for (long i = 0; i < n; i++) d(i) = rho[i] * 0.5;
    VectorXd e1, va, vb, e2, ca, cb;
    std::shared_ptr<Functional> f;
    f->xalpha_x_functional(d, d, e1, va, vb);
    f->vwm_c_functional(d, d, e2, ca, cb);
    for (long i = 0; i < n; i++) {
        ex[i] = e1(i);
        vx[i] = va(i);
        ec[i] = e2(i);
        vc[i] = ca(i);
    }
}

// ---- full-DFT mode only -------------------------------------------------
// which: 0 S, 1 T, 2 V, 3 H, 4 X, 5 P, 6 J, 7 XC, 8 C
int ref_get_matrix(refdft* h, int which, double* out) {
    if (!h->dft) return -1;
    DFT& d = *h->dft;
    const MatrixXXd* M[] = {&d.S, &d.T, &d.V, &d.H, &d.X, &d.P, &d.J, &d.XC, &d.C};
    if (which < 0 || which > 8) return -2;
    std::memcpy(out, M[which]->data(), sizeof(double) * M[which]->size());
    return 0;
}

// out: et, exc, enuc, e_one, e_J, sum(w rho)
int ref_get_energies(refdft* h, double* out) {
    if (!h->dft) return -1;
    DFT& d = *h->dft;
    out[0] = d.et;
    out[1] = d.exc;
    out[2] = d.enuc;
    out[3] = d.single_electron_energy;
    out[4] = d.electronic_repulsion;
    out[5] = d.molgrid->calculate_density();
    return 0;
}

// one pass of the DFT::scf loop body (src/dft.cpp:100-103); returns total energy
double ref_scf_step(refdft* h) {
    if (!h->dft) return NAN;
    CoutSilencer sil(h->quiet);
    DFT& d = *h->dft;
    d.calculate_density_matrix();
    d.calculate_electronic_repulsion_matrix();
    d.calculate_exchange_correlation_matrix();
    d.calculate_energy();
    return d.et;
}

// CPU baseline: wall-clock of one iteration's grid work for a fixed P — the reference's own methods, in the
// order DFT::scf issues them (src/dft.cpp:100-102 minus the eigen-solve).  phases_ms: set_density+correct,
// hartree, xc, electron count.  Returns the total in ms.
double ref_time_iteration(refdft* h, const double* P, double* phases_ms, double* J, double* XC, double* exc) {
    CoutSilencer sil(h->quiet);
    double t0 = now_ms();
    ref_set_density(h, P);
    double t1 = now_ms();
    ref_hartree(h, J);
    double t2 = now_ms();
    ref_xc(h, XC, exc);
    double t3 = now_ms();
    volatile double ne = ref_electron_count(h);
    (void)ne;
    double t4 = now_ms();
    if (phases_ms) {
        phases_ms[0] = t1 - t0;
        phases_ms[1] = t2 - t1;
        phases_ms[2] = t3 - t2;
        phases_ms[3] = t4 - t3;
    }
    return t4 - t0;
}

}  // extern "C"
