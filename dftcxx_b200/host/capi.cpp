// C entry points of the host program for tests (ctypes): run the SCF, or just the host-only pieces (integrals,
// eigen-solver) that can be checked on a machine without a GPU.
#include <cstring>
#include <string>

#include "dft.hpp"

namespace {
thread_local std::string g_err;
std::vector<dftcxx::ScfRecord> g_last;  // the SCF table of the most recent dfthost_scf* run
}

extern "C" {

const char* dfthost_last_error() { return g_err.c_str(); }

// energies: [max_iter][6] = et, exc, e_one, e_j, nelec_grid, ms ; returns the number of iterations run, < 0 on error.
// fixed_iterations > 0 runs exactly that many loop bodies (comparison at equal iteration index), otherwise the
// reference's stopping rule (|dE| <= 1e-4 and >= 3 iterations) applies.
// ngpus: devices driven by the one process; scf_mode: 0 device-resident algebra, 1 host eigen-solver + fused Fock call,
// 2 host eigen-solver + the reference's four grid calls (J and XC separately); -1 for either = the input file's keys.
// Pout (optional, nb x nb): the final density matrix.
int dfthost_scf2(const char* infile, int device, int ngpus, int scf_mode, int fixed_iterations, int max_iter, double* energies, double* enuc,
                 double* Pout) {
    try {
        dftcxx::DFT dft(infile, device, false, ngpus, scf_mode);
        if (fixed_iterations > 0)
            for (int i = 0; i < fixed_iterations && i < max_iter; i++) dft.scf_step();
        else
            dft.scf((unsigned)max_iter);
        const auto& h = dft.history();
        g_last = h;
        int n = 0;
        for (const auto& r : h) {
            if (n >= max_iter) break;
            double* e = energies + 6 * n++;
            e[0] = r.et;
            e[1] = r.exc;
            e[2] = r.e_one;
            e[3] = r.e_j;
            e[4] = r.nelec_grid;
            e[5] = r.ms;
        }
        if (enuc) *enuc = dft.nuclear_repulsion();
        if (Pout) {
            const dftcxx::Mat P = dft.density_matrix();
            std::memcpy(Pout, P.data(), sizeof(double) * P.rows() * P.cols());
        }
        return n;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

int dfthost_scf(const char* infile, int device, int fixed_iterations, int max_iter, double* energies, double* enuc) {
    return dfthost_scf2(infile, device, -1, -1, fixed_iterations, max_iter, energies, enuc, nullptr);
}

// per iteration of the most recent run: wall ms, device ms of the SCF algebra, device ms of the grid path, purification steps
int dfthost_last_timings(double* out, int max_iter) {
    int n = 0;
    for (const auto& r : g_last) {
        if (n >= max_iter) break;
        double* t = out + 4 * n++;
        t[0] = r.ms;
        t[1] = r.ms_algebra;
        t[2] = r.ms_grid;
        t[3] = r.purification_steps;
    }
    return n;
}

// host-only: S, T, V (nb x nb each) for an input file; no GPU needed.  Returns nb or < 0.
int dfthost_one_electron(const char* infile, int nb_cap, double* S, double* T, double* V) {
    try {
        auto st = std::make_shared<dftcxx::Settings>(infile);
        auto mol = std::make_shared<dftcxx::Molecule>(infile, st, false);
        const int n = (int)mol->get_nr_bfs();
        if (n > nb_cap) return n;
        dftcxx::Integrator integ;
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) {
                S[i * n + j] = integ.overlap(mol->get_cgf(i), mol->get_cgf(j));
                T[i * n + j] = integ.kinetic(mol->get_cgf(i), mol->get_cgf(j));
                double v = 0.0;
                for (unsigned k = 0; k < mol->get_nr_atoms(); k++)
                    v += integ.nuclear(mol->get_cgf(i), mol->get_cgf(j), mol->get_atomic_position(k), mol->get_atomic_charge(k));
                V[i * n + j] = v;
            }
        return n;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

// host-only: symmetric eigen-solver; A row-major n x n, w ascending, eigenvectors in the columns of Vout (row-major)
int dfthost_sym_eigen(int n, const double* A, double* w, double* Vout) {
    try {
        dftcxx::Mat M(n, n);
        std::memcpy(M.data(), A, sizeof(double) * n * n);
        std::vector<double> ev;
        dftcxx::Mat V;
        dftcxx::sym_eigen(M, ev, V);
        std::memcpy(w, ev.data(), sizeof(double) * n);
        std::memcpy(Vout, V.data(), sizeof(double) * n * n);
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

// host-only: the engine keys of an input file (Settings): out = { gpus, scf mode, density_dump present, density_dump_size,
// density_dump_points }.  Returns 0 or < 0.
int dfthost_settings(const char* infile, double* out5) {
    try {
        dftcxx::Settings st(infile);
        out5[0] = st.get_gpus();
        out5[1] = st.get_scf_mode();
        out5[2] = st.has("density_dump") ? 1.0 : 0.0;
        out5[3] = st.get_density_dump_size();
        out5[4] = st.get_density_dump_points();
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

void dfthost_boys(int nmax, double x, double* F) { dftcxx::Integrator::boys(nmax, x, F); }
}
