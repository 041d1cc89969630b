#include "integrals.hpp"

#include <cmath>

namespace dftcxx {

namespace {

const int LMAX = 2;            // d functions
const int JMAX = LMAX + 2;     // the kinetic operator raises the ket by two
const int TMAX = LMAX + JMAX;  // highest Hermite index in one dimension

// Hermite expansion coefficients of the 1-D overlap distribution x_A^i x_B^j exp(-a x_A^2 - b x_B^2), with the
// Gaussian-product exponential factored out (E_0^{00} = 1).  E[i][j][t], t <= i + j.
struct Hermite1D {
    double E[LMAX + 1][JMAX + 1][TMAX + 2];
    Hermite1D(int imax, int jmax, double p, double xpa, double xpb) {
        for (auto& a : E)
            for (auto& b : a)
                for (double& c : b) c = 0.0;
        const double h = 0.5 / p;
        E[0][0][0] = 1.0;
        for (int i = 0; i <= imax; i++) {
            if (i > 0)  // raise i from (i-1, 0)
                for (int t = 0; t <= i; t++)
                    E[i][0][t] = (t > 0 ? h * E[i - 1][0][t - 1] : 0.0) + xpa * E[i - 1][0][t] + (t + 1) * E[i - 1][0][t + 1];
            for (int j = 1; j <= jmax; j++)  // raise j from (i, j-1)
                for (int t = 0; t <= i + j; t++)
                    E[i][j][t] = (t > 0 ? h * E[i][j - 1][t - 1] : 0.0) + xpb * E[i][j - 1][t] + (t + 1) * E[i][j - 1][t + 1];
        }
    }
};

struct PairGeometry {
    double p, pre;  // total exponent, exp(-a b |AB|^2 / p)
    vec3 P, PA, PB;
    PairGeometry(const GTO& a, const GTO& b) {
        const double aa = a.get_alpha(), bb = b.get_alpha();
        p = aa + bb;
        double rab2 = 0.0;
        for (int d = 0; d < 3; d++) {
            const double ab = a.get_position()[d] - b.get_position()[d];
            rab2 += ab * ab;
            P[d] = (aa * a.get_position()[d] + bb * b.get_position()[d]) / p;
            PA[d] = P[d] - a.get_position()[d];
            PB[d] = P[d] - b.get_position()[d];
        }
        pre = std::exp(-aa * bb * rab2 / p);
    }
};

const double kPi = 3.141592653589793238462643383279502884;

template <typename F>
double contract(const CGF& a, const CGF& b, F&& prim) {
    double sum = 0.0;
    for (unsigned int k = 0; k < a.size(); k++)
        for (unsigned int l = 0; l < b.size(); l++)
            sum += a.get_norm_gto(k) * b.get_norm_gto(l) * a.get_coefficient_gto(k) * b.get_coefficient_gto(l) * prim(a.get_gto(k), b.get_gto(l));
    return sum;
}

}  // namespace

// ---- Boys function -------------------------------------------------------------------------------------------
void Integrator::boys(int nmax, double x, double* F) {
    if (x < 35.0) {
        // ascending series at the top order, then stable downward recursion
        const double ex = std::exp(-x);
        double term = 1.0 / (2.0 * nmax + 1.0), sum = term;
        for (int k = 1; k < 400; k++) {
            term *= 2.0 * x / (2.0 * nmax + 2.0 * k + 1.0);
            sum += term;
            if (term < 1e-18 * sum) break;
        }
        F[nmax] = ex * sum;
        for (int n = nmax; n > 0; n--) F[n - 1] = (2.0 * x * F[n] + ex) / (2.0 * n - 1.0);
    } else {
        // large argument: F_0 from erf, upward recursion (stable for x >> n)
        const double ex = std::exp(-x);
        F[0] = 0.5 * std::sqrt(kPi / x) * std::erf(std::sqrt(x));
        for (int n = 0; n < nmax; n++) F[n + 1] = ((2.0 * n + 1.0) * F[n] - ex) / (2.0 * x);
    }
}

// ---- primitives -------------------------------------------------------------------------------------------------
double Integrator::overlap(const GTO& a, const GTO& b) const {
    const PairGeometry g(a, b);
    const Hermite1D ex((int)a.get_l(), (int)b.get_l(), g.p, g.PA[0], g.PB[0]);
    const Hermite1D ey((int)a.get_m(), (int)b.get_m(), g.p, g.PA[1], g.PB[1]);
    const Hermite1D ez((int)a.get_n(), (int)b.get_n(), g.p, g.PA[2], g.PB[2]);
    return std::pow(kPi / g.p, 1.5) * g.pre * ex.E[a.get_l()][b.get_l()][0] * ey.E[a.get_m()][b.get_m()][0] * ez.E[a.get_n()][b.get_n()][0];
}

double Integrator::kinetic(const GTO& a, const GTO& b) const {
    // -1/2 <a| nabla^2 |b> through overlaps with the ket's powers shifted by +-2 (src/integrals.cpp:107-129 uses
    // the same textbook identity)
    const PairGeometry g(a, b);
    const int la[3] = {(int)a.get_l(), (int)a.get_m(), (int)a.get_n()};
    const int lb[3] = {(int)b.get_l(), (int)b.get_m(), (int)b.get_n()};
    const Hermite1D e[3] = {Hermite1D(la[0], lb[0] + 2, g.p, g.PA[0], g.PB[0]), Hermite1D(la[1], lb[1] + 2, g.p, g.PA[1], g.PB[1]),
                            Hermite1D(la[2], lb[2] + 2, g.p, g.PA[2], g.PB[2])};
    auto s1 = [&](int d, int shift) -> double {
        const int j = lb[d] + shift;
        return j < 0 ? 0.0 : e[d].E[la[d]][j][0];
    };
    const double s[3] = {s1(0, 0), s1(1, 0), s1(2, 0)};
    const double beta = b.get_alpha();
    const double base = std::pow(kPi / g.p, 1.5) * g.pre;
    const double term0 = beta * (2.0 * (lb[0] + lb[1] + lb[2]) + 3.0) * s[0] * s[1] * s[2];
    const double term1 = -2.0 * beta * beta * (s1(0, 2) * s[1] * s[2] + s[0] * s1(1, 2) * s[2] + s[0] * s[1] * s1(2, 2));
    const double term2 = -0.5 * (lb[0] * (lb[0] - 1) * s1(0, -2) * s[1] * s[2] + lb[1] * (lb[1] - 1) * s[0] * s1(1, -2) * s[2] +
                                 lb[2] * (lb[2] - 1) * s[0] * s[1] * s1(2, -2));
    return base * (term0 + term1 + term2);
}

double Integrator::nuclear(const GTO& a, const GTO& b, const vec3& C) const {
    const double pi_ref = 3.14159265359;  // the reference's prefactor (src/integrals.cpp:343)
    const PairGeometry g(a, b);
    const int la[3] = {(int)a.get_l(), (int)a.get_m(), (int)a.get_n()};
    const int lb[3] = {(int)b.get_l(), (int)b.get_m(), (int)b.get_n()};
    const Hermite1D ex(la[0], lb[0], g.p, g.PA[0], g.PB[0]);
    const Hermite1D ey(la[1], lb[1], g.p, g.PA[1], g.PB[1]);
    const Hermite1D ez(la[2], lb[2], g.p, g.PA[2], g.PB[2]);
    const int tm = la[0] + lb[0], um = la[1] + lb[1], vm = la[2] + lb[2];
    const int nmax = tm + um + vm;
    const double pc[3] = {g.P[0] - C[0], g.P[1] - C[1], g.P[2] - C[2]};
    const double x = std::max(std::fabs(g.p * (pc[0] * pc[0] + pc[1] * pc[1] + pc[2] * pc[2])), 1e-8);  // clamp: src/gamma.cpp:41-42
    const int NM = 4 * LMAX + 1;
    double F[NM];
    boys(nmax, x, F);
    // R[n][t][u][v] = Hermite Coulomb integrals, filled for decreasing auxiliary index n
    static thread_local double R[NM][2 * LMAX + 1][2 * LMAX + 1][2 * LMAX + 1];
    double m2p = 1.0;
    for (int n = 0; n <= nmax; n++) {
        R[n][0][0][0] = m2p * F[n];
        m2p *= -2.0 * g.p;
    }
    for (int n = nmax - 1; n >= 0; n--) {
        const int budget = nmax - n;
        for (int t = 0; t <= tm; t++)
            for (int u = 0; u <= um; u++)
                for (int v = 0; v <= vm; v++) {
                    if (t + u + v == 0 || t + u + v > budget) continue;
                    double val;
                    if (t > 0)
                        val = (t > 1 ? (t - 1) * R[n + 1][t - 2][u][v] : 0.0) + pc[0] * R[n + 1][t - 1][u][v];
                    else if (u > 0)
                        val = (u > 1 ? (u - 1) * R[n + 1][t][u - 2][v] : 0.0) + pc[1] * R[n + 1][t][u - 1][v];
                    else
                        val = (v > 1 ? (v - 1) * R[n + 1][t][u][v - 2] : 0.0) + pc[2] * R[n + 1][t][u][v - 1];
                    R[n][t][u][v] = val;
                }
    }
    double sum = 0.0;
    for (int t = 0; t <= tm; t++)
        for (int u = 0; u <= um; u++)
            for (int v = 0; v <= vm; v++) sum += ex.E[la[0]][lb[0]][t] * ey.E[la[1]][lb[1]][u] * ez.E[la[2]][lb[2]][v] * R[0][t][u][v];
    return -2.0 * pi_ref / g.p * g.pre * sum;
}

// ---- contracted ------------------------------------------------------------------------------------------------
double Integrator::overlap(const CGF& a, const CGF& b) const {
    return contract(a, b, [this](const GTO& x, const GTO& y) { return overlap(x, y); });
}
double Integrator::kinetic(const CGF& a, const CGF& b) const {
    return contract(a, b, [this](const GTO& x, const GTO& y) { return kinetic(x, y); });
}
double Integrator::nuclear(const CGF& a, const CGF& b, const vec3& nucleus, unsigned int charge) const {
    return contract(a, b, [&](const GTO& x, const GTO& y) { return nuclear(x, y, nucleus); }) * (double)charge;
}

}  // namespace dftcxx
