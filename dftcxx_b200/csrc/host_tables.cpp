// Host tables for the grid engine (see host_tables.h).  Compiled with -ffp-contract=off so that every
// expression rounds exactly like the reference's scalar code; formulas cite the reference lines they restate.
#include "host_tables.h"

#include <cmath>
#include <cstdlib>
#include <stdexcept>

namespace dfg {

const int kLebedevCounts[11] = {6, 14, 26, 38, 50, 74, 86, 110, 146, 170, 194};
const double kLebedevTable[][4] = {
#include "lebedev_table.inc"
};

int lebedev_offset(int order) {
    int off = 0;
    for (int i = 0; i < order; i++) off += kLebedevCounts[i];
    return off;
}

static const double kPi = 3.14159265358979323846;  // src/atomicgrid.h:33

void make_radial(int N, std::vector<double>& r, std::vector<double>& w) {
    r.assign(N, 0.0);
    w.assign(N, 0.0);
    const double f = kPi / (double)(N + 1);
    for (int p = 1; p <= N; p++) {
        const double s = std::sin(f * (double)p);
        double wp = f * std::pow(s, 2.0);
        const double x = std::cos(f * (double)p);
        r[p - 1] = (1.0 + x) / (1.0 - x);
        wp = wp / std::sqrt(1.0 - std::pow(x, 2.0)) * 2.0 / std::pow(1.0 - x, 2.0);
        w[p - 1] = wp;
    }
}

// ---- real spherical harmonics ------------------------------------------------------------------------
static double factorial_d(int n) {
    double f = 1.0;
    for (int i = 2; i <= n; i++) f *= (double)i;
    return f;
}

// src/spherical_harmonics.cpp:28-33
static double sh_prefactor(int l, int m) {
    static const double pre = 1.0 / std::sqrt(4 * M_PI);
    const int am = std::abs(m);
    return pre * (m == 0 ? 1 : std::sqrt(2.0)) * std::sqrt((double)(2 * l + 1) * factorial_d(l - am) / factorial_d(l + am));
}

// associated Legendre P_n^m(x) with Condon-Shortley phase, same recurrences as src/spherical_harmonics.cpp:81-117
static double assoc_legendre(int n, int m, double x) {
    std::vector<double> v(n + 1, 0.0);
    if (m <= n) {
        v[m] = 1.0;
        double fact = 1.0;
        for (int k = 0; k < m; k++) {
            v[m] *= -fact * std::sqrt(1.0 - x * x);
            fact += 2.0;
        }
    }
    if (m + 1 <= n) v[m + 1] = x * (double)(2 * m + 1) * v[m];
    for (int j = m + 2; j <= n; j++)
        v[j] = ((double)(2 * j - 1) * x * v[j - 1] + (double)(-j - m + 1) * v[j - 2]) / (double)(j - m);
    return v[n];
}

void make_ylm_table(int leb_offset, int nang, int lmax, std::vector<double>& Y, std::vector<double>& pre) {
    const int nlm = (lmax + 1) * (lmax + 1);
    Y.assign((size_t)nang * nlm, 0.0);
    pre.assign((size_t)(lmax + 1) * (lmax + 1), 0.0);
    for (int l = 0; l <= lmax; l++)
        for (int m = 0; m <= l; m++) pre[l * (lmax + 1) + m] = sh_prefactor(l, m);
    for (int j = 0; j < nang; j++) {
        const double* v = kLebedevTable[leb_offset + j];
        const double rr = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        const double azimuth = std::atan2(v[1], v[0]);
        const double pole = std::acos(v[2] / rr);
        for (int l = 0; l <= lmax; l++)
            for (int m = -l; m <= l; m++) {
                const double polar = assoc_legendre(l, std::abs(m), std::cos(pole));
                double az = 1.0;
                if (m > 0) az = std::cos((double)m * azimuth);
                if (m < 0) az = std::sin(-(double)m * azimuth);
                Y[(size_t)j * nlm + (l * l + l + m)] = sh_prefactor(l, m) * (polar * az);
            }
    }
}

// ---- radial Poisson operators ----------------------------------------------------------------------------
// src/atomicgrid.cpp:502-510 with m = 1
static double d2zdr2(double r, double m) {
    const double nom = m * m * (m + 3.0 * r);
    const double denom = 2.0 * M_PI * std::pow((m * r) / ((m + r) * (m + r)), 1.5) * std::pow(m + r, 5.0);
    return nom / denom;
}
static double dzdrsq(double r, double m) { return m / (M_PI * M_PI * r * (m + r) * (m + r)); }

namespace {
struct Stencil {
    int ncol;
    double d1, d2;  // c1 /= d1*h*h ; c2 /= d2*h
    double k1[7], k2[7];
};
}  // namespace

static void fd_matrix(int N, const std::vector<double>& r, std::vector<double>& A) {
    const int n = N + 2;
    const double h = 1.0 / (double)(N + 1);
    A.assign((size_t)n * n, 0.0);
    // one-sided 5/6-point rows next to both boundaries and the centred 7-point interior row
    // (src/atomicgrid.cpp:334-388); coefficient tables multiply c1 (2nd derivative) and c2 (1st derivative).
    static const Stencil row1 = {5, 12.0, 12.0, {11, -20, 6, 4, -1}, {-3, -10, 18, -6, 1}};
    static const Stencil row2 = {6, 12.0, 60.0, {-1, 16, -30, 16, -1, 0}, {3, -30, -20, 60, -15, 2}};
    static const Stencil rowNm1 = {6, 12.0, 60.0, {0, -1, 16, -30, 16, -1}, {-2, 15, -60, 20, 30, -3}};  // cols N-4..N+1
    static const Stencil rowN = {5, 12.0, 12.0, {-1, 4, 6, -20, 11}, {-1, 6, -18, 10, 3}};                 // cols N-3..N+1
    static const Stencil mid = {7, 180.0, 60.0, {2, -27, 270, -490, 270, -27, 2}, {-1, 9, -45, 0, 45, -9, 1}};
    for (int i = 0; i < n; i++) {
        if (i == 0 || i == N + 1) {  // Dirichlet rows U(r=inf), U(r=0) (src/atomicgrid.cpp:330-333,376-379); N >= 6 keeps all row kinds distinct
            A[(size_t)i * n + i] = 1.0;
            continue;
        }
        double c1 = dzdrsq(r[i - 1], 1.0), c2 = d2zdr2(r[i - 1], 1.0);
        const Stencil* s;
        int col0;
        if (i == 1) {
            s = &row1;
            col0 = 0;
        } else if (i == 2) {
            s = &row2;
            col0 = 0;
        } else if (i == N - 1) {
            s = &rowNm1;
            col0 = N - 4;
        } else if (i == N) {
            s = &rowN;
            col0 = N - 3;
        } else {
            s = &mid;
            col0 = i - 3;
        }
        c1 /= s->d1 * h * h;
        c2 /= s->d2 * h;
        for (int k = 0; k < s->ncol; k++) {
            const int col = col0 + k;
            if (col < 0 || col >= n) throw std::runtime_error("radial_points too small for the 7-point Poisson stencil");
            A[(size_t)i * n + col] = s->k1[k] * c1 + s->k2[k] * c2;
        }
    }
}

void make_poisson_lu(int N, int lmax, const std::vector<double>& r, PoissonLU& out) {
    if (N < 6) throw std::runtime_error("radial_points must be at least 6");
    const int n = N + 2;
    out.n = n;
    out.nl = lmax + 1;
    out.lu.assign((size_t)out.nl * n * n, 0.0);
    out.perm.assign((size_t)out.nl * n, 0);
    out.lo.assign((size_t)out.nl * n, 0);
    out.hi.assign((size_t)out.nl * n, 0);
    std::vector<double> A;
    fd_matrix(N, r, A);
    for (int l = 0; l <= lmax; l++) {
        double* M = &out.lu[(size_t)l * n * n];
        int* perm = &out.perm[(size_t)l * n];
        for (size_t t = 0; t < (size_t)n * n; t++) M[t] = A[t];
        for (int i = 1; i < N + 1; i++) M[(size_t)i * n + i] -= (double)l * (double)(l + 1) / (r[i - 1] * r[i - 1]);
        for (int i = 0; i < n; i++) perm[i] = i;
        // right-looking LU with row partial pivoting
        for (int k = 0; k < n; k++) {
            int p = k;
            double best = std::fabs(M[(size_t)k * n + k]);
            for (int i = k + 1; i < n; i++) {
                const double v = std::fabs(M[(size_t)i * n + k]);
                if (v > best) {
                    best = v;
                    p = i;
                }
            }
            if (p != k) {
                for (int j = 0; j < n; j++) {
                    const double t = M[(size_t)k * n + j];
                    M[(size_t)k * n + j] = M[(size_t)p * n + j];
                    M[(size_t)p * n + j] = t;
                }
                const int t = perm[k];
                perm[k] = perm[p];
                perm[p] = t;
            }
            const double piv = M[(size_t)k * n + k];
            if (piv == 0.0) throw std::runtime_error("singular radial Poisson operator");
            for (int i = k + 1; i < n; i++) {
                double& lik = M[(size_t)i * n + k];
                if (lik == 0.0) continue;
                lik /= piv;
                for (int j = k + 1; j < n; j++) {
                    const double ukj = M[(size_t)k * n + j];
                    if (ukj != 0.0) M[(size_t)i * n + j] -= lik * ukj;
                }
            }
        }
        for (int i = 0; i < n; i++) {
            int lo = i, hi = i;
            for (int j = 0; j < i; j++)
                if (M[(size_t)i * n + j] != 0.0) {
                    lo = j;
                    break;
                }
            for (int j = n - 1; j > i; j--)
                if (M[(size_t)i * n + j] != 0.0) {
                    hi = j;
                    break;
                }
            out.lo[(size_t)l * n + i] = lo;
            out.hi[(size_t)l * n + i] = hi;
        }
    }
}

// ---- spline system -------------------------------------------------------------------------------------
void make_spline_system(int N, const std::vector<double>& r, SplineSystem& s) {
    if (N < 3) throw std::runtime_error("Cspline data range has to contain 3 of more items");
    s.x.assign(N, 0.0);
    for (int i = 0; i < N; i++) s.x[i] = r[N - 1 - i];  // ascending r (src/atomicgrid.cpp:540-543)
    for (int i = 1; i < N; i++)
        if (s.x[i] <= s.x[i - 1]) throw std::runtime_error("Spline x-data should be continuously increasing");
    s.A.assign(N, 0.0);
    s.Cp.assign(N, 0.0);
    s.den.assign(N, 0.0);
    s.h.assign(N, 0.0);
    s.rh.assign(N, 0.0);
    std::vector<double> B(N, 0.0), C(N, 0.0);
    for (int i = 0; i + 1 < N; i++) {
        s.h[i] = s.x[i + 1] - s.x[i];
        s.rh[i] = 1.0 / (s.x[i + 1] - s.x[i]);
    }
    double h0 = s.x[1] - s.x[0], h1 = s.x[2] - s.x[1];
    B[0] = h1 * (h0 + h1);
    C[0] = (h0 + h1) * (h0 + h1);
    s.first_w[0] = 3 * h0 * h1 + 2 * h1 * h1;
    s.first_w[1] = h0;
    s.first_w[2] = 0.0;
    for (int i = 1; i < N - 1; i++) {
        h0 = s.x[i] - s.x[i - 1];
        h1 = s.x[i + 1] - s.x[i];
        s.A[i] = h1;
        B[i] = 2 * (h0 + h1);
        C[i] = h0;
    }
    s.A[N - 1] = (h0 + h1) * (h0 + h1);
    B[N - 1] = h0 * (h0 + h1);
    s.last_w[0] = h1;
    s.last_w[1] = 3 * h0 * h1 + 2 * h0 * h0;
    s.last_w[2] = 0.0;
    s.b0_c0[0] = B[0];
    s.b0_c0[1] = C[0];
    s.Cp[0] = C[0] / B[0];
    s.den[0] = B[0];
    for (int i = 1; i < N - 1; i++) {
        s.den[i] = B[i] - s.A[i] * s.Cp[i - 1];
        s.Cp[i] = C[i] / s.den[i];
    }
    s.den[N - 1] = B[N - 1] - s.A[N - 1] * s.Cp[N - 2];
}

// ---- LDA constants ---------------------------------------------------------------------------------------
void make_lda_constants(LdaConstants& k) {
    const double pi = 3.14159265358979323846;  // src/functionals.h:49
    const double xalpha = 2.0 / 3.0;
    k.fac = -2.25 * xalpha * std::pow(3.0 / 4.0 / pi, 1.0 / 3.0);
    k.vfac = 4.0 / 3.0 * k.fac;
    k.x_pref = 3.0 / 4.0 / pi;
    k.a = 0.0310907;
    k.x0 = -0.10498;
    k.b = 3.72744;
    k.c = 12.9352;
    k.q = std::sqrt(4.0 * k.c - k.b * k.b);
    k.Xx0 = k.x0 * k.x0 + k.b * k.x0 + k.c;
    k.bx0_over_Xx0 = k.b * (k.x0 / k.Xx0);
    k.atan_pref = (2.0 * k.b / k.q) * (1.0 - (k.x0 * (2.0 * k.x0 + k.b) / k.Xx0));
}

}  // namespace dfg
