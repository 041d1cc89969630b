// Host-side (CPU, libm) construction of the small point-independent tables the grid kernels consume.
// They are built once per grid on the host on purpose: the reference evaluates exactly these quantities
// with glibc's cos/sin/pow/acos/atan2, and computing them with the same library keeps the radial nodes,
// quadrature weights, Y_lm tables and Poisson operators bit-identical to the reference's.
#pragma once
#include <vector>

namespace dfg {

extern const int kLebedevCounts[11];
extern const double kLebedevTable[][4];  // 914 rows x,y,z,w (reference src/quadrature.h:53-968, digit for digit)
int lebedev_offset(int order);           // reference src/moleculargrid.cpp:209-212

// Gauss-Chebyshev (2nd kind) radial nodes mapped to [0,inf): reference src/atomicgrid.cpp:46-63.
// r[i], w[i] for i = p-1 = 0..N-1 (r descending).
void make_radial(int N, std::vector<double>& r, std::vector<double>& w);

// Real spherical harmonics incl. prefactor at the Lebedev directions, as the reference evaluates them in
// calculate_rho_lm / calculate_U_lm (src/atomicgrid.cpp:268-284, src/spherical_harmonics.cpp:24-47,81-117).
// Y[j*nlm + lm], lm = l*l + l + m.  pre[l*(lmax+1)+|m|] = prefactor_spherical_harmonic(l,m).
void make_ylm_table(int leb_offset, int nang, int lmax, std::vector<double>& Y, std::vector<double>& pre);

// Radial Poisson operators of calculate_U_lm (src/atomicgrid.cpp:316-389,402,419,424): for each l the
// (N+2)x(N+2) finite-difference matrix minus l(l+1)/r^2 on the interior diagonal, factorised by
// partial-pivot LU (what Eigen::PartialPivLU does).  LU[l] is row-major (N+2)^2 with unit-lower L below
// the diagonal; perm[l*(N+2)+i] = source row of pivoted row i; lo/hi = first/last non-zero column per row.
struct PoissonLU {
    int n;  // N+2
    int nl; // lmax+1
    std::vector<double> lu;
    std::vector<int> perm, lo, hi;
};
void make_poisson_lu(int N, int lmax, const std::vector<double>& r, PoissonLU& out);

// x-dependent part of the not-a-knot spline system of Cspline::generate_spline (src/cspline.cpp:66-127) for
// the common abscissa x = r ascending: sub-diagonal A, swept C' and the pivots B[i]-A[i]*C'[i-1], plus
// h = x[i+1]-x[i] and 1/h.  All arrays length N (unused tail entries are zero).
struct SplineSystem {
    std::vector<double> x, A, Cp, den, h, rh;
    double b0_c0[2];      // B[0], C[0] of the first row
    double first_w[3];    // Y[0] = r0*first_w[0] + r1*first_w[1]          (3*h0*h1+2*h1*h1, h0*h0)
    double last_w[3];     // Y[N-1] = r0*last_w[0] + r1*last_w[1]           (h1*h1, 3*h0*h1+2*h0*h0)
};
void make_spline_system(int N, const std::vector<double>& r, SplineSystem& out);

// Constants of the Slater-Xalpha + VWN5 functional exactly as src/functionals.cpp:30-32,120-150 forms them.
struct LdaConstants {
    double fac;         // -2.25*(2/3)*pow(3/4/pi,1/3)
    double vfac;        // 4/3*fac
    double x_pref;      // 3/4/pi
    double a, x0, b, c; // VWN5 paramagnetic
    double q, Xx0, bx0_over_Xx0, atan_pref;
};
void make_lda_constants(LdaConstants& out);

}  // namespace dfg
