// One-electron integrals over contracted Cartesian Gaussians (s, p, d): overlap, kinetic energy and nuclear
// attraction: the host path (`integrals = host`; the default takes them from the device, csrc/kernels_integrals.cuh, which
// evaluates the same closed forms) and the CPU-testable statement of the scheme.  The evaluation scheme
// here is McMurchie-Davidson (Hermite expansion coefficients E_t^{ij} and Hermite Coulomb integrals R_tuv), not the
// reference's Taketa-Huzinaga-O-ohata sums; both are exact closed forms in the Boys function, so the numbers agree
// to rounding provided the reference's two numerical conventions are kept:
//   * the nuclear-attraction prefactor uses pi = 3.14159265359 (src/integrals.cpp:343) and
//   * the Boys-function argument is clamped from below at 1e-8 (src/gamma.cpp:40-45).
#pragma once
#include "molecule.hpp"

namespace dftcxx {

class Integrator {
public:
    double overlap(const CGF& a, const CGF& b) const;
    double kinetic(const CGF& a, const CGF& b) const;
    double nuclear(const CGF& a, const CGF& b, const vec3& nucleus, unsigned int charge) const;

    double overlap(const GTO& a, const GTO& b) const;
    double kinetic(const GTO& a, const GTO& b) const;
    double nuclear(const GTO& a, const GTO& b, const vec3& nucleus) const;

    // F_n(x) = int_0^1 t^{2n} exp(-x t^2) dt for n = 0..nmax
    static void boys(int nmax, double x, double* F);
};

}  // namespace dftcxx
