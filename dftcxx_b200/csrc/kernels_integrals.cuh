// One-electron integrals over contracted Cartesian Gaussians (s, p, d) on the device: overlap S, kinetic energy T and nuclear
// attraction V (reference: DFT::construct_matrices src/dft.cpp:185-198, src/integrals.cpp:43-387).
// Scheme: McMurchie-Davidson — Hermite expansion coefficients E_t^{ij} per Cartesian direction and Hermite Coulomb integrals
// R_tuv from the Boys function — the same closed forms the C++ host evaluates (dftcxx_b200/host/integrals.cpp, checked
// against the reference's Taketa-Huzinaga-O-ohata sums), with the reference's two numerical conventions kept: the
// nuclear-attraction prefactor uses pi = 3.14159265359 (src/integrals.cpp:343) and the Boys argument is clamped from
// below at 1e-8 (src/gamma.cpp:40-45).  sm_100a only.
#pragma once
#include "common.cuh"

namespace dfg {

constexpr int kIntL = 2;               // d functions
constexpr int kIntJ = kIntL + 2;       // the kinetic operator raises the ket by two
constexpr int kIntT = kIntL + kIntJ;   // highest Hermite index in one dimension
constexpr int kIntWarps = 4;           // (i, j) pairs per CTA, one warp each
constexpr double kIntPi = 3.141592653589793238462643383279502884;

// E[i][j][t], t <= i + j, of x_A^i x_B^j exp(-a x_A^2 - b x_B^2) with the Gaussian-product factor taken out (E_0^{00} = 1)
struct IntHermite {
    double E[kIntL + 1][kIntJ + 1][kIntT + 2];
    __device__ void build(int imax, int jmax, double p, double xpa, double xpb) {
#pragma unroll 1
        for (int i = 0; i <= kIntL; i++)
#pragma unroll 1
            for (int j = 0; j <= kIntJ; j++)
#pragma unroll 1
                for (int t = 0; t < kIntT + 2; t++) E[i][j][t] = 0.0;
        const double h = 0.5 / p;
        E[0][0][0] = 1.0;
#pragma unroll 1
        for (int i = 0; i <= imax; i++) {
            if (i > 0)
#pragma unroll 1
                for (int t = 0; t <= i; t++)
                    E[i][0][t] = (t > 0 ? h * E[i - 1][0][t - 1] : 0.0) + xpa * E[i - 1][0][t] + (t + 1) * E[i - 1][0][t + 1];
#pragma unroll 1
            for (int j = 1; j <= jmax; j++)
#pragma unroll 1
                for (int t = 0; t <= i + j; t++)
                    E[i][j][t] = (t > 0 ? h * E[i][j - 1][t - 1] : 0.0) + xpb * E[i][j - 1][t] + (t + 1) * E[i][j - 1][t + 1];
        }
    }
};

// F_n(x), n = 0..nmax: ascending series at the top order + downward recursion for x < 35, erf + upward recursion beyond
__device__ inline void int_boys(int nmax, double x, double* F) {
    const double ex = exp(-x);
    if (x < 35.0) {
        double term = 1.0 / (2.0 * nmax + 1.0), sum = term;
#pragma unroll 1
        for (int k = 1; k < 400; k++) {
            term *= 2.0 * x / (2.0 * nmax + 2.0 * k + 1.0);
            sum += term;
            if (term < 1e-18 * sum) break;
        }
        F[nmax] = ex * sum;
#pragma unroll 1
        for (int n = nmax; n > 0; n--) F[n - 1] = (2.0 * x * F[n] + ex) / (2.0 * n - 1.0);
    } else {
        F[0] = 0.5 * sqrt(kIntPi / x) * erf(sqrt(x));
#pragma unroll 1
        for (int n = 0; n < nmax; n++) F[n + 1] = ((2.0 * n + 1.0) * F[n] - ex) / (2.0 * x);
    }
}

// One warp per CGF pair (i <= j) of the upper triangle.  Every lane walks the primitive pairs (Hermite tables per pair, the
// same in all lanes); S and T come out of the tables directly, the nuclear attraction is split over the lanes by nucleus
// (lane takes nuclei lane, lane + 32, ...) and summed by a fixed-order lane tree, so results are bit-identical run to run.
// Zq[k] = nuclear charge as a double.  Outputs nb x nb, both triangles written.
__global__ void __launch_bounds__(kIntWarps * 32)
k_one_electron(int nbf, int natoms, const int* __restrict__ bf_center, const int* __restrict__ bf_prim_off, const double* __restrict__ center_xyz,
               const int* __restrict__ prim_exp, const double* __restrict__ exp_alpha, const double* __restrict__ prim_coeff,
               const double* __restrict__ prim_norm, const int* __restrict__ prim_lmn, const double* __restrict__ atom_xyz,
               const double* __restrict__ Zq, double* __restrict__ S, double* __restrict__ T, double* __restrict__ V) {
    const int lane = threadIdx.x & 31;
    const long pair = (long)blockIdx.x * kIntWarps + (threadIdx.x >> 5);
    const long npair = (long)nbf * (nbf + 1) / 2;
    if (pair >= npair) return;
    // pair -> (i, j), i <= j, row-major over the upper triangle
    int i = (int)(((2.0 * nbf + 1.0) - sqrt((2.0 * nbf + 1.0) * (2.0 * nbf + 1.0) - 8.0 * (double)pair)) * 0.5);
    while ((long)i * (2 * nbf - i + 1) / 2 > pair) i--;
    while ((long)(i + 1) * (2 * nbf - i) / 2 <= pair) i++;
    const int j = i + (int)(pair - (long)i * (2 * nbf - i + 1) / 2);
    const int ca = bf_center[i], cb = bf_center[j];
    const double A[3] = {center_xyz[3 * ca], center_xyz[3 * ca + 1], center_xyz[3 * ca + 2]};
    const double B[3] = {center_xyz[3 * cb], center_xyz[3 * cb + 1], center_xyz[3 * cb + 2]};
    double rab2 = 0.0;
#pragma unroll 1
    for (int d = 0; d < 3; d++) rab2 += (A[d] - B[d]) * (A[d] - B[d]);
    double s_sum = 0.0, t_sum = 0.0, v_sum = 0.0;
    IntHermite e[3];
#pragma unroll 1
    for (int ka = bf_prim_off[i]; ka < bf_prim_off[i + 1]; ka++) {
        const double aa = exp_alpha[prim_exp[ka]];
        const int lmna = prim_lmn[ka];
        const int la[3] = {lmna & 15, (lmna >> 4) & 15, (lmna >> 8) & 15};
#pragma unroll 1
        for (int kb = bf_prim_off[j]; kb < bf_prim_off[j + 1]; kb++) {
            const double bb = exp_alpha[prim_exp[kb]];
            const int lmnb = prim_lmn[kb];
            const int lb[3] = {lmnb & 15, (lmnb >> 4) & 15, (lmnb >> 8) & 15};
            const double cc = prim_norm[ka] * prim_norm[kb] * prim_coeff[ka] * prim_coeff[kb];
            const double p = aa + bb;
            double P[3];
#pragma unroll 1
            for (int d = 0; d < 3; d++) {
                P[d] = (aa * A[d] + bb * B[d]) / p;
                e[d].build(la[d], lb[d] + 2, p, P[d] - A[d], P[d] - B[d]);
            }
            const double pre = exp(-aa * bb * rab2 / p);
            const double base = pow(kIntPi / p, 1.5) * pre;
            // overlap and kinetic energy (ket powers shifted by +-2)
            const double s0[3] = {e[0].E[la[0]][lb[0]][0], e[1].E[la[1]][lb[1]][0], e[2].E[la[2]][lb[2]][0]};
            double sp[3], sm[3];
#pragma unroll 1
            for (int d = 0; d < 3; d++) {
                sp[d] = e[d].E[la[d]][lb[d] + 2][0];
                sm[d] = lb[d] >= 2 ? e[d].E[la[d]][lb[d] - 2][0] : 0.0;
            }
            s_sum += cc * (base * s0[0] * s0[1] * s0[2]);
            const double term0 = bb * (2.0 * (lb[0] + lb[1] + lb[2]) + 3.0) * s0[0] * s0[1] * s0[2];
            const double term1 = -2.0 * bb * bb * (sp[0] * s0[1] * s0[2] + s0[0] * sp[1] * s0[2] + s0[0] * s0[1] * sp[2]);
            const double term2 = -0.5 * (lb[0] * (lb[0] - 1) * sm[0] * s0[1] * s0[2] + lb[1] * (lb[1] - 1) * s0[0] * sm[1] * s0[2] +
                                         lb[2] * (lb[2] - 1) * s0[0] * s0[1] * sm[2]);
            t_sum += cc * (base * (term0 + term1 + term2));
            // nuclear attraction: this lane's nuclei
            const int tm = la[0] + lb[0], um = la[1] + lb[1], vm = la[2] + lb[2];
            const int nmax = tm + um + vm;
            const double vpre = -2.0 * 3.14159265359 / p * pre;
#pragma unroll 1
            for (int k = lane; k < natoms; k += 32) {
                const double pc[3] = {P[0] - atom_xyz[3 * k], P[1] - atom_xyz[3 * k + 1], P[2] - atom_xyz[3 * k + 2]};
                const double x = fmax(fabs(p * (pc[0] * pc[0] + pc[1] * pc[1] + pc[2] * pc[2])), 1e-8);
                double F[2 * kIntL * 2 + 1];
                int_boys(nmax, x, F);
                // Hermite Coulomb integrals R^n_tuv, layer n from layer n + 1 (two layers alive)
                double R[2][2 * kIntL + 1][2 * kIntL + 1][2 * kIntL + 1];
                double m2p[2 * kIntL * 2 + 1];
                m2p[0] = 1.0;
#pragma unroll 1
                for (int n = 1; n <= nmax; n++) m2p[n] = m2p[n - 1] * (-2.0 * p);
#pragma unroll 1
                for (int n = nmax; n >= 0; n--) {
                    double(*cur)[2 * kIntL + 1][2 * kIntL + 1] = R[n & 1];
                    double(*nxt)[2 * kIntL + 1][2 * kIntL + 1] = R[(n + 1) & 1];
                    const int budget = nmax - n;
                    cur[0][0][0] = m2p[n] * F[n];
#pragma unroll 1
                    for (int t = 0; t <= tm; t++)
#pragma unroll 1
                        for (int u = 0; u <= um; u++)
#pragma unroll 1
                            for (int v = 0; v <= vm; v++) {
                                if (t + u + v == 0 || t + u + v > budget) continue;
                                double val;
                                if (t > 0)
                                    val = (t > 1 ? (t - 1) * nxt[t - 2][u][v] : 0.0) + pc[0] * nxt[t - 1][u][v];
                                else if (u > 0)
                                    val = (u > 1 ? (u - 1) * nxt[t][u - 2][v] : 0.0) + pc[1] * nxt[t][u - 1][v];
                                else
                                    val = (v > 1 ? (v - 1) * nxt[t][u][v - 2] : 0.0) + pc[2] * nxt[t][u][v - 1];
                                cur[t][u][v] = val;
                            }
                }
                double sum = 0.0;
#pragma unroll 1
                for (int t = 0; t <= tm; t++)
#pragma unroll 1
                    for (int u = 0; u <= um; u++)
#pragma unroll 1
                        for (int v = 0; v <= vm; v++) sum += e[0].E[la[0]][lb[0]][t] * e[1].E[la[1]][lb[1]][u] * e[2].E[la[2]][lb[2]][v] * R[0][t][u][v];
                v_sum += cc * (vpre * sum) * Zq[k];
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v_sum += __shfl_xor_sync(0xffffffffu, v_sum, o);
    if (lane == 0) {
        S[(size_t)i * nbf + j] = S[(size_t)j * nbf + i] = s_sum;
        T[(size_t)i * nbf + j] = T[(size_t)j * nbf + i] = t_sum;
        V[(size_t)i * nbf + j] = V[(size_t)j * nbf + i] = v_sum;
    }
}

}  // namespace dfg
