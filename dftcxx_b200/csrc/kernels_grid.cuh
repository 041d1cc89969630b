// Grid construction kernels: points + raw quadrature weights, Becke fuzzy-cell weights, CGF amplitudes.
// One-time work per molecule (reference: MolecularGrid::create_grid, src/moleculargrid.cpp:193-261).
#pragma once
#include "common.cuh"

namespace dfg {

// ---------------------------------------------------------------------------------------------------------
// Points and raw weights.  Restates AtomicGrid::create_atomic_grid (src/atomicgrid.cpp:50-87): the radial
// nodes r_tab / weights wrad_tab come from the host (libm, bit-identical to the reference); here
// pos = R_atom + leb*r and w = w_rad*w_leb*(|leb*r|^2*4*pi), every product and sum rounded separately
// (__dmul_rn/__dadd_rn are never contracted into FMAs) so positions and weights match the CPU bit for bit.
__global__ void k_points(GridShape g, const double* __restrict__ atom_xyz, const double* __restrict__ r_tab,
                         const double* __restrict__ wrad_tab, const double* __restrict__ leb /*[nang][4]*/,
                         double* __restrict__ px, double* __restrict__ py, double* __restrict__ pz,
                         double* __restrict__ w) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= g.nloc) return;
    const long gs = g.shell0 + t / g.nang;
    const int a = (int)(t % g.nang);
    const int atom = (int)(gs / g.nrad), i = (int)(gs % g.nrad);
    const double r = r_tab[i];
    const double qx = __dmul_rn(leb[4 * a + 0], r), qy = __dmul_rn(leb[4 * a + 1], r), qz = __dmul_rn(leb[4 * a + 2], r);
    px[t] = __dadd_rn(atom_xyz[3 * atom + 0], qx);
    py[t] = __dadd_rn(atom_xyz[3 * atom + 1], qy);
    pz[t] = __dadd_rn(atom_xyz[3 * atom + 2], qz);
    const double sq = __dadd_rn(__dadd_rn(__dmul_rn(qx, qx), __dmul_rn(qy, qy)), __dmul_rn(qz, qz));
    const double jac = __dmul_rn(__dmul_rn(sq, 4.0), 3.14159265358979323846);
    w[t] = __dmul_rn(__dmul_rn(wrad_tab[i], leb[4 * a + 3]), jac);
}

// ---------------------------------------------------------------------------------------------------------
// Becke fuzzy-cell weights (src/moleculargrid.cpp:228-254, 275-329).
//   P_k(p) = prod_{j != k, j ascending} 0.5*(1 - f3(mu_kj)),  mu_kj = (|p-R_k| - |p-R_j|) / |R_j-R_k|,
//   f(mu) = 1.5*mu - 0.5*pow(mu,3) applied three times;  wb = P_own / sum_k P_k (k ascending).
// One warp per point, lane = atom k (strided by 32): every lane runs the j-product in the reference's order,
// so P_k is reproduced operation for operation; distances sit in shared memory and are broadcast.  The
// cancellation-prone cell function is evaluated with separately rounded operations; mu^3 uses an error-free
// product so it is correctly rounded like glibc's pow(mu, 3.0) (which is what the reference calls).
__device__ __forceinline__ double cube_rn(double m) {
    const double p = __dmul_rn(m, m);
    const double e = __fma_rn(m, m, -p);
    const double c = __dmul_rn(p, m);
    const double ce = __fma_rn(p, m, -c);
    return __dadd_rn(c, __fma_rn(e, m, ce));
}
__device__ __forceinline__ double becke_cutoff(double mu) {
#pragma unroll
    for (int it = 0; it < 3; it++) mu = __dsub_rn(__dmul_rn(1.5, mu), __dmul_rn(0.5, cube_rn(mu)));
    return __dmul_rn(0.5, __dsub_rn(1.0, mu));
}

constexpr int kBeckeWarps = 4;

__global__ void __launch_bounds__(kBeckeWarps * 32)
k_becke(GridShape g, const double* __restrict__ atom_xyz, const double* __restrict__ Rinv_unused,
        const double* __restrict__ Rdist /*[natoms][natoms] |R_j-R_k|*/, const double* __restrict__ px,
        const double* __restrict__ py, const double* __restrict__ pz, double* __restrict__ w, double* __restrict__ wb) {
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int na = g.natoms;
    double* d = sm + (size_t)warp * 2 * na;  // distances
    double* P = d + na;                      // cell products
    const long t = (long)blockIdx.x * kBeckeWarps + warp;
    if (t >= g.nloc) return;
    const double x = px[t], y = py[t], z = pz[t];
    for (int k = lane; k < na; k += 32) {
        const double dx = __dsub_rn(x, atom_xyz[3 * k]), dy = __dsub_rn(y, atom_xyz[3 * k + 1]), dz = __dsub_rn(z, atom_xyz[3 * k + 2]);
        d[k] = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
    }
    __syncwarp();
    for (int k = lane; k < na; k += 32) {
        const double dk = d[k];
        double prod = 1.0;
        for (int j = 0; j < na; j++) {
            if (j == k) continue;
            const double mu = __ddiv_rn(__dsub_rn(dk, d[j]), Rdist[(size_t)j * na + k]);
            prod = __dmul_rn(prod, becke_cutoff(mu));
        }
        P[k] = prod;
    }
    __syncwarp();
    if (lane == 0) {
        const int own = (int)((g.shell0 + t / g.nang) / g.nrad);
        double denom = 0.0;
        for (int k = 0; k < na; k++) denom = __dadd_rn(denom, P[k]);
        const double v = __ddiv_rn(P[own], denom);
        wb[t] = v;
        w[t] = __dmul_rn(w[t], v);
    }
}

// ---------------------------------------------------------------------------------------------------------
// CGF amplitudes Phi[p][b] (src/gridpoint.cpp:45-52, src/cgf.cpp:146-154, 49-57):
//   phi_b(p) = sum_k c_k * ( N_k * dx^l * dy^m * dz^n * exp(-alpha_k r^2) ), products taken left to right.
// Thread = point; columns are produced in order, atom after atom; each distinct exponent of an atom is
// exponentiated once per point (the reference recomputes it per primitive: px/py/pz and sp shells share them).
// exp() is skipped where alpha*r^2 > 746 (the result is exactly +0 in FP64 there, as on the CPU).
// A [kPhiPts x kPhiCols] tile is staged in shared memory so that every Phi row is written with full 256-byte
// coalesced segments; pad columns [nbf, nbp) are written as zeros.
constexpr int kPhiPts = 128;
constexpr int kPhiCols = 32;    // one 256-byte row segment per point and pass
constexpr int kPhiMaxExp = 24;  // distinct exponents on one centre (STO-6G third row needs 18)

struct PhiPrim {  // one primitive term, 32 bytes = two 128-bit loads
    double coeff, norm;
    int exp_idx;  // absolute index into exp_alpha
    int lmn;      // l | m<<4 | n<<8
    int pad[2];
};

struct PhiBasis {
    int nbf, nbp;
    const int* bf_atom;       // [nbf] centre of column b
    const int* bf_prim_off;   // [nbf+1]
    const int* atom_exp_off;  // [ncentres+1]
    const double* exp_alpha;  // distinct exponents, centre after centre
    const PhiPrim* prims;     // [nprim]
    const double* atom_xyz;   // [ncentres][3]
};

__device__ __forceinline__ double ipow_rn(double acc, double x, int n) {
    // acc * pow(x, n) for n in {0,1,2}: pow(x,1) = x and pow(x,2) = x*x exactly rounded; factor 1.0 is exact
    if (n == 1) return __dmul_rn(acc, x);
    if (n == 2) return __dmul_rn(acc, __dmul_rn(x, x));
    return acc;
}

__global__ void __launch_bounds__(kPhiPts, 3)
k_phi(long nloc, PhiBasis B, const double* __restrict__ px, const double* __restrict__ py,
      const double* __restrict__ pz, double* __restrict__ phi) {
    extern __shared__ double sm[];
    double* tile = sm;                                          // [kPhiPts][kPhiCols+1]
    double* ex = sm + (size_t)kPhiPts * (kPhiCols + 1);          // [kPhiMaxExp][kPhiPts]
    const int tid = threadIdx.x;
    const long p0 = (long)blockIdx.x * kPhiPts;
    const long p = p0 + tid;
    const bool live = p < nloc;
    const double x = live ? px[p] : 0.0, y = live ? py[p] : 0.0, z = live ? pz[p] : 0.0;
    int cur_atom = -1, e0 = 0;
    double dx = 0, dy = 0, dz = 0;
    for (int c0 = 0; c0 < B.nbp; c0 += kPhiCols) {
        const int c1 = min(c0 + kPhiCols, B.nbp);
        for (int b = c0; b < c1; b++) {
            double val = 0.0;
            if (b < B.nbf) {
                const int atom = __ldg(B.bf_atom + b);
                if (atom != cur_atom) {
                    cur_atom = atom;
                    dx = __dsub_rn(x, __ldg(B.atom_xyz + 3 * atom));
                    dy = __dsub_rn(y, __ldg(B.atom_xyz + 3 * atom + 1));
                    dz = __dsub_rn(z, __ldg(B.atom_xyz + 3 * atom + 2));
                    const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                    e0 = __ldg(B.atom_exp_off + atom);
                    const int e1 = __ldg(B.atom_exp_off + atom + 1);
                    for (int u = e0; u < e1; u++) {
                        const double arg = __dmul_rn(__ldg(B.exp_alpha + u), r2);
                        ex[(size_t)(u - e0) * kPhiPts + tid] = arg > 746.0 ? 0.0 : exp(-arg);
                    }
                }
                const int k0 = __ldg(B.bf_prim_off + b), k1 = __ldg(B.bf_prim_off + b + 1);
                for (int k = k0; k < k1; k++) {
                    const double2 cn = __ldg(reinterpret_cast<const double2*>(B.prims + k));
                    const int2 il = __ldg(reinterpret_cast<const int2*>(B.prims + k) + 2);
                    double a = cn.y;
                    a = ipow_rn(a, dx, il.y & 15);
                    a = ipow_rn(a, dy, (il.y >> 4) & 15);
                    a = ipow_rn(a, dz, (il.y >> 8) & 15);
                    a = __dmul_rn(a, ex[(size_t)(il.x - e0) * kPhiPts + tid]);
                    val = __dadd_rn(val, __dmul_rn(cn.x, a));
                }
            }
            tile[(size_t)tid * (kPhiCols + 1) + (b - c0)] = val;
        }
        __syncthreads();
        const int ncol = c1 - c0;
        for (int row = tid >> 5; row < kPhiPts; row += kPhiPts / 32) {
            const long pr = p0 + row;
            if (pr >= nloc) break;
            const int c = tid & 31;
            if (c < ncol) phi[pr * B.nbp + c0 + c] = tile[(size_t)row * (kPhiCols + 1) + c];
        }
        __syncthreads();
    }
}

}  // namespace dfg
