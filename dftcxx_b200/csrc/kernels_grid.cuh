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
    // m^3 rounded once: m*m = p + e exactly (e from the FMA residual), so m^3 = p*m + e*m and the final FMA rounds
    // the exact p*m plus the tiny (rounded) e*m.  Verified against the 6-operation double-double variant on 4e7
    // samples (identical) — see DESIGN.md for how it compares with glibc's pow(m, 3.0).
    const double p = __dmul_rn(m, m);
    const double e = __fma_rn(m, m, -p);
    return __fma_rn(p, m, __dmul_rn(e, m));
}
__device__ __forceinline__ double becke_cutoff(double mu) {
    // f(mu) = 1.5*mu - 0.5*mu^3, three times: 0.5*cube is exact, so one FMA reproduces the separately rounded
    // product and difference of the reference expression (src/moleculargrid.cpp:325)
#pragma unroll
    for (int it = 0; it < 3; it++) mu = __fma_rn(-0.5, cube_rn(mu), __dmul_rn(1.5, mu));
    return __dmul_rn(0.5, __dsub_rn(1.0, mu));
}

constexpr int kBeckeWarps = 4;

__global__ void __launch_bounds__(kBeckeWarps * 32)
k_becke(GridShape g, const double* __restrict__ atom_xyz,
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
// Thread = point.  The host groups the columns into SHELLS — runs of consecutive CGFs on one centre that share
// their primitives: S (1 column), P (px,py,pz), D (xx,xy,xz,yy,yz,zz), or a generic single column — so the
// primitive record (coefficient, norms, exponent slot) is loaded once per shell and the inner loops are branch
// free.  Each distinct exponent of a centre is exponentiated once per point (the reference recomputes it per
// primitive); exp() is skipped where alpha*r^2 > 746 (the result is exactly +0 in FP64 there, as on the CPU).
// Multiplying by pow(x,0) = 1.0 is exact, so the generic path multiplies by a selected factor instead of branching.
// A [kPhiPts x kPhiCols] tile is staged in shared memory so that every Phi row is written with full 256-byte
// coalesced segments; pad columns [nbf, nbp) are written as zeros.
constexpr int kPhiPts = 128;
constexpr int kPhiCols = 32;    // one 256-byte row segment per point and pass
constexpr int kPhiBatch = 3;    // primitives of a shell fetched together (6-31G shells have 1, 3 or 6)
constexpr int kPhiMaxExp = 24;  // distinct exponents on one centre (STO-6G third row needs 18)

enum { kShellGeneric = 0, kShellS = 1, kShellP = 2, kShellD = 3 };

struct PhiPrim {  // one primitive of a shell, 32 bytes
    double coeff;
    double norm_a;  // S, P: the norm; D: norm of xx, yy, zz; generic: the norm
    double norm_b;  // D: norm of xy, xz, yz
    int exp_slot;   // index into the centre's distinct-exponent list
    int lmn;        // generic only: l | m<<4 | n<<8
};

struct PhiShell {  // 32 bytes
    int type, col, centre, prim_off, nprim, ncol, pad0, pad1;
};

struct PhiBasis {
    int nbf, nbp, nshell;
    const PhiShell* shells;     // ordered by first column; columns are contiguous across shells
    const PhiPrim* prims;
    const int* centre_exp_off;  // [ncentres+1]
    const double* exp_alpha;    // distinct exponents, centre after centre
    const double* centre_xyz;   // [ncentres][3]
    const int* pass_shell_rng;  // [npass][2]: first / one-past-last shell intersecting column pass c (kPhiCols columns each)
};

__global__ void __launch_bounds__(kPhiPts, 4)
k_phi(long nloc, PhiBasis B, const double* __restrict__ px, const double* __restrict__ py,
      const double* __restrict__ pz, double* __restrict__ phi) {
    extern __shared__ double sm[];
    double* tile = sm;                                   // [kPhiPts][kPhiCols+1]
    double* ex = sm + (size_t)kPhiPts * (kPhiCols + 1);  // [largest centre's exponent count][kPhiPts]
    const int tid = threadIdx.x;
    const long p0 = (long)blockIdx.x * kPhiPts;
    const long p = p0 + tid;
    const bool live = p < nloc;
    const double x = live ? px[p] : 0.0, y = live ? py[p] : 0.0, z = live ? pz[p] : 0.0;
    double* trow = tile + (size_t)tid * (kPhiCols + 1);
    const double* exl = ex + tid;
    int cur = -1;
    double dx = 0, dy = 0, dz = 0;
    const int npass = B.nbp / kPhiCols;
    for (int pass = 0; pass < npass; pass++) {
        const int c0 = pass * kPhiCols;
#pragma unroll 4
        for (int c = 0; c < kPhiCols; c++) trow[c] = 0.0;  // pad columns and columns of shells outside this pass
        const int s0 = __ldg(B.pass_shell_rng + 2 * pass), s1 = __ldg(B.pass_shell_rng + 2 * pass + 1);
        // the shell and primitive records are the same for every thread (warp-uniform loads); what costs is their
        // latency in front of each dependent step, so the next shell's record is fetched one iteration ahead and a
        // shell's primitives are fetched together before any of them is used
        int4 sa_n = make_int4(0, 0, 0, 0);
        int2 sb_n = make_int2(0, 0);
        if (s0 < s1) {
            sa_n = __ldg(reinterpret_cast<const int4*>(B.shells + s0));
            sb_n = __ldg(reinterpret_cast<const int2*>(B.shells + s0) + 2);
        }
        for (int si = s0; si < s1; si++) {
            const int4 sa = sa_n;  // type, col, centre, prim_off
            const int2 sb = sb_n;  // nprim, ncol
            if (si + 1 < s1) {
                sa_n = __ldg(reinterpret_cast<const int4*>(B.shells + si + 1));
                sb_n = __ldg(reinterpret_cast<const int2*>(B.shells + si + 1) + 2);
            }
            if (sa.z != cur) {
                cur = sa.z;
                dx = __dsub_rn(x, __ldg(B.centre_xyz + 3 * cur));
                dy = __dsub_rn(y, __ldg(B.centre_xyz + 3 * cur + 1));
                dz = __dsub_rn(z, __ldg(B.centre_xyz + 3 * cur + 2));
                const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                const int e0 = __ldg(B.centre_exp_off + cur), e1 = __ldg(B.centre_exp_off + cur + 1);
                // exponentials of this centre, four at a time: the four evaluations are independent instruction chains
                // (one exp() is ~25 dependent FP64 operations), and a group is skipped altogether when no lane of the warp
                // needs it (alpha r^2 > 746 underflows to exactly +0, the common case for tight exponents far away)
                for (int u = e0; u < e1; u += 4) {
                    double arg[4];
#pragma unroll
                    for (int t = 0; t < 4; t++) arg[t] = u + t < e1 ? __dmul_rn(__ldg(B.exp_alpha + u + t), r2) : 1e300;
                    const bool need = arg[0] <= 746.0 || arg[1] <= 746.0 || arg[2] <= 746.0 || arg[3] <= 746.0;
                    double ev[4] = {0.0, 0.0, 0.0, 0.0};
                    if (__any_sync(0xffffffffu, need)) {
#pragma unroll
                        for (int t = 0; t < 4; t++) {
                            const double e_ = exp(-fmin(arg[t], 746.0));
                            ev[t] = arg[t] > 746.0 ? 0.0 : e_;
                        }
                    }
#pragma unroll
                    for (int t = 0; t < 4; t++)
                        if (u + t < e1) ex[(size_t)(u + t - e0) * kPhiPts + tid] = ev[t];
                }
            }
            const PhiPrim* pr = B.prims + sa.w;
            const int rel = sa.y - c0;  // first column of the shell relative to this pass (may be negative)
            if (sa.x == kShellS) {
                double v = 0.0;
                for (int k0 = 0; k0 < sb.x; k0 += kPhiBatch) {
                    double2 cn[kPhiBatch];
                    double e[kPhiBatch];
#pragma unroll
                    for (int u = 0; u < kPhiBatch; u++)
                        if (k0 + u < sb.x) cn[u] = __ldg(reinterpret_cast<const double2*>(pr + k0 + u));
#pragma unroll
                    for (int u = 0; u < kPhiBatch; u++)
                        if (k0 + u < sb.x) e[u] = exl[(size_t)__ldg(reinterpret_cast<const int*>(pr + k0 + u) + 6) * kPhiPts];
#pragma unroll
                    for (int u = 0; u < kPhiBatch; u++)
                        if (k0 + u < sb.x) v = __dadd_rn(v, __dmul_rn(cn[u].x, __dmul_rn(cn[u].y, e[u])));
                }
                trow[rel] = v;
            } else if (sa.x == kShellP) {
                double v0 = 0.0, v1 = 0.0, v2 = 0.0;
                for (int k0 = 0; k0 < sb.x; k0 += kPhiBatch) {
                    double2 cn[kPhiBatch];
                    double e[kPhiBatch];
#pragma unroll
                    for (int u = 0; u < kPhiBatch; u++)
                        if (k0 + u < sb.x) cn[u] = __ldg(reinterpret_cast<const double2*>(pr + k0 + u));
#pragma unroll
                    for (int u = 0; u < kPhiBatch; u++)
                        if (k0 + u < sb.x) e[u] = exl[(size_t)__ldg(reinterpret_cast<const int*>(pr + k0 + u) + 6) * kPhiPts];
#pragma unroll
                    for (int u = 0; u < kPhiBatch; u++)
                        if (k0 + u < sb.x) {
                            v0 = __dadd_rn(v0, __dmul_rn(cn[u].x, __dmul_rn(__dmul_rn(cn[u].y, dx), e[u])));
                            v1 = __dadd_rn(v1, __dmul_rn(cn[u].x, __dmul_rn(__dmul_rn(cn[u].y, dy), e[u])));
                            v2 = __dadd_rn(v2, __dmul_rn(cn[u].x, __dmul_rn(__dmul_rn(cn[u].y, dz), e[u])));
                        }
                }
                if (rel >= 0 && rel < kPhiCols) trow[rel] = v0;
                if (rel + 1 >= 0 && rel + 1 < kPhiCols) trow[rel + 1] = v1;
                if (rel + 2 >= 0 && rel + 2 < kPhiCols) trow[rel + 2] = v2;
            } else if (sa.x == kShellD) {
                double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                const double xx = __dmul_rn(dx, dx), yy = __dmul_rn(dy, dy), zz = __dmul_rn(dz, dz);
                for (int k = 0; k < sb.x; k++) {
                    const double2 cn = __ldg(reinterpret_cast<const double2*>(pr + k));
                    const double nb_ = __ldg(reinterpret_cast<const double*>(pr + k) + 2);
                    const int slot = __ldg(reinterpret_cast<const int*>(pr + k) + 6);
                    const double e = exl[(size_t)slot * kPhiPts];
                    v[0] = __dadd_rn(v[0], __dmul_rn(cn.x, __dmul_rn(__dmul_rn(cn.y, xx), e)));
                    v[1] = __dadd_rn(v[1], __dmul_rn(cn.x, __dmul_rn(__dmul_rn(__dmul_rn(nb_, dx), dy), e)));
                    v[2] = __dadd_rn(v[2], __dmul_rn(cn.x, __dmul_rn(__dmul_rn(__dmul_rn(nb_, dx), dz), e)));
                    v[3] = __dadd_rn(v[3], __dmul_rn(cn.x, __dmul_rn(__dmul_rn(cn.y, yy), e)));
                    v[4] = __dadd_rn(v[4], __dmul_rn(cn.x, __dmul_rn(__dmul_rn(__dmul_rn(nb_, dy), dz), e)));
                    v[5] = __dadd_rn(v[5], __dmul_rn(cn.x, __dmul_rn(__dmul_rn(cn.y, zz), e)));
                }
#pragma unroll
                for (int t = 0; t < 6; t++)
                    if (rel + t >= 0 && rel + t < kPhiCols) trow[rel + t] = v[t];
            } else {  // generic single column: factors selected, never branched on
                double v = 0.0;
                const double xx = __dmul_rn(dx, dx), yy = __dmul_rn(dy, dy), zz = __dmul_rn(dz, dz);
                for (int k = 0; k < sb.x; k++) {
                    const double2 cn = __ldg(reinterpret_cast<const double2*>(pr + k));
                    const int2 sl = __ldg(reinterpret_cast<const int2*>(pr + k) + 3);  // exp_slot, lmn
                    const int l = sl.y & 15, m = (sl.y >> 4) & 15, n = (sl.y >> 8) & 15;
                    const double fx = l == 0 ? 1.0 : (l == 1 ? dx : xx);
                    const double fy = m == 0 ? 1.0 : (m == 1 ? dy : yy);
                    const double fz = n == 0 ? 1.0 : (n == 1 ? dz : zz);
                    const double a = __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(cn.y, fx), fy), fz), exl[(size_t)sl.x * kPhiPts]);
                    v = __dadd_rn(v, __dmul_rn(cn.x, a));
                }
                trow[rel] = v;
            }
        }
        __syncthreads();
        for (int row = tid >> 5; row < kPhiPts; row += kPhiPts / 32) {
            const long pr_ = p0 + row;
            if (pr_ >= nloc) break;
            phi[pr_ * B.nbp + c0 + (tid & 31)] = tile[(size_t)row * (kPhiCols + 1) + (tid & 31)];
        }
        __syncthreads();
    }
}

}  // namespace dfg
