// Pointwise LDA, charge sums and the Becke/Poisson Hartree potential kernels.
#pragma once
#include "common.cuh"
#include "host_tables.h"
#include <type_traits>

namespace dfg {

// ---------------------------------------------------------------------------------------------------------
// Per-shell sums  out[gshell*stride + slot] = sum_a w*v  (warp per local shell, fixed reduction order).
// Used for sum(w rho) (src/atomicgrid.cpp:520-530).
__global__ void k_shell_sum(GridShape g, const double* __restrict__ w, const double* __restrict__ v,
                            double* __restrict__ out, int stride, int slot) {
    const long s = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (s >= g.nshell_loc) return;
    double acc = 0.0;
    for (int a = lane; a < g.nang; a += 32) acc = fma(w[s * g.nang + a], v[s * g.nang + a], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if (lane == 0) out[(g.shell0 + s) * stride + slot] = acc;
}

// Single-block bookkeeping: per-atom charges q_atom = sum over the atom's shells (radial order), total, and
// (mode 0) the rescale factor sum(Z)/sum(w rho) of MolecularGrid::correct_densities (src/moleculargrid.cpp:132-146),
// or (mode 1) electron count + E_xc written to results[0..1].
__global__ void k_totals(GridShape g, const double* __restrict__ shellsum, int stride, double zsum, int mode,
                         double* __restrict__ q_atom, double* __restrict__ scalars /*[0]=scale [1]=nel [2]=exc*/) {
    __shared__ double part[256];
    for (int a = threadIdx.x; a < g.natoms; a += blockDim.x) {
        double q = 0.0;
        for (int i = 0; i < g.nrad; i++) q += shellsum[((long)a * g.nrad + i) * stride + 0];
        q_atom[a] = q;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int a = 0; a < g.natoms; a++) tot += q_atom[a];
        if (mode == 0) scalars[0] = zsum / tot;
        if (mode == 1) scalars[1] = tot;
    }
    if (mode == 1) {
        // E_xc = sum over shells of the per-shell sums (fixed order: strided partials, then tree)
        double e = 0.0;
        const long ns = (long)g.natoms * g.nrad;
        for (long s = threadIdx.x; s < ns; s += blockDim.x) e += shellsum[s * stride + 1];
        part[threadIdx.x] = e;
        __syncthreads();
        for (int o = blockDim.x / 2; o > 0; o >>= 1) {
            if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) scalars[2] = part[0];
    }
}

// Fused Fock build: one weight vector for F_grid = 2J + XC = Phi^T diag(w (V + v_xc)) Phi (src/dft.cpp:334 only ever
// uses the sum H + 2J + XC).  dJ = w V, dxc = w v_xc; without_xc reproduces the reference's first iteration, whose F holds
// J(P0) but XC = 0 (src/dft.cpp:219-226).
__global__ void k_fock_weights(long nloc, const double* __restrict__ dJ, const double* __restrict__ dxc, int include_xc, double* __restrict__ dF) {
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nloc) return;
    dF[p] = include_xc ? dJ[p] + dxc[p] : dJ[p];
}

// E_J = 2 tr(P J) (src/dft.cpp:443) without forming J: 2 sum_ij P_ij J_ij = sum_p w_p V_p (rho_raw,p / 2), rho_raw the
// density of P BEFORE the rescale = rho / scale.  shellsum[s] = sum over the shell's points of (w V) * rho (k_shell_sum),
// summed per atom in radial order, then over the atoms in order (sharding independent).  out = [e_j, exc, nel].
__global__ void k_fock_scalars(GridShape g, const double* __restrict__ shellsum, const double* __restrict__ scalars, double* __restrict__ q_tmp,
                               double* __restrict__ out) {
    for (int a = threadIdx.x; a < g.natoms; a += blockDim.x) {
        double q = 0.0;
        for (int i = 0; i < g.nrad; i++) q += shellsum[(long)a * g.nrad + i];
        q_tmp[a] = q;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int a = 0; a < g.natoms; a++) tot += q_tmp[a];
        out[0] = 0.5 * tot / scalars[0];
        out[1] = scalars[2];
        out[2] = scalars[1];
    }
}

// ---------------------------------------------------------------------------------------------------------
// rho *= scale, then Slater-Xalpha exchange + VWN5 correlation for a closed shell (rho_a = rho_b = rho/2):
// src/functionals.cpp:24-63 (exchange), 65-114 with zeta = 0 => g = 0 < tol, paramagnetic branch only, 116-150.
//   dxc[p] = w * ((vxa+vxb+vca+vcb) * 0.5)      (src/dft.cpp:424)
//   exw[p] = ex + ec                             (E_xc = sum w*(ex+ec), src/dft.cpp:417)
__device__ __forceinline__ double vwn_X(double x, double b, double c) { return x * x + b * x + c; }

__global__ void k_scale_xc(long nloc, LdaConstants K, const double* __restrict__ scalars, const double* __restrict__ w,
                           double* __restrict__ rho, double* __restrict__ dxc, double* __restrict__ exw) {
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nloc) return;
    const double r = rho[p] * scalars[0];
    rho[p] = r;
    const double ra = r * 0.5;
    double ex = 0.0, vxa = 0.0;
    if (!(ra < 1e-10)) {
        const double rho3 = pow(ra, 1.0 / 3.0);
        const double t = K.fac * ra * rho3;
        ex = t + t;
        vxa = K.vfac * rho3;
    }
    double ec = 0.0, vca = 0.0;
    const double dens = ra + ra;
    if (!(dens < 1e-10)) {
        const double x = pow(K.x_pref / dens, 1.0 / 6.0);
        const double Xx = vwn_X(x, K.b, K.c);
        const double twoxb = 2.0 * x + K.b;
        const double epsp = K.a * (log(x * x / Xx) - K.bx0_over_Xx0 * log((x - K.x0) * (x - K.x0) / Xx) + K.atan_pref * atan(K.q / twoxb));
        const double den2 = twoxb * twoxb + K.q * K.q;
        const double depsp = K.a * (2.0 / x - twoxb / Xx - 4.0 * K.b / den2 -
                                    (K.b * K.x0 / K.Xx0) * (2.0 / (x - K.x0) - twoxb / Xx - 4.0 * (2.0 * K.x0 + K.b) / den2));
        ec = epsp * dens;
        vca = epsp - (x / 6.0) * depsp;
    }
    dxc[p] = w[p] * ((vxa + vxa + vca + vca) * 0.5);
    exw[p] = ex + ec;
}

// ---------------------------------------------------------------------------------------------------------
// rho_lm(i, lm) = 4 pi sum_j rho(i,j) * Y_lm(j) * w_leb(j) * w_becke(i,j)   (src/atomicgrid.cpp:258-295)
// Block per local shell, thread per lm; j runs in the reference's order.
__global__ void k_rho_lm(GridShape g, const double* __restrict__ rho, const double* __restrict__ wb,
                         const double* __restrict__ leb, const double* __restrict__ Y /*[nang][nlm]*/,
                         double* __restrict__ rho_lm /*[natoms*nrad][nlm]*/) {
    extern __shared__ double sm[];
    double* rs = sm;               // rho
    double* ws = sm + g.nang;      // lebedev weight
    double* bs = sm + 2 * g.nang;  // becke weight
    const long s = blockIdx.x;
    for (int j = threadIdx.x; j < g.nang; j += blockDim.x) {
        rs[j] = rho[s * g.nang + j];
        ws[j] = leb[4 * j + 3];
        bs[j] = wb[s * g.nang + j];
    }
    __syncthreads();
    const int lm = threadIdx.x;
    if (lm >= g.nlm) return;
    double acc = 0.0;
    for (int j = 0; j < g.nang; j++) acc += rs[j] * Y[(long)j * g.nlm + lm] * ws[j] * bs[j];
    rho_lm[(g.shell0 + s) * g.nlm + lm] = acc * (4.0 * 3.14159265358979323846);
}

// ---------------------------------------------------------------------------------------------------------
// Radial Poisson solve for every (atom, lm): M_l U = g with the pre-factorised operators (host_tables.h),
// g_0 = sqrt(4 pi) q_atom for lm = 0 else 0, g_i = -4 pi r_i rho_lm(i), g_{N+1} = 0  (src/atomicgrid.cpp:395-432).
// Thread per system, atom fastest so a warp shares l (the LU entries are then warp-uniform broadcasts).
// work: [(N+2)][nsys] scratch.
__global__ void k_poisson(GridShape g, int n /*N+2*/, const double* __restrict__ lu, const int* __restrict__ perm,
                          const int* __restrict__ lo, const int* __restrict__ hi, const double* __restrict__ r_tab,
                          const double* __restrict__ rho_lm, const double* __restrict__ q_atom,
                          double* __restrict__ work, double* __restrict__ U_lm /*[natoms*nrad][nlm]*/, bool use_smem) {
    const long nsys = (long)g.natoms * g.nlm;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nsys) return;
    const int atom = (int)(t % g.natoms), lm = (int)(t / g.natoms);
    int l = 0;
    while ((l + 1) * (l + 1) <= lm) l++;
    const double* M = lu + (size_t)l * n * n;
    const int* pr = perm + (size_t)l * n;
    const int* lor = lo + (size_t)l * n;
    const int* hir = hi + (size_t)l * n;
    const int N = n - 2;
    const double sqrt4pi = 3.5449077018110318;  // sqrt(4*M_PI)
    auto rhs = [&](int i) -> double {
        if (i == 0) return lm == 0 ? sqrt4pi * q_atom[atom] : 0.0;
        if (i == N + 1) return 0.0;
        return -4.0 * 3.14159265358979323846 * r_tab[i - 1] * rho_lm[((long)atom * g.nrad + (i - 1)) * g.nlm + lm];
    };
    // scratch vector of this system: shared memory [n][blockDim] when the launch provides it (small radial grids: the
    // substitution is a chain of dependent loads, ~30 clocks each from shared memory against a global round trip),
    // else the global work array [n][nsys]
    extern __shared__ double sx[];
    double* x = use_smem ? sx + threadIdx.x : work + t;
    const size_t xs_ = use_smem ? (size_t)blockDim.x : (size_t)nsys;
    for (int i = 0; i < n; i++) {  // L y = P g
        double acc = rhs(pr[i]);
        for (int j = lor[i]; j < i; j++) acc -= M[(size_t)i * n + j] * x[(size_t)j * xs_];
        x[(size_t)i * xs_] = acc;
    }
    for (int i = n - 1; i >= 0; i--) {  // U x = y
        double acc = x[(size_t)i * xs_];
        for (int j = hir[i]; j > i; j--) acc -= M[(size_t)i * n + j] * x[(size_t)j * xs_];
        x[(size_t)i * xs_] = acc / M[(size_t)i * n + i];
    }
    for (int i = 1; i < N + 1; i++) U_lm[((long)atom * g.nrad + (i - 1)) * g.nlm + lm] = x[(size_t)i * xs_];
}

// ---------------------------------------------------------------------------------------------------------
// Not-a-knot cubic splines through (r ascending, U_lm) for every (atom, lm): Cspline::generate_spline
// (src/cspline.cpp:66-142) with the x-only part of the tridiagonal sweep precomputed on the host.
// Output table coef[atom][interval][lm][4] (a,b,c,d); interval N-1 is the clamp y.back() (src/cspline.cpp:159-161).
// Position of (l, m) inside one (atom, interval) block of the coefficient table: the +m and -m records sit next to
// each other (64 contiguous bytes), because the interpolation kernel always needs them together.
//   slot(l, 0) = l^2,  slot(l, +m) = l^2 + 2m - 1,  slot(l, -m) = l^2 + 2m
__host__ __device__ __forceinline__ int coef_slot_lm(int l, int m) { return l * l + (m == 0 ? 0 : (m > 0 ? 2 * m - 1 : -2 * m)); }
__host__ __device__ __forceinline__ int coef_slot(int lm) {
    int l = 0;
    while ((l + 1) * (l + 1) <= lm) l++;
    return coef_slot_lm(l, lm - l * l - l);
}

struct SplineDev {
    const double* x;    // [N]
    const double* A;    // [N]
    const double* Cp;   // [N]
    const double* den;  // [N]
    const double* h;    // [N]
    const double* rh;   // [N]
    double fw0, fh0, lh1, lw1;
};

__global__ void k_spline(GridShape g, SplineDev S, const double* __restrict__ U_lm, const double* __restrict__ pre /*[(lmax+1)^2]*/,
                         double* __restrict__ work /*[2][N][nsys]*/, double* __restrict__ coef, bool use_smem) {
    const long nsys = (long)g.natoms * g.nlm;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nsys) return;
    const int lm = (int)(t % g.nlm), atom = (int)(t / g.nlm);
    const int N = g.nrad;
    auto y = [&](int i) -> double { return U_lm[((long)atom * N + (N - 1 - i)) * g.nlm + lm]; };  // ascending r
    const int slot = coef_slot(lm);
    // the records carry the real-Y_lm prefactor pre_{l|m|} (src/spherical_harmonics.cpp:28-33), so the interpolation
    // kernels multiply P_l^|m| {cos|sin} by the record value only
    double pf;
    {
        int l = 0;
        while ((l + 1) * (l + 1) <= lm) l++;
        const int m = lm - l * l - l;
        pf = pre[l * (g.lmax + 1) + (m < 0 ? -m : m)];
    }
    // scratch [2][N] per system: shared memory [2N][blockDim] when provided, else the global work array (stride nsys)
    extern __shared__ double sx[];
    double* Y = use_smem ? sx + threadIdx.x : work + t;
    double* D = use_smem ? sx + (size_t)N * blockDim.x + threadIdx.x : work + (size_t)N * nsys + t;
    const size_t ws_ = use_smem ? (size_t)blockDim.x : (size_t)nsys;
    // right-hand sides (src/cspline.cpp:81-109)
    {
        const double r0 = (y(1) - y(0)) / S.h[0], r1 = (y(2) - y(1)) / S.h[1];
        Y[0] = r0 * S.fw0 + r1 * S.fh0 * S.fh0;
    }
    double r0 = 0.0, r1 = 0.0;
    for (int i = 1; i < N - 1; i++) {
        r0 = (y(i) - y(i - 1)) / S.h[i - 1];
        r1 = (y(i + 1) - y(i)) / S.h[i];
        Y[(size_t)i * ws_] = 3 * (r0 * S.h[i] + r1 * S.h[i - 1]);
    }
    Y[(size_t)(N - 1) * ws_] = r0 * S.lh1 * S.lh1 + r1 * S.lw1;
    // forward sweep and back substitution (src/cspline.cpp:111-127)
    Y[0] = Y[0] / S.den[0];
    for (int i = 1; i < N; i++) Y[(size_t)i * ws_] = (Y[(size_t)i * ws_] - S.A[i] * Y[(size_t)(i - 1) * ws_]) / S.den[i];
    D[(size_t)(N - 1) * ws_] = Y[(size_t)(N - 1) * ws_];
    for (int i = N - 1; i > 0; i--) D[(size_t)(i - 1) * ws_] = Y[(size_t)(i - 1) * ws_] - S.Cp[i - 1] * D[(size_t)i * ws_];
    // polynomial coefficients (src/cspline.cpp:130-139)
    for (int i = 0; i < N - 1; i++) {
        const double dx = S.rh[i];
        const double dy = (y(i + 1) - y(i)) * dx;
        const double Di = D[(size_t)i * ws_], Dn = D[(size_t)(i + 1) * ws_];
        double4 c;
        c.x = pf * y(i);
        c.y = pf * Di;
        c.z = pf * (dx * (3 * dy - 2 * Di - Dn));
        c.w = pf * (dx * dx * (-2 * dy + Di + Dn));
        *reinterpret_cast<double4*>(coef + (((size_t)atom * N + i) * g.nlm + slot) * 4) = c;
    }
    *reinterpret_cast<double4*>(coef + (((size_t)atom * N + (N - 1)) * g.nlm + slot) * 4) = make_double4(pf * y(N - 1), 0.0, 0.0, 0.0);
}

// ---------------------------------------------------------------------------------------------------------
// Own-cell potential V_fuzzy(p) = sum_lm (1/r) * Y_lm(p) * U_lm(i)  (src/atomicgrid.cpp:437-462).
// Block per local shell; Yt is the transposed table [nlm][nang].
__global__ void k_v_own(GridShape g, const double* __restrict__ r_tab, const double* __restrict__ leb,
                        const double* __restrict__ Yt, const double* __restrict__ U_lm, double* __restrict__ Vown) {
    extern __shared__ double sm[];
    const long s = blockIdx.x;
    const long gs = g.shell0 + s;
    const int i = (int)(gs % g.nrad);
    for (int lm = threadIdx.x; lm < g.nlm; lm += blockDim.x) sm[lm] = U_lm[gs * g.nlm + lm];
    __syncthreads();
    const double r = r_tab[i];
    for (int j = threadIdx.x; j < g.nang; j += blockDim.x) {
        const double qx = __dmul_rn(leb[4 * j], r), qy = __dmul_rn(leb[4 * j + 1], r), qz = __dmul_rn(leb[4 * j + 2], r);
        const double rr = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(qx, qx), __dmul_rn(qy, qy)), __dmul_rn(qz, qz)));
        const double rinv = 1.0 / rr;
        double v = 0.0;
        for (int lm = 0; lm < g.nlm; lm++) v += rinv * Yt[(long)lm * g.nang + j] * sm[lm];
        Vown[s * g.nang + j] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Cross-atom interpolation (src/moleculargrid.cpp:342-380): for point p of atom i
//   V(p) = sum_k [ k == i ? V_fuzzy(p) : sum_lm (1/r) pre_lm P_l^|m|(cos th) {cos m ph | sin |m| ph} S_{k,lm}(r) ]
// with r, th, ph the spherical coordinates of p - R_k and S the clamped cubic spline (src/cspline.cpp:151-172).
// Thread per point, atoms k in ascending order.  Y_lm is generated once per (p,k) by the reference's own
// Legendre recurrences (src/spherical_harmonics.cpp:81-117) run column-wise in m, and cos/sin(m ph) by the
// angle-addition recurrence from cos ph = x/rho_xy, sin ph = y/rho_xy (atan2(0,0) = 0 => cos = 1, sin = 0).
// The radial interval is found by bisection on the shared abscissa.  dJ[p] = w[p]*V[p] feeds the J contraction.
constexpr int kMaxL = 15;

__global__ void __launch_bounds__(128)
k_interp(GridShape g, const double* __restrict__ atom_xyz, const double* __restrict__ px, const double* __restrict__ py,
         const double* __restrict__ pz, const double* __restrict__ w, const double* __restrict__ Vown,
         const double* __restrict__ xs /*[N] ascending*/, const double* __restrict__ pre /*[(lmax+1)^2]*/,
         const double* __restrict__ coef, double* __restrict__ Vpart /*[gridDim.y][nloc]*/) {
    extern __shared__ double sm[];
    const int N = g.nrad, L = g.lmax;
    double* xsh = sm;            // [N]
    double* rj = sm + N;         // [2L+2] reciprocals 1/(j-m)
    for (int i = threadIdx.x; i < N; i += blockDim.x) xsh[i] = xs[i];
    for (int i = threadIdx.x; i < 2 * L + 2; i += blockDim.x) rj[i] = i > 0 ? 1.0 / (double)i : 0.0;
    __syncthreads();
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= g.nloc) return;
    const int own = (int)((g.shell0 + p / g.nang) / g.nrad);
    const double x = px[p], y = py[p], z = pz[p];
    double Vacc = 0.0;
    const int k_begin = (int)((long)g.natoms * blockIdx.y / gridDim.y), k_end = (int)((long)g.natoms * (blockIdx.y + 1) / gridDim.y);
    for (int k = k_begin; k < k_end; k++) {
        if (k == own) {
            Vacc += Vown[p];
            continue;
        }
        const double dx = x - atom_xyz[3 * k], dy = y - atom_xyz[3 * k + 1], dz = z - atom_xyz[3 * k + 2];
        const double r = sqrt(dx * dx + dy * dy + dz * dz);
        const double rinv = 1.0 / r;
        // spline interval: r < x0 -> (0, t=0); r >= x_last -> (N-1, t=0); else first i with r <= x_i -> (i-1, r - x_{i-1})
        int iv;
        double tt;
        if (r < xsh[0]) {
            iv = 0;
            tt = 0.0;
        } else if (r >= xsh[N - 1]) {
            iv = N - 1;
            tt = 0.0;
        } else {
            int lo_ = 0, hi_ = N - 1;  // invariant: x[lo_] < r <= x[hi_]  (r == x[0] handled: first i with r <= x_i is 1)
            if (r == xsh[0]) {
                hi_ = 1;
            } else {
                while (hi_ - lo_ > 1) {
                    const int mid = (lo_ + hi_) >> 1;
                    if (r <= xsh[mid])
                        hi_ = mid;
                    else
                        lo_ = mid;
                }
            }
            iv = hi_ - 1;
            tt = r - xsh[iv];
        }
        const double* cf = coef + ((size_t)k * N + iv) * g.nlm * 4;
        const double ct = dz / r;  // cos(theta) = z/r by true division (exactly +-1 on the axis, like the reference's cos(acos(z/r)))
        const double st = sqrt(1.0 - ct * ct);
        const double rxy = sqrt(dx * dx + dy * dy);
        double c1 = 1.0, s1 = 0.0;
        if (rxy > 0.0) {
            c1 = dx / rxy;
            s1 = dy / rxy;
        }
        double cm = 1.0, sn = 0.0;  // cos(m phi), sin(m phi)
        double pmm = 1.0;           // P_m^m
        double fact = 1.0;
        double sum = 0.0;
        for (int m = 0; m <= L; m++) {
            if (m > 0) {
                pmm *= -fact * st;
                fact += 2.0;
                const double cn = cm * c1 - sn * s1;
                sn = sn * c1 + cm * s1;
                cm = cn;
            }
            double pl2 = 0.0, pl1 = pmm;  // P_{l-2}^m, P_{l-1}^m while iterating l
            for (int l = m; l <= L; l++) {
                double pl;
                if (l == m)
                    pl = pmm;
                else if (l == m + 1)
                    pl = ct * (double)(2 * m + 1) * pmm;
                else
                    pl = ((double)(2 * l - 1) * ct * pl1 + (double)(-l - m + 1) * pl2) * rj[l - m];
                pl2 = pl1;
                pl1 = pl;
                const double pf = rinv;  // pre_lm is folded into the spline records
                const int lmp = coef_slot_lm(l, m);
                {
                    const double4 c = *reinterpret_cast<const double4*>(cf + (size_t)lmp * 4);
                    const double sv = c.x + c.y * tt + c.z * tt * tt + c.w * tt * tt * tt;
                    sum += pf * (pl * cm) * sv;
                }
                if (m > 0) {
                    const double4 c = *reinterpret_cast<const double4*>(cf + (size_t)(lmp + 1) * 4);
                    const double sv = c.x + c.y * tt + c.z * tt * tt + c.w * tt * tt * tt;
                    sum += pf * (pl * sn) * sv;
                }
            }
        }
        Vacc += sum;
    }
    Vpart[(size_t)blockIdx.y * g.nloc + p] = Vacc;
}

// Fully unrolled variant for a compile-time lmax (the presets use 5, 8, 10, 11): all (l, m) loop indices, recurrence
// constants and table offsets become immediates, which removes the integer/convert/branch overhead that dominates the
// generic kernel (about 30 instructions per (l,m) term there).  Same arithmetic, same accumulation order.
template <int L, int MINB>
__global__ void __launch_bounds__(128, MINB)
k_interp_t(GridShape g, const double* __restrict__ atom_xyz, const double* __restrict__ px, const double* __restrict__ py,
           const double* __restrict__ pz, const double* __restrict__ w, const double* __restrict__ Vown,
           const double* __restrict__ xs, const double* __restrict__ pre, const double* __restrict__ coef,
           double* __restrict__ Vpart /*[gridDim.y][nloc]*/) {
    extern __shared__ __align__(128) double sm[];
    const int N = g.nrad;
    constexpr int NLM = (L + 1) * (L + 1);
    // Spline-record staging: the 32 points of a warp usually fall into one or two adjacent radial intervals of a source
    // atom, so the warp copies those one or two (atom, interval) rows (NLM records of 32 bytes, contiguous) into shared
    // memory once and every lane then reads its records as broadcast LDS instead of 2*NLM divergent global loads.
    // Warps spanning more than two intervals fall back to per-lane global loads through the same (generic) pointer.
    double4* rows = reinterpret_cast<double4*>(sm) + (size_t)(threadIdx.x >> 5) * 2 * NLM;  // [4 warps][2][NLM]
    double* xsh = sm + (size_t)4 * 2 * NLM * 4;                                              // [N]
    for (int i = threadIdx.x; i < N; i += blockDim.x) xsh[i] = xs[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long pr = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = pr < g.nloc;
    const long p = live ? pr : g.nloc - 1;  // idle lanes of the last warp shadow the last point (never stored)
    const int own = (int)((g.shell0 + p / g.nang) / g.nrad);
    const double x = px[p], y = py[p], z = pz[p];
    const double vown = Vown[p];
    double Vacc = 0.0;
    // blockIdx.y selects a contiguous chunk of source atoms: short CTAs keep the tail wave small when a rank holds
    // only a fraction of the points; the chunk sums are added in chunk order by k_finish_potential
    const int k_begin = (int)((long)g.natoms * blockIdx.y / gridDim.y), k_end = (int)((long)g.natoms * (blockIdx.y + 1) / gridDim.y);
    for (int k = k_begin; k < k_end; k++) {
        const double dx = x - atom_xyz[3 * k], dy = y - atom_xyz[3 * k + 1], dz = z - atom_xyz[3 * k + 2];
        const double r = sqrt(dx * dx + dy * dy + dz * dz);
        const double rinv = 1.0 / r;
        int iv;
        double tt;
        if (r < xsh[0]) {
            iv = 0;
            tt = 0.0;
        } else if (r >= xsh[N - 1]) {
            iv = N - 1;
            tt = 0.0;
        } else {
            int lo_ = 0, hi_ = N - 1;  // x[lo_] < r <= x[hi_] (or r == x[0]): first i with r <= x_i
            while (hi_ - lo_ > 1) {
                const int mid = (lo_ + hi_) >> 1;
                if (r <= xsh[mid])
                    hi_ = mid;
                else
                    lo_ = mid;
            }
            iv = hi_ - 1;
            tt = r - xsh[iv];
        }
        const double4* grow = reinterpret_cast<const double4*>(coef + ((size_t)k * N) * NLM * 4);  // atom k, interval 0
        const int iv_min = __reduce_min_sync(0xffffffffu, iv), iv_max = __reduce_max_sync(0xffffffffu, iv);
        const double4* cf;
        __syncwarp();  // the previous atom's rows are no longer being read
        if (iv_max - iv_min <= 1) {
            const int nrec = (iv_max - iv_min + 1) * NLM;  // the two rows are contiguous in the table
            const double4* src = grow + (size_t)iv_min * NLM;
            for (int i = lane; i < nrec; i += 32) rows[i] = src[i];
            __syncwarp();
            cf = rows + (iv - iv_min) * NLM;
        } else {
            cf = grow + (size_t)iv * NLM;
        }
        const double tt2 = tt * tt, tt3 = tt2 * tt;
        const double ct = dz / r;
        const double st = sqrt(1.0 - ct * ct);
        const double rxy = sqrt(dx * dx + dy * dy);
        double c1 = 1.0, s1 = 0.0;
        if (rxy > 0.0) {
            c1 = dx / rxy;
            s1 = dy / rxy;
        }
        double cm = 1.0, sn = 0.0, pmm = 1.0;
        double acc4[4] = {0.0, 0.0, 0.0, 0.0};  // independent partial sums: no single serial FMA chain
#pragma unroll
        for (int m = 0; m <= L; m++) {
            if (m > 0) {
                pmm *= -(double)(2 * m - 1) * st;
                const double cn = cm * c1 - sn * s1;
                sn = sn * c1 + cm * s1;
                cm = cn;
            }
            double pl2 = 0.0, pl1 = pmm;
#pragma unroll
            for (int l = m; l <= L; l++) {
                double pl;
                if (l == m)
                    pl = pmm;
                else if (l == m + 1)
                    pl = ct * (double)(2 * m + 1) * pmm;
                else
                    pl = ((double)(2 * l - 1) * ct * pl1 + (double)(-l - m + 1) * pl2) * (1.0 / (double)(l - m));
                pl2 = pl1;
                pl1 = pl;
                const double pf = rinv;  // pre_lm is folded into the spline records
                {
                    const double4 c = cf[l * l + (m == 0 ? 0 : 2 * m - 1)];
                    const double sv = c.x + c.y * tt + c.z * tt2 + c.w * tt3;
                    acc4[(l & 1) * 2] += pf * (pl * cm) * sv;
                }
                if (m > 0) {
                    const double4 c = cf[l * l + 2 * m];
                    const double sv = c.x + c.y * tt + c.z * tt2 + c.w * tt3;
                    acc4[(l & 1) * 2 + 1] += pf * (pl * sn) * sv;
                }
            }
        }
        const double sum = (acc4[0] + acc4[1]) + (acc4[2] + acc4[3]);
        Vacc += k == own ? vown : sum;  // own cell: the tabulated V_fuzzy instead of the interpolated expansion
    }
    if (live) Vpart[(size_t)blockIdx.y * g.nloc + p] = Vacc;
}

// V = sum of the atom-chunk partial potentials in chunk order; dJ = w * V feeds the J contraction
__global__ void k_finish_potential(long nloc, int nchunk, const double* __restrict__ Vpart, const double* __restrict__ w,
                                   double* __restrict__ V, double* __restrict__ dJ) {
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nloc) return;
    double v = 0.0;
    for (int c = 0; c < nchunk; c++) v += Vpart[(size_t)c * nloc + p];
    V[p] = v;
    dJ[p] = w[p] * v;
}

// ---------------------------------------------------------------------------------------------------------
// Binned cross-atom interpolation.
//
// The unrolled kernel above is bound by the register-fill bandwidth of the on-chip load path: every lane needs the
// 32-byte spline record of each (l,m) term (8 L1/shared wavefronts per warp and term against 4.5 clocks of FP64
// work).  Which record a (point p, source atom k) pair needs depends only on the geometry: the radial interval
// ("row") of |p - R_k| on the shared abscissa.  So at grid-build time all pairs are sorted into bins keyed by
// (k, row); at every SCF iteration one warp evaluates 32*R pairs of ONE bin, each lane holding R points in registers,
// and every record fetched from the warp's staged copy of the row is used R times.  Per-pair results go to out[slot]
// and k_finish_binned adds a point's contributions in source-atom order, exactly the order of the reference's loop
// (src/moleculargrid.cpp:350-377), so the result does not depend on how the bins were filled.
__device__ __forceinline__ double pair_dist(double dx, double dy, double dz) { return sqrt(dx * dx + dy * dy + dz * dz); }

// Clamped interval lookup of Cspline::eval (src/cspline.cpp:151-172): row N-1 is the y.back() clamp.
__device__ __forceinline__ int spline_row(double r, const double* xsh, int N) {
    if (r < xsh[0]) return 0;
    if (r >= xsh[N - 1]) return N - 1;
    int lo_ = 0, hi_ = N - 1;  // x[lo_] < r <= x[hi_] (or r == x[0]): first i with r <= x_i
    while (hi_ - lo_ > 1) {
        const int mid = (lo_ + hi_) >> 1;
        if (r <= xsh[mid])
            hi_ = mid;
        else
            lo_ = mid;
    }
    return hi_ - 1;
}

// Pass 1 (cursor == nullptr): counts[k*N + row] += 1 for every local (p, k != own) pair.
// Pass 2: slot = binoff[key] + position inside the bin; pair_point[slot] = p, slot_of[k][p] = slot (-1 for k == own).
// Lanes of a warp that hit the same bin are aggregated into one atomic.
__global__ void __launch_bounds__(256)
k_bin_pairs(GridShape g, const double* __restrict__ atom_xyz, const double* __restrict__ px, const double* __restrict__ py,
            const double* __restrict__ pz, const double* __restrict__ xs, int* __restrict__ counts, const int* __restrict__ binoff,
            int* __restrict__ cursor, int* __restrict__ pair_point, int* __restrict__ slot_of) {
    extern __shared__ double xsh[];
    const int N = g.nrad;
    for (int i = threadIdx.x; i < N; i += blockDim.x) xsh[i] = xs[i];
    __syncthreads();
    const long pr = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = pr < g.nloc;
    const long p = live ? pr : g.nloc - 1;
    const int lane = threadIdx.x & 31;
    const int own = (int)((g.shell0 + p / g.nang) / g.nrad);
    const double x = px[p], y = py[p], z = pz[p];
    for (int k = 0; k < g.natoms; k++) {
        int row = -1;
        if (live && k != own) row = spline_row(pair_dist(x - atom_xyz[3 * k], y - atom_xyz[3 * k + 1], z - atom_xyz[3 * k + 2]), xsh, N);
        const unsigned grp = __match_any_sync(0xffffffffu, row);
        const int leader = __ffs(grp) - 1;
        const int key = k * N + row;
        if (cursor == nullptr) {
            if (row >= 0 && lane == leader) atomicAdd(&counts[key], __popc(grp));
        } else {
            int base = 0;
            if (row >= 0 && lane == leader) base = atomicAdd(&cursor[key], __popc(grp));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (row >= 0) {
                const int slot = binoff[key] + base + __popc(grp & ((1u << lane) - 1u));
                pair_point[slot] = (int)p;
                slot_of[(size_t)k * g.nloc + p] = slot;
            } else if (live) {
                slot_of[(size_t)k * g.nloc + p] = -1;
            }
        }
    }
}

// item_key[item] = bin of the item's 32*R slots: the largest key with binoff[key] <= first slot (empty bins repeat the
// offset of the next one, so the last of a run of equal offsets is the non-empty bin).
__global__ void k_item_keys(const int* __restrict__ binoff, int nkeys, int unit, long nitems, int* __restrict__ item_key) {
    const long item = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= nitems) return;
    const long slot0 = item * unit;
    int key = 0, hi_ = nkeys;
    while (hi_ - key > 1) {
        const int mid = (key + hi_) >> 1;
        if ((long)binoff[mid] <= slot0)
            key = mid;
        else
            hi_ = mid;
    }
    item_key[item] = key;
}

// Scale of the Legendre recurrence used by k_interp_bin (see there): K_l^m, and the one remaining constant A'_l^m.
__host__ __device__ constexpr double legendre_scale(int l, int m) {
    return l <= m + 1 ? 1.0 : (double)(l + m - 1) / (double)(l - m) * legendre_scale(l - 2, m);
}
__host__ __device__ constexpr double legendre_scaled_a(int l, int m) {
    return (double)(2 * l - 1) / (double)(l - m) * legendre_scale(l - 1, m) / legendre_scale(l, m);
}

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

// Geometry of one (point, source atom) pair as the interpolation consumes it: 1/r, the spline abscissa t = r - x_row
// (0 below the first node: the spline is clamped to y.front()), cos / sin(theta), cos / sin(phi).
struct PairGeo {
    double rinv, tt, ct, st, c1, s1;
};
__device__ __forceinline__ PairGeo pair_geometry(double dx, double dy, double dz, double x_first, double x_row) {
    PairGeo q;
    const double r = pair_dist(dx, dy, dz);
    q.rinv = 1.0 / r;
    q.tt = r < x_first ? 0.0 : r - x_row;
    q.ct = dz / r;  // cos(theta) = z/r by true division: exactly +-1 on the axis, like cos(acos(z/r))
    q.st = sqrt(1.0 - q.ct * q.ct);
    const double rxy = sqrt(dx * dx + dy * dy);
    q.c1 = 1.0;
    q.s1 = 0.0;  // atan2(0,0) = 0
    if (rxy > 0.0) {
        const double ri = 1.0 / rxy;
        q.c1 = dx * ri;
        q.s1 = dy * ri;
    }
    return q;
}

// The pairs never move during an SCF (points and nuclei are fixed), so their geometry is evaluated ONCE at grid build, in
// bin order, as six SoA arrays geo[q][slot] (48 bytes per pair: 5.1 GB at (H2O)64).  k_interp_bin<.., GEO = true> then starts
// from six coalesced loads instead of a dependent gather (slot -> point -> x, y, z) followed by two square roots and three
// divisions per pair: that prologue held 30 % of the kernel's stall samples (profiles/r02c_ncu_hot_h2o64.txt).  Pad slots
// get the same harmless dummy the kernel used to build for them.
__global__ void k_pair_geometry(GridShape g, const double* __restrict__ atom_xyz, const double* __restrict__ px, const double* __restrict__ py,
                                const double* __restrict__ pz, const double* __restrict__ xs, const int* __restrict__ item_key,
                                const int* __restrict__ pair_point, int unit, long nslots, double* __restrict__ geo) {
    const long slot = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= nslots) return;
    const int key = item_key[slot / unit];
    const int N = g.nrad, k = key / N, row = key - k * N;
    const int p = pair_point[slot];
    double dx = 1.0, dy = 1.0, dz = 1.0;
    if (p >= 0) {
        dx = px[p] - atom_xyz[3 * k];
        dy = py[p] - atom_xyz[3 * k + 1];
        dz = pz[p] - atom_xyz[3 * k + 2];
    }
    const PairGeo q = pair_geometry(dx, dy, dz, xs[0], xs[row]);
    geo[slot] = q.rinv;
    geo[nslots + slot] = q.tt;
    geo[2 * nslots + slot] = q.ct;
    geo[3 * nslots + slot] = q.st;
    geo[4 * nslots + slot] = q.c1;
    geo[5 * nslots + slot] = q.s1;
}

constexpr int kBinWarps = 4;  // warps (= work items) per CTA of the binned kernel

// One warp per item = 32*R consecutive slots of one bin (bins are padded to a multiple of 32*R; pad slots hold -1).
// Arithmetic per pair: Legendre columns by the reference's recurrences (src/spherical_harmonics.cpp:81-117),
// cos/sin(m phi) by angle addition, cubic in Horner form; 1/r is applied once to the pair's sum.
template <int L, int R, int MINB, bool GEO = false>
__global__ void __launch_bounds__(kBinWarps * 32, MINB)
k_interp_bin(GridShape g, const double* __restrict__ atom_xyz, const double* __restrict__ px, const double* __restrict__ py,
             const double* __restrict__ pz, const double* __restrict__ xs, const double* __restrict__ coef,
             const int* __restrict__ item_key, const int* __restrict__ pair_point, long nitems, double* __restrict__ out,
             const double* __restrict__ geo = nullptr, long nslots = 0) {
    constexpr int NLM = (L + 1) * (L + 1);
    extern __shared__ __align__(128) double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long item = (long)blockIdx.x * kBinWarps + warp;
    if (item >= nitems) return;
    double4* rowbuf = reinterpret_cast<double4*>(sm) + (size_t)warp * NLM;
    const long slot0 = item * (32 * R);
    // independent loads first: the item's bin and its R point indices per lane
    const int key = item_key[item];
    int pidx[R];
#pragma unroll
    for (int j = 0; j < R; j++) pidx[j] = pair_point[slot0 + j * 32 + lane];
    const int N = g.nrad;
    const int k = key / N, row = key - k * N;
    {
        // stage the bin's row of records with asynchronous copies; they land while the lane's geometry is set up
        const double2* src = reinterpret_cast<const double2*>(coef) + (size_t)key * (2 * NLM);  // 16-byte units
        double2* dst2 = reinterpret_cast<double2*>(rowbuf);
        for (int i = lane; i < 2 * NLM; i += 32) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(dst2 + i);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src + i) : "memory");
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    }
    double tt[R], ct[R], st[R], c1[R], s1[R], rinv[R], cm[R], sn[R], pmm[R], acc[R];
    bool live[R];
    if constexpr (GEO) {
        (void)row;
#pragma unroll
        for (int j = 0; j < R; j++) {
            const long s = slot0 + j * 32 + lane;
            live[j] = true;  // pad slots hold the dummy geometry and own a result slot nobody reads
            rinv[j] = geo[s];
            tt[j] = geo[nslots + s];
            ct[j] = geo[2 * nslots + s];
            st[j] = geo[3 * nslots + s];
            c1[j] = geo[4 * nslots + s];
            s1[j] = geo[5 * nslots + s];
            cm[j] = 1.0;
            sn[j] = 0.0;
            pmm[j] = 1.0;
            acc[j] = 0.0;
        }
    } else {
        const double ax = atom_xyz[3 * k], ay = atom_xyz[3 * k + 1], az = atom_xyz[3 * k + 2];
        const double x_first = xs[0], x_row = xs[row];
#pragma unroll
        for (int j = 0; j < R; j++) {
            const int p = pidx[j];
            live[j] = p >= 0;
            double dx = 1.0, dy = 1.0, dz = 1.0;  // pad lanes evaluate a harmless dummy
            if (live[j]) {
                dx = px[p] - ax;
                dy = py[p] - ay;
                dz = pz[p] - az;
            }
            const PairGeo q = pair_geometry(dx, dy, dz, x_first, x_row);
            rinv[j] = q.rinv;
            tt[j] = q.tt;
            ct[j] = q.ct;
            st[j] = q.st;
            c1[j] = q.c1;
            s1[j] = q.s1;
            cm[j] = 1.0;
            sn[j] = 0.0;
            pmm[j] = 1.0;
            acc[j] = 0.0;
        }
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    __syncwarp();
    // compile-time (m, l) loops: every recurrence constant and record offset is an immediate.  The Legendre column is run
    // in the scaled form s_l = P_l^m / K_l^m with K chosen so that the three-term recurrence loses its second constant,
    //   (l-m) P_l = (2l-1) x P_{l-1} - (l+m-1) P_{l-2}   ->   s_l = A'_l x s_{l-1} - s_{l-2},
    //   K_m = K_{m+1} = 1,  K_l = (l+m-1)/(l-m) K_{l-2},  A'_l = (2l-1)/(l-m) K_{l-1}/K_l     (legendre_scale below),
    // two FP64 operations per (l, m) instead of three; K_l^m is folded into the spline records by k_spline (scaled
    // prefactor table), like the Y_lm prefactor.  The cos / sin(m phi) factors are applied once per m to the two l-sums.
    static_for<0, L + 1>([&](auto mc) {
        constexpr int m = decltype(mc)::value;
        double pl1[R], pl2[R], accc[R], accs[R];
#pragma unroll
        for (int j = 0; j < R; j++) {
            if (m > 0) {
                pmm[j] *= -(double)(2 * m - 1) * st[j];
                const double cn = cm[j] * c1[j] - sn[j] * s1[j];
                sn[j] = sn[j] * c1[j] + cm[j] * s1[j];
                cm[j] = cn;
            }
            pl2[j] = 0.0;
            pl1[j] = pmm[j];
            accc[j] = 0.0;
            accs[j] = 0.0;
        }
        static_for<m, L + 1>([&](auto lc) {
            constexpr int l = decltype(lc)::value;
            const double4 cc = rowbuf[l * l + (m == 0 ? 0 : 2 * m - 1)];
            double4 cs = make_double4(0.0, 0.0, 0.0, 0.0);
            if (m > 0) cs = rowbuf[l * l + 2 * m];
#pragma unroll
            for (int j = 0; j < R; j++) {
                double pl;
                if (l == m) {
                    pl = pmm[j];
                } else if (l == m + 1) {
                    pl = ct[j] * (double)(2 * m + 1) * pmm[j];
                } else {
                    constexpr double AP = legendre_scaled_a(l, m);
                    pl = fma(AP * ct[j], pl1[j], -pl2[j]);
                }
                pl2[j] = pl1[j];
                pl1[j] = pl;
                const double svc = fma(fma(fma(cc.w, tt[j], cc.z), tt[j], cc.y), tt[j], cc.x);
                accc[j] = fma(pl, svc, accc[j]);
                if (m > 0) {
                    const double svs = fma(fma(fma(cs.w, tt[j], cs.z), tt[j], cs.y), tt[j], cs.x);
                    accs[j] = fma(pl, svs, accs[j]);
                }
            }
        });
#pragma unroll
        for (int j = 0; j < R; j++) {
            if (m > 0)
                acc[j] += fma(accs[j], sn[j], accc[j] * cm[j]);
            else
                acc[j] += accc[j];
        }
    });
#pragma unroll
    for (int j = 0; j < R; j++)
        if (live[j]) out[slot0 + j * 32 + lane] = acc[j] * rinv[j];
}

// V(p) = sum over source atoms in ascending order of (own cell: tabulated V_fuzzy | interpolated expansion);
// dJ = w * V feeds the J contraction.
__global__ void k_finish_binned(GridShape g, const int* __restrict__ slot_of, const double* __restrict__ out,
                                const double* __restrict__ Vown, const double* __restrict__ w, double* __restrict__ V,
                                double* __restrict__ dJ) {
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= g.nloc) return;
    const double vown = Vown[p];
    double v = 0.0;
    int k = 0;
    for (; k + 4 <= g.natoms; k += 4) {
        int s[4];
#pragma unroll
        for (int u = 0; u < 4; u++) s[u] = slot_of[(size_t)(k + u) * g.nloc + p];
        double c[4];
#pragma unroll
        for (int u = 0; u < 4; u++) c[u] = s[u] < 0 ? vown : out[s[u]];
#pragma unroll
        for (int u = 0; u < 4; u++) v += c[u];
    }
    for (; k < g.natoms; k++) {
        const int s = slot_of[(size_t)k * g.nloc + p];
        v += s < 0 ? vown : out[s];
    }
    V[p] = v;
    dJ[p] = w[p] * v;
}

}  // namespace dfg
