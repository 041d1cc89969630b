// Density and density gradient on a rectangular grid — the reference's RectangularGrid (src/rectangulargrid.cpp:34-95),
// the data behind DFT::finalize's density dump (src/dft.cpp:489-504).  sm_100a only.
#pragma once
#include "common.cuh"

namespace dfg {

constexpr int kRectThreads = 256;

// x^e for the Cartesian powers the basis allows (e <= 2; e = -1 never reaches here)
__device__ __forceinline__ double rect_pow(double x, int e) { return e == 0 ? 1.0 : (e == 1 ? x : x * x); }

// CTA = PT consecutive grid points.  Phase 1: amplitudes and basis-function gradients of the PT points into shared memory
// (one (point, CGF) pair per thread and pass).  Phase 2: one warp per row i of the density matrix, lanes stride the row
// (coalesced), the four vectors (P phi, P g_x, P g_y, P g_z)_i of every point come out of one pass over the row; fixed-order
// reductions (lane tree, then warps in order), so results are bit-identical run to run.
//
// Point (i, j, k) -> index (i * dp + j) * dp + k at (k, j, i) * size/(dp-1) - size/2  (src/rectangulargrid.cpp:39-46).
// Amplitude: CGF::get_amp / GTO::get_amp (src/cgf.cpp:49-57, 146-154).  Gradient: CGF::get_grad / GTO::get_grad
// (src/cgf.cpp:67-94, 164-172) AS THE REFERENCE EVALUATES IT: separable exponentials per axis, the derivative of the
// monomial WITHOUT its factor l (x^(l-1), not l x^(l-1)), the contraction coefficient applied twice and no normalisation
// constant.  These are the reference's formulas, reproduced so that the dump is a drop-in; they are not "fixed" here.
// Density: GridPoint::set_density (src/gridpoint.cpp:82-84); gradient: GridPoint::set_gradient (src/gridpoint.cpp:94-109),
// g_x = 2 phi.(P d_x) + 2 d_x.(P phi), both terms kept (P is not assumed symmetric).
template <int PT>
__global__ void __launch_bounds__(kRectThreads)
k_rect_density(int nbf, const int* __restrict__ bf_center, const int* __restrict__ bf_prim_off, const double* __restrict__ center_xyz,
               const int* __restrict__ prim_exp, const double* __restrict__ exp_alpha, const double* __restrict__ prim_coeff,
               const double* __restrict__ prim_norm, const int* __restrict__ prim_lmn, const double* __restrict__ P /* [nbf][nbf] */,
               double size, int dp, long npts, double* __restrict__ pos /* [npts][3] */, double* __restrict__ rho /* [npts] */,
               double* __restrict__ grad /* [npts][3] */) {
    extern __shared__ __align__(16) double sm[];
    double* vec = sm;  // [4][PT][nbf]: phi, d_x, d_y, d_z
    __shared__ double red[kRectThreads / 32][PT][4];
    __shared__ double pxyz[PT][3];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long p0 = (long)blockIdx.x * PT;
    const double gd = size / (double)(dp - 1);
    if (tid < PT) {
        const long p = p0 + tid;
        const long k = p % dp, j = (p / dp) % dp, i = p / ((long)dp * dp);
        // separately rounded product and difference, like the reference's C++ (no FMA contraction): bit-identical points
        pxyz[tid][0] = __dsub_rn(__dmul_rn((double)k, gd), size / 2.0);
        pxyz[tid][1] = __dsub_rn(__dmul_rn((double)j, gd), size / 2.0);
        pxyz[tid][2] = __dsub_rn(__dmul_rn((double)i, gd), size / 2.0);
    }
    __syncthreads();
    for (int idx = tid; idx < PT * nbf; idx += kRectThreads) {
        const int pt = idx / nbf, b = idx - pt * nbf;
        const int c = bf_center[b];
        const double dx = pxyz[pt][0] - center_xyz[3 * c], dy = pxyz[pt][1] - center_xyz[3 * c + 1], dz = pxyz[pt][2] - center_xyz[3 * c + 2];
        const double r2 = dx * dx + dy * dy + dz * dz;
        double a = 0.0, gx = 0.0, gy = 0.0, gz = 0.0;
        for (int t = bf_prim_off[b]; t < bf_prim_off[b + 1]; t++) {
            const double al = exp_alpha[prim_exp[t]], cf = prim_coeff[t];
            const int lmn = prim_lmn[t], l = lmn & 15, m = (lmn >> 4) & 15, n = (lmn >> 8) & 15;
            a += cf * (prim_norm[t] * rect_pow(dx, l) * rect_pow(dy, m) * rect_pow(dz, n) * exp(-al * r2));
            const double ex = exp(-al * (dx * dx)), ey = exp(-al * (dy * dy)), ez = exp(-al * (dz * dz));
            const double fx = rect_pow(dx, l) * ex, fy = rect_pow(dy, m) * ey, fz = rect_pow(dz, n) * ez;
            double qx = -2.0 * al * dx * fx, qy = -2.0 * al * dy * fy, qz = -2.0 * al * dz * fz;
            if (l > 0) qx += rect_pow(dx, l - 1) * ex;
            if (m > 0) qy += rect_pow(dy, m - 1) * ey;
            if (n > 0) qz += rect_pow(dz, n - 1) * ez;
            gx += cf * (cf * qx * fy * fz);
            gy += cf * (cf * fx * qy * fz);
            gz += cf * (cf * fx * fy * qz);
        }
        vec[(0 * PT + pt) * nbf + b] = a;
        vec[(1 * PT + pt) * nbf + b] = gx;
        vec[(2 * PT + pt) * nbf + b] = gy;
        vec[(3 * PT + pt) * nbf + b] = gz;
    }
    __syncthreads();
    // per-warp running sums over its rows: s[pt][0] = phi.(P phi), s[pt][1..3] = phi.(P d_a) + d_a.(P phi)
    double s[PT][4];
#pragma unroll
    for (int pt = 0; pt < PT; pt++)
#pragma unroll
        for (int a = 0; a < 4; a++) s[pt][a] = 0.0;
    for (int i = warp; i < nbf; i += kRectThreads / 32) {
        double t[PT][4];
#pragma unroll
        for (int pt = 0; pt < PT; pt++)
#pragma unroll
            for (int a = 0; a < 4; a++) t[pt][a] = 0.0;
        const double* row = P + (size_t)i * nbf;
        for (int j = lane; j < nbf; j += 32) {
            const double pij = row[j];
#pragma unroll
            for (int pt = 0; pt < PT; pt++)
#pragma unroll
                for (int a = 0; a < 4; a++) t[pt][a] = fma(pij, vec[(a * PT + pt) * nbf + j], t[pt][a]);
        }
#pragma unroll
        for (int pt = 0; pt < PT; pt++)
#pragma unroll
            for (int a = 0; a < 4; a++) {
                double v = t[pt][a];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                t[pt][a] = v;  // all lanes hold (P v_a)_i
            }
#pragma unroll
        for (int pt = 0; pt < PT; pt++) {
            const double ph = vec[(0 * PT + pt) * nbf + i];
            s[pt][0] = fma(ph, t[pt][0], s[pt][0]);
#pragma unroll
            for (int a = 1; a < 4; a++) s[pt][a] += ph * t[pt][a] + vec[(a * PT + pt) * nbf + i] * t[pt][0];
        }
    }
    if (lane == 0)
#pragma unroll
        for (int pt = 0; pt < PT; pt++)
#pragma unroll
            for (int a = 0; a < 4; a++) red[warp][pt][a] = s[pt][a];
    __syncthreads();
    if (tid < PT * 4) {
        const int pt = tid >> 2, a = tid & 3;
        const long p = p0 + pt;
        if (p < npts) {
            double v = 0.0;
            for (int w = 0; w < kRectThreads / 32; w++) v += red[w][pt][a];
            v *= 2.0;
            if (a == 0) {
                rho[p] = v;
                pos[3 * p] = pxyz[pt][0];
                pos[3 * p + 1] = pxyz[pt][1];
                pos[3 * p + 2] = pxyz[pt][2];
            } else {
                grad[3 * p + a - 1] = v;
            }
        }
    }
}

}  // namespace dfg
