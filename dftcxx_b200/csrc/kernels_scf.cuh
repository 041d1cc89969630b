// Device-resident SCF algebra: what DFT::calculate_density_matrix and DFT::calculate_energy do on the
// host in the reference (src/dft.cpp:330-366, 441-447), as FP64 tensor-core (DMMA) products that keep H, X, F, P in HBM:
//
//   F  = H + F_grid                                   (F_grid = 2J + XC from the fused contraction)
//   F' = X^T F X                                      two products
//   D' = projector onto the nocc lowest eigenvectors of F'   — the only thing the reference uses its eigenvectors for
//        (Pnew = C_occ C_occ^T, C = X C'; orbital energies and C are never read again, src/dft.cpp:343-352)
//   Pnew = X D' X^T,  P = (1 - alpha) Pnew + alpha P  (first density: P = Pnew)
//   E_one = 2 tr(P H)
//
// D' is computed without an eigen-decomposition by Palser-Manolopoulos canonical purification (trace-conserving, needs no
// chemical potential): D0 = (lambda/n)(mu I - F') + (nocc/n) I from Gershgorin bounds, then
//   c = tr(D^2 - D^3) / tr(D - D^2);   D <- ((1-2c) D + (1+c) D^2 - D^3)/(1-c)  if c <= 1/2,  ((1+c) D^2 - D^3)/c otherwise
// until the idempotency error |tr(D - D^2)| stops falling (quadratic convergence; ~20 iterations at a 0.2 Ha gap, ~40 at
// 1e-4 Ha).  Every step is a dense symmetric n x n product — DMMA work — and agrees with the eigenvector construction
// to the conditioning of the occupied subspace (eps * width / gap), which is also the eigen-solver's own accuracy.
// Deterministic: fixed tile order, no atomics.  A matrix with no gap at the Fermi level does not converge; the host then
// falls back to its eigen-solver.
//
// All matrices here are row-major [np][np], np = nbf rounded up to 64, zero padded.
#pragma once
#include "common.cuh"
#include "kernels_dense.cuh"

namespace dfg {

constexpr int kGemmTile = 64;
constexpr int kGemmBK = 32;
constexpr int kGemmLdA = kGemmBK + 4;    // 36: "row = lane/4, col = lane%4" fragment loads hit 16 distinct 8-byte banks per half warp
constexpr int kGemmLdB = kGemmTile + 4;  // 68: "row = lane%4, col = lane/4" fragment loads likewise (68 mod 16 = 4)
constexpr int kGemmStages = 3;
constexpr int kGemmThreads = 128;
constexpr int kGemmStageDoubles = kGemmTile * kGemmLdA + kGemmBK * kGemmLdB;
constexpr size_t kGemmSmemBytes = (size_t)kGemmStages * kGemmStageDoubles * sizeof(double);

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// C = A B (np x np, row-major).  CTA = one 64 x 64 tile of C, 4 warps (2 x 2), warp tile 32 x 32 = 4 x 4 DMMA tiles;
// operands staged by cp.async through a 3-stage ring.  SYM: C is known to be symmetric (A, B symmetric and commuting, or
// B = A^T-like products): only the tiles on or above the diagonal are computed (grid.x = nt (nt+1)/2) and mirrored on store,
// which also makes the result symmetric to the bit.  skip: device flag; when set the kernel does nothing (lets the host
// queue a fixed batch of purification steps and stop early without a synchronisation per step).
template <bool SYM>
__global__ void __launch_bounds__(kGemmThreads) k_gemm_nn(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C, int np,
                                                          const int* __restrict__ skip) {
    if (skip && *skip) return;
    extern __shared__ __align__(16) double gsm[];
    const int nt = np / kGemmTile;
    int ti, tj;
    if (SYM) {
        int b = blockIdx.x;
        ti = 0;
        while (b >= nt - ti) {
            b -= nt - ti;
            ti++;
        }
        tj = ti + b;
    } else {
        ti = blockIdx.y;
        tj = blockIdx.x;
    }
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wm = warp >> 1, wn = warp & 1, g = lane >> 2, q = lane & 3;
    const double* Ag = A + (size_t)ti * kGemmTile * np;
    const double* Bg = B + (size_t)tj * kGemmTile;
    auto load_stage = [&](int s, int k0) {
        double* As = gsm + (size_t)s * kGemmStageDoubles;
        double* Bs = As + kGemmTile * kGemmLdA;
#pragma unroll
        for (int c = tid; c < kGemmTile * kGemmBK / 2; c += kGemmThreads) {
            const int r = c >> 4, cc = (c & 15) * 2;
            cp_async16(As + r * kGemmLdA + cc, Ag + (size_t)r * np + k0 + cc);
        }
#pragma unroll
        for (int c = tid; c < kGemmBK * kGemmTile / 2; c += kGemmThreads) {
            const int r = c >> 5, cc = (c & 31) * 2;
            cp_async16(Bs + r * kGemmLdB + cc, Bg + (size_t)(k0 + r) * np + cc);
        }
    };
    double acc[4][4][2];
#pragma unroll
    for (int mt = 0; mt < 4; mt++)
#pragma unroll
        for (int n = 0; n < 4; n++) acc[mt][n][0] = acc[mt][n][1] = 0.0;
    const int nk = np / kGemmBK;
    for (int s = 0; s < kGemmStages - 1; s++) {
        if (s < nk) load_stage(s, s * kGemmBK);
        cp_async_commit();
    }
    for (int k = 0; k < nk; k++) {
        cp_async_wait<kGemmStages - 2>();
        __syncthreads();
        const int kn = k + kGemmStages - 1;
        if (kn < nk) load_stage(kn % kGemmStages, kn * kGemmBK);
        cp_async_commit();
        const double* As = gsm + (size_t)(k % kGemmStages) * kGemmStageDoubles;
        const double* Bs = As + kGemmTile * kGemmLdA;
#pragma unroll
        for (int kk = 0; kk < kGemmBK; kk += 4) {
            double a[4], b[4];
#pragma unroll
            for (int mt = 0; mt < 4; mt++) a[mt] = As[(wm * 32 + mt * 8 + g) * kGemmLdA + kk + q];
#pragma unroll
            for (int n = 0; n < 4; n++) b[n] = Bs[(kk + q) * kGemmLdB + wn * 32 + n * 8 + g];
#pragma unroll
            for (int mt = 0; mt < 4; mt++)
#pragma unroll
                for (int n = 0; n < 4; n++) dmma884(acc[mt][n][0], acc[mt][n][1], a[mt], b[n]);
        }
    }
    cp_async_wait<0>();
    const int row0 = ti * kGemmTile + wm * 32 + g, col0 = tj * kGemmTile + wn * 32 + q * 2;
#pragma unroll
    for (int mt = 0; mt < 4; mt++)
#pragma unroll
        for (int n = 0; n < 4; n++) {
            const int r = row0 + mt * 8, c = col0 + n * 8;
            if (!SYM) {
                *reinterpret_cast<double2*>(C + (size_t)r * np + c) = make_double2(acc[mt][n][0], acc[mt][n][1]);
            } else if (ti != tj) {
                *reinterpret_cast<double2*>(C + (size_t)r * np + c) = make_double2(acc[mt][n][0], acc[mt][n][1]);
                C[(size_t)c * np + r] = acc[mt][n][0];
                C[(size_t)(c + 1) * np + r] = acc[mt][n][1];
            } else {
#pragma unroll
                for (int e = 0; e < 2; e++)
                    if (c + e >= r) {
                        C[(size_t)r * np + c + e] = acc[mt][n][e];
                        C[(size_t)(c + e) * np + r] = acc[mt][n][e];
                    }
            }
        }
}

// Symmetric product on 64 x 32 tiles.  With 64 x 64 tiles a symmetric n = 832 product has 91 CTAs for 148 SMs and each CTA
// (4 warps) keeps the DMMA pipe only ~66 % busy (profiles/r02_ncu_full_scf_h2o64.txt: 44.9 us per product).  Half-width
// tiles give 182 CTAs of half the work, two of which fit an SM (83 KB of shared memory each): every SM is busy and the SMs
// that hold two CTAs overlap their latencies.  Same operands, same k order per element as k_gemm_nn<true>: identical bits.
constexpr int kGemmBN2 = 32;
constexpr int kGemmLdB2 = kGemmBN2 + 4;  // 36 (36 mod 16 = 4, see kGemmLdB)
constexpr int kGemmStageDoubles2 = kGemmTile * kGemmLdA + kGemmBK * kGemmLdB2;
constexpr size_t kGemmSmemBytes2 = (size_t)kGemmStages * kGemmStageDoubles2 * sizeof(double);

__global__ void __launch_bounds__(kGemmThreads, 2) k_gemm_sym32(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C, int np,
                                                                const int* __restrict__ skip) {
    if (skip && *skip) return;
    extern __shared__ __align__(16) double gsm[];
    const int nt = np / kGemmTile;
    // row tile ti (64 rows) against the column tiles tj (32 columns) that reach the diagonal or lie above it: tj >= 2 ti
    int b = blockIdx.x, ti = 0;
    while (b >= 2 * (nt - ti)) {
        b -= 2 * (nt - ti);
        ti++;
    }
    const int tj = 2 * ti + b;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wm = warp >> 1, wn = warp & 1, g = lane >> 2, q = lane & 3;
    const double* Ag = A + (size_t)ti * kGemmTile * np;
    const double* Bg = B + (size_t)tj * kGemmBN2;
    auto load_stage = [&](int s, int k0) {
        double* As = gsm + (size_t)s * kGemmStageDoubles2;
        double* Bs = As + kGemmTile * kGemmLdA;
#pragma unroll
        for (int c = tid; c < kGemmTile * kGemmBK / 2; c += kGemmThreads) {
            const int r = c >> 4, cc = (c & 15) * 2;
            cp_async16(As + r * kGemmLdA + cc, Ag + (size_t)r * np + k0 + cc);
        }
#pragma unroll
        for (int c = tid; c < kGemmBK * kGemmBN2 / 2; c += kGemmThreads) {
            const int r = c >> 4, cc = (c & 15) * 2;
            cp_async16(Bs + r * kGemmLdB2 + cc, Bg + (size_t)(k0 + r) * np + cc);
        }
    };
    double acc[4][2][2];
#pragma unroll
    for (int mt = 0; mt < 4; mt++)
#pragma unroll
        for (int n = 0; n < 2; n++) acc[mt][n][0] = acc[mt][n][1] = 0.0;
    const int nk = np / kGemmBK;
    for (int s = 0; s < kGemmStages - 1; s++) {
        if (s < nk) load_stage(s, s * kGemmBK);
        cp_async_commit();
    }
    for (int k = 0; k < nk; k++) {
        cp_async_wait<kGemmStages - 2>();
        __syncthreads();
        const int kn = k + kGemmStages - 1;
        if (kn < nk) load_stage(kn % kGemmStages, kn * kGemmBK);
        cp_async_commit();
        const double* As = gsm + (size_t)(k % kGemmStages) * kGemmStageDoubles2;
        const double* Bs = As + kGemmTile * kGemmLdA;
#pragma unroll
        for (int kk = 0; kk < kGemmBK; kk += 4) {
            double a[4], bb[2];
#pragma unroll
            for (int mt = 0; mt < 4; mt++) a[mt] = As[(wm * 32 + mt * 8 + g) * kGemmLdA + kk + q];
#pragma unroll
            for (int n = 0; n < 2; n++) bb[n] = Bs[(kk + q) * kGemmLdB2 + wn * 16 + n * 8 + g];
#pragma unroll
            for (int mt = 0; mt < 4; mt++)
#pragma unroll
                for (int n = 0; n < 2; n++) dmma884(acc[mt][n][0], acc[mt][n][1], a[mt], bb[n]);
        }
    }
    cp_async_wait<0>();
    const int row0 = ti * kGemmTile + wm * 32 + g, col0 = tj * kGemmBN2 + wn * 16 + q * 2;
    const bool above = tj * kGemmBN2 >= (ti + 1) * kGemmTile;  // the whole tile lies above the diagonal
#pragma unroll
    for (int mt = 0; mt < 4; mt++)
#pragma unroll
        for (int n = 0; n < 2; n++) {
            const int r = row0 + mt * 8, c = col0 + n * 8;
            if (above) {
                *reinterpret_cast<double2*>(C + (size_t)r * np + c) = make_double2(acc[mt][n][0], acc[mt][n][1]);
                C[(size_t)c * np + r] = acc[mt][n][0];
                C[(size_t)(c + 1) * np + r] = acc[mt][n][1];
            } else {
#pragma unroll
                for (int e = 0; e < 2; e++)
                    if (c + e >= r) {
                        C[(size_t)r * np + c + e] = acc[mt][n][e];
                        C[(size_t)(c + e) * np + r] = acc[mt][n][e];
                    }
            }
        }
}


// dst[np][np] (zero padded) = a[nb][nb] (+ b[nb][nb])
__global__ void k_scf_pad_sum(const double* __restrict__ a, const double* __restrict__ b, int nb, int np, double* __restrict__ dst) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)np * np) return;
    const int i = (int)(t / np), j = (int)(t % np);
    double v = 0.0;
    if (i < nb && j < nb) v = a[(size_t)i * nb + j] + (b ? b[(size_t)i * nb + j] : 0.0);
    dst[t] = v;
}

// dst[np][np] = transpose of src[np][np]
__global__ void k_scf_transpose(const double* __restrict__ src, int np, double* __restrict__ dst) {
    __shared__ double tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) tile[r][threadIdx.x] = src[(size_t)(by + r) * np + bx + threadIdx.x];
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) dst[(size_t)(bx + r) * np + by + threadIdx.x] = tile[threadIdx.x][r];
}

// Purification state in device memory.
struct PmState {
    double a, b;              // D0 = a I + b F'
    double c, err, err_prev;  // current coefficient, |tr(D - D^2)| of this and of the previous step
    double emin, emax, mu;    // Gershgorin bounds and mean of the spectrum of F'
    int done, iters, failed, pad;
};

// Gershgorin discs of the n x n matrix F (row-major, leading dimension np): warp per row.
__global__ void k_pm_gershgorin(const double* __restrict__ F, int n, int np, double* __restrict__ lo, double* __restrict__ hi, double* __restrict__ diag) {
    const int row = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (row >= n) return;
    double s = 0.0;
    for (int j = lane; j < n; j += 32)
        if (j != row) s += fabs(F[(size_t)row * np + j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) {
        const double d = F[(size_t)row * np + row];
        lo[row] = d - s;
        hi[row] = d + s;
        diag[row] = d;
    }
}

// fixed-order block reduction helper (blockDim.x = 256)
__device__ __forceinline__ double block_sum256(double v, double* part) {
    part[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
        __syncthreads();
    }
    const double r = part[0];
    __syncthreads();
    return r;
}

// One block: spectrum bounds -> coefficients of D0 (Palser & Manolopoulos, Phys. Rev. B 58, 12704, eq. 14-16).
__global__ void k_pm_setup(const double* __restrict__ lo, const double* __restrict__ hi, const double* __restrict__ diag, int n, int nocc,
                           PmState* __restrict__ st) {
    __shared__ double part[256];
    __shared__ double smin[256], smax[256];
    double mn = 1e300, mx = -1e300, tr = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        mn = fmin(mn, lo[i]);
        mx = fmax(mx, hi[i]);
        tr += diag[i];
    }
    smin[threadIdx.x] = mn;
    smax[threadIdx.x] = mx;
    const double trace = block_sum256(tr, part);
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            smin[threadIdx.x] = fmin(smin[threadIdx.x], smin[threadIdx.x + o]);
            smax[threadIdx.x] = fmax(smax[threadIdx.x], smax[threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double emin = smin[0], emax = smax[0], mu = trace / n;
        const double l1 = (double)nocc / (emax - mu), l2 = (double)(n - nocc) / (mu - emin);
        const double lam = fmin(l1, l2);
        st->emin = emin;
        st->emax = emax;
        st->mu = mu;
        st->b = -lam / n;
        st->a = lam * mu / n + (double)nocc / n;
        st->c = 0.0;
        st->err = st->err_prev = 1e300;
        st->done = 0;
        st->iters = 0;
        st->failed = (emax - mu > 0.0 && mu - emin > 0.0) ? 0 : 1;  // a multiple of the identity: nothing to purify
        if (st->failed) st->done = 1;
    }
}

// D = a I + b F' on the leading n x n block, zero elsewhere.
__global__ void k_pm_init(const double* __restrict__ F, int n, int np, const PmState* __restrict__ st, double* __restrict__ D) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)np * np) return;
    const int i = (int)(t / np), j = (int)(t % np);
    double v = 0.0;
    if (i < n && j < n) v = st->b * F[t] + (i == j ? st->a : 0.0);
    D[t] = v;
}

// One block: traces from the diagonals -> c, idempotency error, stopping decision.
__global__ void k_pm_coeff(const double* __restrict__ D, const double* __restrict__ D2, const double* __restrict__ D3, int n, int np,
                           PmState* __restrict__ st) {
    __shared__ double part[256];
    if (st->done) return;
    double num = 0.0, den = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const size_t d = (size_t)i * np + i;
        num += D2[d] - D3[d];
        den += D[d] - D2[d];
    }
    num = block_sum256(num, part);
    den = block_sum256(den, part);
    if (threadIdx.x == 0) {
        const double e = fabs(den), ep = st->err;
        st->err_prev = ep;
        st->err = e;
        // quadratic end game: once the error is small and has stopped falling by orders of magnitude it sits on the
        // rounding floor; den == 0 exactly means D is idempotent to the last bit
        if (e == 0.0 || (e < 1e-7 && ep < 1e299 && (e > 0.01 * ep || e < 1e-14))) {
            st->done = 1;
        } else {
            st->c = num / den;
            st->iters += 1;
        }
    }
}

// D <- PM update with the coefficient of k_pm_coeff (elementwise, keeps the zero padding and the symmetry).
__global__ void k_pm_update(double* __restrict__ D, const double* __restrict__ D2, const double* __restrict__ D3, int np, const PmState* __restrict__ st) {
    if (st->done) return;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)np * np) return;
    const double c = st->c, d = D[t], d2 = D2[t], d3 = D3[t];
    D[t] = c <= 0.5 ? ((1.0 - 2.0 * c) * d + (1.0 + c) * d2 - d3) / (1.0 - c) : ((1.0 + c) * d2 - d3) / c;
}

// D = identity on the leading n x n block (every orbital occupied: nocc == n)
__global__ void k_scf_identity(int n, int np, double* __restrict__ D) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)np * np) return;
    const int i = (int)(t / np), j = (int)(t % np);
    D[t] = (i == j && i < n) ? 1.0 : 0.0;
}

// P[nb][nb] = first ? Pnew : (1 - alpha) Pnew + alpha P   (src/dft.cpp:354-359); Pnew is [np][np]
__global__ void k_scf_mix(const double* __restrict__ Pnew, int nb, int np, double alpha, int first, double* __restrict__ P) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)nb * nb) return;
    const int i = (int)(t / nb), j = (int)(t % nb);
    const double pn = Pnew[(size_t)i * np + j];
    P[t] = first ? pn : (1.0 - alpha) * pn + alpha * P[t];
}

// rowsum[i] = sum_j A[i][j] B[i][j] (warp per row, fixed order); k_scf_trace_finish adds the rows in order: 2 tr(A B) for symmetric A, B
__global__ void k_scf_rowdot(const double* __restrict__ A, const double* __restrict__ B, int nb, double* __restrict__ rowsum) {
    const int row = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (row >= nb) return;
    double s = 0.0;
    for (int j = lane; j < nb; j += 32) s = fma(A[(size_t)row * nb + j], B[(size_t)row * nb + j], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) rowsum[row] = s;
}
__global__ void k_scf_trace_finish(const double* __restrict__ rowsum, int nb, double scale, double* __restrict__ out) {
    __shared__ double part[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) s += rowsum[i];
    s = block_sum256(s, part);
    if (threadIdx.x == 0) *out = scale * s;
}

}  // namespace dfg
