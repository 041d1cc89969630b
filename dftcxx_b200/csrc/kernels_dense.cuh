// Dense FP64 contractions on the tensor pipe (DMMA, mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4 on sm_100a;
// tcgen05/wgmma have no f64 kind, so warp-level DMMA fed from cp.async-staged shared memory is the
// Blackwell FP64 tensor path).  Measured register-resident peak on this pool's B200: 37.05 TFLOP/s
// (profiles/r01_fp64_peak_microbench.txt), identical to the DFMA peak, at a quarter of the issue slots.
//
//   k_rho      : rho_p = 2 * sum_n Phi[p][n] * (sum_k Phi[p][k] P[k][n])   (src/gridpoint.cpp:82-84)
//   k_contract : C_z[i][j] = sum_p Phi[p][i] d_z[p] Phi[p][j], z in {XC, J} (src/dft.cpp:424-432, src/atomicgrid.cpp:471-488)
//   k_contract_reduce : fixed-order sum of the split-K partials, mirrored into both triangles.
//
// Shared-memory tiles are padded so that every fragment load is bank-conflict free:
//   [rows][32+4] doubles for "row = lane/4, col = lane%4" accesses, [rows][128+8] for "row = lane%4, col = lane/4".
#pragma once
#include "common.cuh"

namespace dfg {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int sz = valid ? 16 : 0;  // src-size 0 => 16 zero bytes are written
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

constexpr int kDenseThreads = 256;  // 8 warps: 4 along M x 2 along N
constexpr int kTileM = 128;
constexpr int kTileN = 128;
constexpr int kTileK = 32;
constexpr int kLdK = kTileK + 4;    // 36
constexpr int kLdN = kTileN + 8;    // 136
constexpr int kStages = 3;

// =========================================================================================================
// rho
// =========================================================================================================
constexpr int kRhoStageDoubles = kTileM * kLdK + kTileK * kLdN;
constexpr size_t kRhoSmemBytes = (size_t)kStages * kRhoStageDoubles * sizeof(double) + 2 * kTileM * sizeof(double);

__device__ __forceinline__ void rho_load_stage(double* st, const double* __restrict__ phi, const double* __restrict__ P,
                                               long p0, long nloc, int nbp, int slab, int kc) {
    double* As = st;                      // [128][36]   Phi[p0+r][kc + c]
    double* Bs = st + kTileM * kLdK;      // [32][136]   P[kc + r][slab + c]
    const int tid = threadIdx.x;
#pragma unroll
    for (int it = 0; it < 8; it++) {
        const int ch = tid + it * kDenseThreads;  // 0..2047
        const int r = ch >> 4, c = (ch & 15) * 2;
        const long p = p0 + r;
        const bool ok = p < nloc;
        cp_async16(As + r * kLdK + c, phi + (ok ? p : 0) * (long)nbp + kc + c, ok);
    }
#pragma unroll
    for (int it = 0; it < 8; it++) {
        const int ch = tid + it * kDenseThreads;
        const int r = ch >> 6, c = (ch & 63) * 2;
        const bool ok = slab + c < nbp;
        cp_async16(Bs + r * kLdN + c, P + (long)(kc + r) * nbp + (ok ? slab + c : 0), ok);
    }
}

template <int NT>
__device__ __forceinline__ void rho_mma_stage(const double* st, double (&acc)[4][8][2], int wm, int wn, int lane) {
    const double* As = st;
    const double* Bs = st + kTileM * kLdK;
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int kk = 0; kk < kTileK; kk += 4) {
        double a[4], b[NT];
#pragma unroll
        for (int mt = 0; mt < 4; mt++) a[mt] = As[(wm * 32 + mt * 8 + g) * kLdK + kk + q];
#pragma unroll
        for (int nt = 0; nt < NT; nt++) b[nt] = Bs[(kk + q) * kLdN + wn * (NT * 8) + nt * 8 + g];
#pragma unroll
        for (int mt = 0; mt < 4; mt++)
#pragma unroll
            for (int nt = 0; nt < NT; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
    }
}

// Schedule of one CTA's (column slab, k-chunk) steps.  P is symmetric, so for column slab J only the k-chunks at or
// beyond the slab are visited: first the chunks past the slab's end (their contribution counts twice: P[n][k] and
// P[k][n]), then — after doubling the accumulators, which is exact — the chunks inside the slab.  This halves the
// DMMA work of rho = 2 sum_n phi_n (sum_k P_nk phi_k) relative to the full product.
struct RhoStep {
    int slab, i, nbp;
    __device__ __forceinline__ int nk() const { return nbp / kTileK; }
    __device__ __forceinline__ int first_chunk() const { return slab * (kTileN / kTileK); }
    __device__ __forceinline__ int end_chunk() const { return min(nk(), (slab + 1) * (kTileN / kTileK)); }
    __device__ __forceinline__ int count() const { return nk() - first_chunk(); }
    __device__ __forceinline__ int outer() const { return nk() - end_chunk(); }  // chunks past the slab
    __device__ __forceinline__ int chunk() const { return i < outer() ? end_chunk() + i : first_chunk() + (i - outer()); }
    __device__ __forceinline__ void advance() {
        if (++i == count()) {
            i = 0;
            slab++;
        }
    }
};

// grid.x = ceil(nloc/128); P is the zero-padded [nbp][nbp] density matrix (symmetric).
__global__ void __launch_bounds__(kDenseThreads, 1)
k_rho(const double* __restrict__ phi, const double* __restrict__ P, double* __restrict__ rho, long nloc, int nbp) {
    extern __shared__ __align__(16) double sm[];
    double* red = sm + (size_t)kStages * kRhoStageDoubles;  // [2][128]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp & 3, wn = warp >> 2;
    const int g = lane >> 2, q = lane & 3;
    const long p0 = (long)blockIdx.x * kTileM;
    const int nk = nbp / kTileK;
    const int nslab = (nbp + kTileN - 1) / kTileN;
    int total = 0;
    for (int J = 0; J < nslab; J++) total += nk - J * (kTileN / kTileK);

    double rowsum[4] = {0.0, 0.0, 0.0, 0.0};
    double acc[4][8][2];

    RhoStep ld{0, 0, nbp}, cs{0, 0, nbp};
    for (int s = 0; s < kStages - 1; s++) {
        if (s < total) {
            rho_load_stage(sm + (size_t)s * kRhoStageDoubles, phi, P, p0, nloc, nbp, ld.slab * kTileN, ld.chunk() * kTileK);
            ld.advance();
        }
        cp_async_commit();
    }
    for (int it = 0; it < total; it++) {
        const int slab = cs.slab * kTileN;
        const int ncols = min(kTileN, nbp - slab);  // 32, 64, 96 or 128
        const bool narrow = ncols <= 64;            // 8 warps as 4 x 2 over [128 x 64]: 4 n-tiles per warp
        if (cs.i == 0) {
#pragma unroll
            for (int mt = 0; mt < 4; mt++)
#pragma unroll
                for (int nt = 0; nt < 8; nt++) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
        }
        if (cs.i == cs.outer()) {  // entering the slab's own k-range: everything so far is an off-diagonal block
#pragma unroll
            for (int mt = 0; mt < 4; mt++)
#pragma unroll
                for (int nt = 0; nt < 8; nt++) {
                    acc[mt][nt][0] *= 2.0;
                    acc[mt][nt][1] *= 2.0;
                }
        }
        cp_async_wait<kStages - 2>();
        __syncthreads();
        if (it + kStages - 1 < total) {
            rho_load_stage(sm + (size_t)((it + kStages - 1) % kStages) * kRhoStageDoubles, phi, P, p0, nloc, nbp, ld.slab * kTileN, ld.chunk() * kTileK);
            ld.advance();
        }
        cp_async_commit();
        const double* st = sm + (size_t)(it % kStages) * kRhoStageDoubles;
        if (narrow)
            rho_mma_stage<4>(st, acc, wm, wn, lane);
        else
            rho_mma_stage<8>(st, acc, wm, wn, lane);
        if (cs.i == cs.count() - 1) {
            // epilogue of this column slab: rowsum += T'[p][n] * Phi[p][n]
            const int ntn = narrow ? 4 : 8;
#pragma unroll
            for (int mt = 0; mt < 4; mt++) {
                const long p = p0 + wm * 32 + mt * 8 + g;
                if (p < nloc) {
#pragma unroll
                    for (int nt = 0; nt < 8; nt++) {
                        if (nt < ntn) {
                            const int col = slab + wn * (ntn * 8) + nt * 8 + q * 2;
                            if (col < nbp) {
                                const double2 f = *reinterpret_cast<const double2*>(phi + p * (long)nbp + col);
                                rowsum[mt] = fma(acc[mt][nt][0], f.x, rowsum[mt]);
                                rowsum[mt] = fma(acc[mt][nt][1], f.y, rowsum[mt]);
                            }
                        }
                    }
                }
            }
        }
        cs.advance();
    }
    cp_async_wait<0>();
    // reduce over the 4 lanes of a quad, then over the two N-warps
#pragma unroll
    for (int mt = 0; mt < 4; mt++) {
        double v = rowsum[mt];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (q == 0) red[wn * kTileM + wm * 32 + mt * 8 + g] = v;
    }
    __syncthreads();
    if (tid < kTileM) {
        const long p = p0 + tid;
        if (p < nloc) rho[p] = 2.0 * (red[tid] + red[kTileM + tid]);
    }
}

// =========================================================================================================
// C_z = Phi^T diag(d_z) Phi  (upper-triangular 128x128 tile pairs, split over point ranges)
// =========================================================================================================
constexpr int kConStageDoubles = 2 * kTileK * kLdN + kTileK;
constexpr size_t kConSmemBytes = (size_t)kStages * kConStageDoubles * sizeof(double);

__device__ __forceinline__ void con_load_stage(double* st, const double* __restrict__ phi, const double* __restrict__ d,
                                               long pk, long pend, int nbp, int ci, int cj, bool diag) {
    double* As = st;                     // [32][136]  Phi[pk+r][ci + c]
    double* Bs = st + kTileK * kLdN;     // [32][136]  Phi[pk+r][cj + c]
    double* ds = st + 2 * kTileK * kLdN; // [32]
    const int tid = threadIdx.x;
#pragma unroll
    for (int it = 0; it < 8; it++) {
        const int ch = tid + it * kDenseThreads;
        const int r = ch >> 6, c = (ch & 63) * 2;
        const long p = pk + r;
        const bool okp = p < pend;
        const bool oki = okp && (ci + c < nbp);
        cp_async16(As + r * kLdN + c, phi + (okp ? p : 0) * (long)nbp + (oki ? ci + c : 0), oki);
        if (!diag) {
            const bool okj = okp && (cj + c < nbp);
            cp_async16(Bs + r * kLdN + c, phi + (okp ? p : 0) * (long)nbp + (okj ? cj + c : 0), okj);
        }
    }
    if (tid < kTileK / 2) {
        const long p = pk + tid * 2;
        // d is padded to an even length and zero beyond the shard, so a 16-byte copy is always in bounds
        cp_async16(ds + tid * 2, d + (p < pend ? p : 0), p < pend);
    }
}

// Warp tiling of the 128 x (128|64) output tile: 8 warps stacked along M, each owning MT = 2 row tiles (16 rows) and
// the whole width (NT = 16 column tiles, 8 for an edge tile).  Per k4-step a warp then scales only 2 A fragments by the
// point weight (the DMULs compete with DMMA for the FP64 pipe) and issues 32 DMMAs from 2 + 16 fragment loads.
template <int MT, int NT>
__device__ __forceinline__ void con_mma_stage(const double* st, bool diag, double (&acc)[32][2], int warp, int lane) {
    const double* As = st;
    const double* Bs = diag ? st : st + kTileK * kLdN;
    const double* ds = st + 2 * kTileK * kLdN;
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int kk = 0; kk < kTileK; kk += 4) {
        double a[MT], b[NT];
        const double dv = ds[kk + q];
#pragma unroll
        for (int mt = 0; mt < MT; mt++) a[mt] = As[(kk + q) * kLdN + warp * (MT * 8) + mt * 8 + g] * dv;
#pragma unroll
        for (int nt = 0; nt < NT; nt++) b[nt] = Bs[(kk + q) * kLdN + nt * 8 + g];
#pragma unroll
        for (int mt = 0; mt < MT; mt++)
#pragma unroll
            for (int nt = 0; nt < NT; nt++) dmma884(acc[mt * NT + nt][0], acc[mt * NT + nt][1], a[mt], b[nt]);
    }
}

// Diagonal tile pair (ti == tj): of the four 64x64 quadrants only (0,0), (0,1) and (1,1) are needed.  Warp w owns row
// tile w of the upper half (against all 16 column tiles) and row tile 8+w of the lower half (against column tiles
// 8..15): 24 instead of 32 DMMAs per k4-step with a fully static register layout.  For a 64-wide edge tile (NT = 8)
// only the upper-left quadrant exists.  The skipped quadrant stays zero; k_contract_reduce never reads it.
template <int NT>
__device__ __forceinline__ void con_mma_stage_diag(const double* st, double (&acc)[32][2], int warp, int lane) {
    const double* As = st;
    const double* ds = st + 2 * kTileK * kLdN;
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int kk = 0; kk < kTileK; kk += 4) {
        double b[NT];
        const double dv = ds[kk + q];
        const double a0 = As[(kk + q) * kLdN + warp * 8 + g] * dv;
#pragma unroll
        for (int nt = 0; nt < NT; nt++) b[nt] = As[(kk + q) * kLdN + nt * 8 + g];
#pragma unroll
        for (int nt = 0; nt < NT; nt++) dmma884(acc[nt][0], acc[nt][1], a0, b[nt]);
        if (NT == 16) {
            const double a1 = As[(kk + q) * kLdN + 64 + warp * 8 + g] * dv;
#pragma unroll
            for (int nt = 8; nt < NT; nt++) dmma884(acc[8 + nt][0], acc[8 + nt][1], a1, b[nt]);
        }
    }
}

template <int NT>
__device__ __forceinline__ void con_store_diag(double* out, const double (&acc)[32][2], int warp, int lane) {
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int nt = 0; nt < NT; nt++)
        *reinterpret_cast<double2*>(out + (warp * 8 + g) * kTileN + nt * 8 + q * 2) = make_double2(acc[nt][0], acc[nt][1]);
    if (NT == 16) {
#pragma unroll
        for (int nt = 8; nt < NT; nt++)
            *reinterpret_cast<double2*>(out + (64 + warp * 8 + g) * kTileN + nt * 8 + q * 2) = make_double2(acc[8 + nt][0], acc[8 + nt][1]);
    }
}

template <int MT, int NT>
__device__ __forceinline__ void con_store(double* out, const double (&acc)[32][2], int warp, int lane) {
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int mt = 0; mt < MT; mt++)
#pragma unroll
        for (int nt = 0; nt < NT; nt++)
            *reinterpret_cast<double2*>(out + (warp * (MT * 8) + mt * 8 + g) * kTileN + nt * 8 + q * 2) =
                make_double2(acc[mt * NT + nt][0], acc[mt * NT + nt][1]);
}

// Work decomposition ("stream-K"): the (matrix z, tile pair) items, each nchunk k-chunks long and weighted by their
// DMMA cost (a 64-wide edge tile costs half), are laid end to end and cut into one equal share per CTA (one CTA per
// SM).  A CTA therefore executes 1-3 segments = (z, pair, [c_begin, c_end)) and writes one partial tile per segment;
// k_contract_reduce adds an item's partial tiles in a fixed order.  The schedule is built on the host (dftgrid_api.cu).
struct ConSeg {
    int z, pair, c_begin, c_end;
};

// grid = number of CTAs in the schedule.  d0/d1: per-point weights of matrix 0/1 (zero-padded past the shard).
// partial: [nseg][128*128].
__global__ void __launch_bounds__(kDenseThreads, 1)
k_contract(const double* __restrict__ phi, const double* __restrict__ d0, const double* __restrict__ d1,
           const int* __restrict__ pair_ij, const ConSeg* __restrict__ segs, const int* __restrict__ cta_seg_off,
           double* __restrict__ partial, long nloc, int nbp) {
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int s_begin = cta_seg_off[blockIdx.x], s_end = cta_seg_off[blockIdx.x + 1];
    for (int sidx = s_begin; sidx < s_end; sidx++) {
        const ConSeg sg = segs[sidx];
        const int ti = pair_ij[2 * sg.pair], tj = pair_ij[2 * sg.pair + 1];
        const int ci = ti * kTileM, cj = tj * kTileN;
        const bool diag = ti == tj;
        const double* d = sg.z == 0 ? d0 : d1;
        const long c_begin = sg.c_begin;
        const int total = sg.c_end - sg.c_begin;
        const bool narrow = min(kTileN, nbp - cj) <= 64;

        double acc[32][2];
#pragma unroll
        for (int t = 0; t < 32; t++) acc[t][0] = acc[t][1] = 0.0;

        for (int s = 0; s < kStages - 1; s++) {
            if (s < total) con_load_stage(sm + (size_t)s * kConStageDoubles, phi, d, (c_begin + s) * kTileK, nloc, nbp, ci, cj, diag);
            cp_async_commit();
        }
        for (int it = 0; it < total; it++) {
            cp_async_wait<kStages - 2>();
            __syncthreads();
            {
                const int nx = it + kStages - 1;
#ifndef DFG_ABLATE_LOADS
                if (nx < total)
                    con_load_stage(sm + (size_t)(nx % kStages) * kConStageDoubles, phi, d, (c_begin + nx) * kTileK, nloc, nbp, ci, cj, diag);
#endif
                cp_async_commit();
            }
            const double* st = sm + (size_t)(it % kStages) * kConStageDoubles;
            if (diag) {
                if (narrow)
                    con_mma_stage_diag<8>(st, acc, warp, lane);
                else
                    con_mma_stage_diag<16>(st, acc, warp, lane);
            } else if (narrow) {
                con_mma_stage<2, 8>(st, false, acc, warp, lane);
            } else {
                con_mma_stage<2, 16>(st, false, acc, warp, lane);
            }
        }
        cp_async_wait<0>();
        __syncthreads();  // every warp is done with the stage buffers before the next segment refills them
        double* out = partial + (size_t)sidx * (size_t)(kTileM * kTileN);
        if (diag) {
            if (narrow)
                con_store_diag<8>(out, acc, warp, lane);
            else
                con_store_diag<16>(out, acc, warp, lane);
        } else if (narrow) {
            con_store<2, 8>(out, acc, warp, lane);
        } else {
            con_store<2, 16>(out, acc, warp, lane);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// TMA-fed variant of k_contract: a ninth (producer) warp streams the Phi rows of every stage into shared memory with
// bulk asynchronous copies (cp.async.bulk -> SASS UBLKCP, one 1 KB row per lane) that signal an mbarrier; the eight
// DMMA warps never touch global memory, wait on the "full" barrier of a stage and release it through an "empty"
// barrier — no __syncthreads() and no per-thread cp.async / address arithmetic in the tensor loop, and the pipeline
// keeps running across segment boundaries.  Same arithmetic, same partial-tile layout as k_contract.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

constexpr int kConTmaThreads = kDenseThreads + 32;
constexpr size_t kConTmaSmemBytes = (size_t)kStages * kConStageDoubles * sizeof(double) + 2 * kStages * sizeof(unsigned long long);

// phi must be readable for whole 32-row chunks (rows past nloc are zero-filled by the host side), d0/d1 likewise.
__global__ void __launch_bounds__(kConTmaThreads, 1)
k_contract_tma(const double* __restrict__ phi, const double* __restrict__ d0, const double* __restrict__ d1,
               const int* __restrict__ pair_ij, const ConSeg* __restrict__ segs, const int* __restrict__ cta_seg_off,
               double* __restrict__ partial, int nbp) {
    extern __shared__ __align__(128) double sm[];
    unsigned long long* full = reinterpret_cast<unsigned long long*>(sm + (size_t)kStages * kConStageDoubles);
    unsigned long long* empty = full + kStages;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < kStages; s++) {
            mbar_init(full + s, 1);   // the producer's arrive.expect_tx
            mbar_init(empty + s, 8);  // one arrival per DMMA warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    const int s_begin = cta_seg_off[blockIdx.x], s_end = cta_seg_off[blockIdx.x + 1];
    unsigned n = 0;  // running stage counter, continues across segments
    if (warp == 8) {
        // ===== producer warp =====
        for (int sidx = s_begin; sidx < s_end; sidx++) {
            const ConSeg sg = segs[sidx];
            const int ti = pair_ij[2 * sg.pair], tj = pair_ij[2 * sg.pair + 1];
            const int ci = ti * kTileM, cj = tj * kTileN;
            const bool diag = ti == tj;
            const double* d = sg.z == 0 ? d0 : d1;
            const unsigned wi = (unsigned)min(kTileM, nbp - ci) * 8u, wj = (unsigned)min(kTileN, nbp - cj) * 8u;  // valid row bytes
            const unsigned bytes = kTileK * (wi + (diag ? 0u : wj)) + kTileK * 8u;
            for (int c = sg.c_begin; c < sg.c_end; c++, n++) {
                const unsigned stage = n % kStages, round = n / kStages;
                double* st = sm + (size_t)stage * kConStageDoubles;
                mbar_wait(empty + stage, (round & 1u) ^ 1u);
                if (lane == 0) mbar_arrive_expect_tx(full + stage, bytes);
                __syncwarp();
                const double* row = phi + ((size_t)c * kTileK + lane) * (size_t)nbp;
                bulk_copy_g2s(st + lane * kLdN, row + ci, wi, full + stage);
                if (!diag) bulk_copy_g2s(st + kTileK * kLdN + lane * kLdN, row + cj, wj, full + stage);
                if (lane == 0) bulk_copy_g2s(st + 2 * kTileK * kLdN, d + (size_t)c * kTileK, kTileK * 8u, full + stage);
            }
        }
        return;
    }
    // ===== DMMA warps =====
    for (int sidx = s_begin; sidx < s_end; sidx++) {
        const ConSeg sg = segs[sidx];
        const int ti = pair_ij[2 * sg.pair], tj = pair_ij[2 * sg.pair + 1];
        const bool diag = ti == tj;
        const bool narrow = min(kTileN, nbp - tj * kTileN) <= 64;
        double acc[32][2];
#pragma unroll
        for (int t = 0; t < 32; t++) acc[t][0] = acc[t][1] = 0.0;
        for (int c = sg.c_begin; c < sg.c_end; c++, n++) {
            const unsigned stage = n % kStages, round = n / kStages;
            const double* st = sm + (size_t)stage * kConStageDoubles;
            mbar_wait(full + stage, round & 1u);
            if (diag) {
                if (narrow)
                    con_mma_stage_diag<8>(st, acc, warp, lane);
                else
                    con_mma_stage_diag<16>(st, acc, warp, lane);
            } else if (narrow) {
                con_mma_stage<2, 8>(st, false, acc, warp, lane);
            } else {
                con_mma_stage<2, 16>(st, false, acc, warp, lane);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + stage);
        }
        double* out = partial + (size_t)sidx * (size_t)(kTileM * kTileN);
        if (diag) {
            if (narrow)
                con_store_diag<8>(out, acc, warp, lane);
            else
                con_store_diag<16>(out, acc, warp, lane);
        } else if (narrow) {
            con_store<2, 8>(out, acc, warp, lane);
        } else {
            con_store<2, 16>(out, acc, warp, lane);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// TMA-fed variant of k_rho (same schedule and arithmetic as k_rho): the producer warp copies, per stage, the 128 Phi
// row pieces (256 B each, four per lane) and the 32 rows of the P block with cp.async.bulk into the padded tiles; the
// eight DMMA warps synchronise through mbarriers only.  phi must be readable for whole 128-row tiles.
constexpr int kRhoTmaThreads = kDenseThreads + 32;
constexpr size_t kRhoTmaSmemBytes = (size_t)kStages * kRhoStageDoubles * sizeof(double) + 2 * kTileM * sizeof(double) + 2 * kStages * sizeof(unsigned long long);

// 9 warps: one SM sub-partition hosts 3 of them, so ptxas caps the kernel at 168 registers per thread
__global__ void __launch_bounds__(kRhoTmaThreads, 1)
k_rho_tma(const double* __restrict__ phi, const double* __restrict__ P, double* __restrict__ rho, long nloc, int nbp) {
    extern __shared__ __align__(128) double sm[];
    double* red = sm + (size_t)kStages * kRhoStageDoubles;  // [2][128]
    unsigned long long* full = reinterpret_cast<unsigned long long*>(red + 2 * kTileM);
    unsigned long long* empty = full + kStages;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < kStages; s++) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    const long p0 = (long)blockIdx.x * kTileM;
    const int nk = nbp / kTileK;
    const int nslab = (nbp + kTileN - 1) / kTileN;
    int total = 0;
    for (int J = 0; J < nslab; J++) total += nk - J * (kTileN / kTileK);

    if (warp == 8) {
        // ===== producer warp =====
        RhoStep ld{0, 0, nbp};
        for (int it = 0; it < total; it++) {
            const unsigned stage = (unsigned)it % kStages, round = (unsigned)it / kStages;
            double* st = sm + (size_t)stage * kRhoStageDoubles;
            const int slab = ld.slab * kTileN, kc = ld.chunk() * kTileK;
            const unsigned wb = (unsigned)min(kTileN, nbp - slab) * 8u;
            mbar_wait(empty + stage, (round & 1u) ^ 1u);
            if (lane == 0) mbar_arrive_expect_tx(full + stage, kTileM * kTileK * 8u + kTileK * wb);
            __syncwarp();
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int row = lane + 32 * r;
                bulk_copy_g2s(st + row * kLdK, phi + (size_t)(p0 + row) * nbp + kc, kTileK * 8u, full + stage);
            }
            bulk_copy_g2s(st + kTileM * kLdK + lane * kLdN, P + (size_t)(kc + lane) * nbp + slab, wb, full + stage);
            ld.advance();
        }
        return;
    }
    // ===== DMMA warps =====
    const int wm = warp & 3, wn = warp >> 2;
    const int g = lane >> 2, q = lane & 3;
    double rowsum[4] = {0.0, 0.0, 0.0, 0.0};
    double acc[4][8][2];
    RhoStep cs{0, 0, nbp};
    for (int it = 0; it < total; it++) {
        const unsigned stage = (unsigned)it % kStages, round = (unsigned)it / kStages;
        const int slab = cs.slab * kTileN;
        const int ncols = min(kTileN, nbp - slab);
        const bool narrow = ncols <= 64;
        if (cs.i == 0) {
#pragma unroll
            for (int mt = 0; mt < 4; mt++)
#pragma unroll
                for (int nt = 0; nt < 8; nt++) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
        }
        if (cs.i == cs.outer()) {
#pragma unroll
            for (int mt = 0; mt < 4; mt++)
#pragma unroll
                for (int nt = 0; nt < 8; nt++) {
                    acc[mt][nt][0] *= 2.0;
                    acc[mt][nt][1] *= 2.0;
                }
        }
        mbar_wait(full + stage, round & 1u);
        const double* st = sm + (size_t)stage * kRhoStageDoubles;
        if (narrow)
            rho_mma_stage<4>(st, acc, wm, wn, lane);
        else
            rho_mma_stage<8>(st, acc, wm, wn, lane);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + stage);
        if (cs.i == cs.count() - 1) {
            const int ntn = narrow ? 4 : 8;
#pragma unroll
            for (int mt = 0; mt < 4; mt++) {
                const long p = p0 + wm * 32 + mt * 8 + g;
                if (p < nloc) {
#pragma unroll
                    for (int nt = 0; nt < 8; nt++) {
                        if (nt < ntn) {
                            const int col = slab + wn * (ntn * 8) + nt * 8 + q * 2;
                            if (col < nbp) {
                                const double2 f = *reinterpret_cast<const double2*>(phi + p * (long)nbp + col);
                                rowsum[mt] = fma(acc[mt][nt][0], f.x, rowsum[mt]);
                                rowsum[mt] = fma(acc[mt][nt][1], f.y, rowsum[mt]);
                            }
                        }
                    }
                }
            }
        }
        cs.advance();
    }
#pragma unroll
    for (int mt = 0; mt < 4; mt++) {
        double v = rowsum[mt];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (q == 0) red[wn * kTileM + wm * 32 + mt * 8 + g] = v;
    }
    // named barrier over the 256 DMMA threads only (the producer warp has already left)
    asm volatile("bar.sync 1, 256;\n" ::: "memory");
    if (tid < kTileM) {
        const long p = p0 + tid;
        if (p < nloc) rho[p] = 2.0 * (red[tid] + red[kTileM + tid]);
    }
}

// out0/out1: nb x nb (symmetric => row/column-major agnostic).  Each thread sums one element over the item's partial
// tiles [item_slot_off[item], item_slot_off[item+1]) in order.  grid = (npairs, 2).
__global__ void k_contract_reduce(const double* __restrict__ partial, const int* __restrict__ pair_ij, const int* __restrict__ item_slot_off,
                                  int npairs, int nb, int nbp, double scale0, double scale1, double* __restrict__ out0, double* __restrict__ out1) {
    const int pair = blockIdx.x, z = blockIdx.y;
    const int ti = pair_ij[2 * pair], tj = pair_ij[2 * pair + 1];
    const int item = z * npairs + pair;
    const int k0 = item_slot_off[item], k1 = item_slot_off[item + 1];
    double* out = z == 0 ? out0 : out1;
    const double scale = z == 0 ? scale0 : scale1;
    for (int e = threadIdx.x; e < kTileM * kTileN; e += blockDim.x) {
        const int r = e / kTileN, c = e % kTileN;
        const int gi = ti * kTileM + r, gj = tj * kTileN + c;
        if (gi >= nb || gj >= nb) continue;
        if (ti == tj && gj < gi) continue;  // lower part of a diagonal tile is the mirror image
        double s = 0.0;
        for (int k = k0; k < k1; k++) s += partial[(size_t)k * (kTileM * kTileN) + e];
        s *= scale;
        out[(size_t)gi * nb + gj] = s;
        out[(size_t)gj * nb + gi] = s;
    }
}

}  // namespace dfg
