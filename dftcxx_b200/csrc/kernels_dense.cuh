// Dense FP64 contractions on the tensor pipe (DMMA, mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4 on sm_100a;
// tcgen05/wgmma have no f64 kind, so warp-level DMMA fed from bulk-copied (cp.async.bulk, SASS UBLKCP) shared
// memory tiles is the Blackwell FP64 tensor path).  Measured register-resident peak on this pool's B200:
// 37.05 TFLOP/s (profiles/r01_fp64_peak_microbench.txt), identical to the DFMA peak, at a quarter of the issue slots.
//
//   k_rho_tma      : rho_p = 2 * sum_n Phi[p][n] * (sum_k Phi[p][k] P[k][n])   (src/gridpoint.cpp:82-84)
//   k_contract_tma : C_z[i][j] = sum_p Phi[p][i] d_z[p] Phi[p][j], z in {XC, J} (src/dft.cpp:424-432, src/atomicgrid.cpp:471-488)
//   k_contract_reduce : fixed-order sum of the split-K partials, mirrored into both triangles.
//   k_chunk_flags  : which 32-point chunks of Phi hold any non-zero amplitude (all-zero chunks are skipped, exactly).
//
// Both tensor kernels are warp-specialised: a ninth (producer) warp streams the tiles of every pipeline stage into
// shared memory with bulk asynchronous copies that signal an mbarrier; the eight DMMA warps never touch global memory
// in their main loop, wait on the "full" barrier of a stage and release it through an "empty" barrier.
//
// Shared-memory tiles are padded so that every fragment load is bank-conflict free: a 64-bit load of a warp is served as
// two half-warp wavefronts of 16 x 8 bytes, so the 16 addresses of a half warp (lane = 4 g + q, g = 0..3, q = 0..3) must
// fall into 16 distinct 8-byte banks, i.e. be distinct mod 16 doubles:
//   [rows][32+4]  for "row = lane/4, col = lane%4" accesses: 36 g + q  = 4 g + q (mod 16)  -> 0..15
//   [rows][128+4] for "row = lane%4, col = lane/4" accesses: 132 q + g = 4 q + g (mod 16)  -> 0..15
// (Round 1 used 128+8 for the second kind: 136 q + g = 8 q + g (mod 16) collides for q = 0 / 2 and 1 / 3 — the 2-way
// conflicts ncu counted as l1tex__data_bank_conflicts_pipe_lsu_mem_shared; rows stay 16-byte aligned for the bulk copies.)
#pragma once
#include "common.cuh"

namespace dfg {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

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

// Test hook (dftgrid_debug_set_stress): pseudo-random delays in the producer (bit 0) and / or consumer (bit 1) warps of the
// two mbarrier pipelines.  Any ordering bug between the bulk copies and the fragment loads (a stage overwritten before its
// last reader, a stage read before its bytes landed) shows up as changed bits under the perturbed timing; the results must
// stay identical to the unperturbed run (tests/test_gpu_stress.py).  Compiled only into libdftgrid_stress.so (-DDFG_STRESS):
// even an untaken branch on a constant costs the contraction 3-5 % (measured 11.2 -> 11.6 ms), so the product has none.
__constant__ int c_stress_mode = 0;
__device__ __forceinline__ void stress_delay(int bit, unsigned n) {
#ifdef DFG_STRESS
    if (c_stress_mode & bit) {
        unsigned h = (n * 2654435761u) ^ (blockIdx.x * 40503u) ^ ((threadIdx.x >> 5) * 9176u);
        h ^= h >> 13;
        h *= 0x5bd1e995u;
        h ^= h >> 15;
        if ((h & 3u) == 0u) __nanosleep(h % 3000u);
    }
#endif
}

constexpr int kDenseThreads = 256;  // 8 DMMA warps
constexpr int kTileM = 128;
constexpr int kTileN = 128;
constexpr int kTileK = 32;
constexpr int kLdK = kTileK + 4;    // 36
#ifndef DFG_LDN_PAD
#define DFG_LDN_PAD 4  // developer A/B: 8 was round 1's value
#endif
constexpr int kLdN = kTileN + DFG_LDN_PAD;    // 132
constexpr int kStages = 3;

// Block-sparsity map of Phi (screening): mask[c] bit b = 1 when any amplitude of the 32 rows of chunk c
// in the 32-column block b exceeds tau in magnitude.  Far from a nucleus exp(-alpha r^2) underflows to exactly +0 (the
// outermost radial shells: whole chunks of exact zeros), and long before that a block's amplitudes are below any
// threshold that could move rho, J or XC: a block with max |phi| <= tau contributes at most tau * |phi_other| * |d| per
// point.  tau = 0 keeps the skipping exact; the default tau = 1e-20 (dftgrid_api.cu) is 10 orders of magnitude below the
// 1e-10 parity tolerance of J / XC.  The tensor kernels skip, per 32-point chunk, the column blocks whose bit is clear:
// whole pipeline stages when no warp needs them, otherwise per warp.  Warp per chunk; nbp / 32 <= 64 blocks.
__global__ void k_chunk_masks(const double* __restrict__ phi, long nchunk, int nbp, double tau, unsigned long long* __restrict__ mask) {
    const long c = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= nchunk) return;
    const double* base = phi + (size_t)c * kTileK * nbp;
    unsigned long long m = 0ull;
    for (int b = 0; b < nbp / 32; b++) {
        double mx = 0.0;
#pragma unroll 8
        for (int r = 0; r < kTileK; r++) mx = fmax(mx, fabs(base[(size_t)r * nbp + b * 32 + lane]));
        if (__ballot_sync(0xffffffffu, mx > tau) != 0u) m |= 1ull << b;
    }
    if (lane == 0) mask[c] = m;
}

// =========================================================================================================
// rho
// =========================================================================================================
// P is symmetric, so for the 128-column slab J of T = Phi P only the k-chunks at or beyond the slab are visited and
// the strictly-lower part is counted twice: rho = 2 sum_n phi_n (2 sum_{k > blk(n)} P_kn phi_k + sum_{k in blk(n)} P_kn phi_k)
// with blk(n) the 32-wide block of column n.  The factor is moved into the operand: Ph = P with its 32x32 diagonal
// blocks halved (exact), T'' = sum_{k >= blk(n)} Phi Ph, rho = 4 sum_n phi_n T''_n.  Inside the slab's own k-range a
// chunk therefore only feeds the column blocks at or before it (10 instead of 16 block-steps per slab).
//
// Warp tiling of the 128 x 128 slab tile: 4 warps along M (32 rows) x 2 along N; an N-warp owns two of the four
// 32-column blocks, {0,3} and {1,2}, so that both do the same number of in-slab block-steps (5 each).
constexpr int kRhoStageDoubles = kTileM * kLdK + kTileK * kLdN;
constexpr int kRhoTmaThreads = kDenseThreads + 128;  // 2 DMMA warpgroups + 1 producer warpgroup
constexpr size_t kRhoTmaSmemBytes = (size_t)kStages * kRhoStageDoubles * sizeof(double) + 2 * kTileM * sizeof(double) + 2 * kStages * sizeof(unsigned long long);

__device__ __forceinline__ int rho_block_of(int wn, int half) { return wn == 0 ? (half == 0 ? 0 : 3) : (half == 0 ? 1 : 2); }

// One k-chunk of one slab for one warp: acc[mt][half*4 + j] += Phi[rows of mt][k] * Ph[k][c_half + 8j ..]
template <bool H0, bool H1>
__device__ __forceinline__ void rho_mma_stage(const double* st, double (&acc)[4][8][2], int wm, int c0, int c1, int lane) {
    const double* As = st;
    const double* Bs = st + kTileM * kLdK;
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int kk = 0; kk < kTileK; kk += 4) {
        double a[4], b[4];
#pragma unroll
        for (int mt = 0; mt < 4; mt++) a[mt] = As[(wm * 32 + mt * 8 + g) * kLdK + kk + q];
        if (H0) {
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = Bs[(kk + q) * kLdN + c0 + j * 8 + g];
#pragma unroll
            for (int mt = 0; mt < 4; mt++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma884(acc[mt][j][0], acc[mt][j][1], a[mt], b[j]);
        }
        if (H1) {
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = Bs[(kk + q) * kLdN + c1 + j * 8 + g];
#pragma unroll
            for (int mt = 0; mt < 4; mt++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma884(acc[mt][4 + j][0], acc[mt][4 + j][1], a[mt], b[j]);
        }
    }
}

// Steps of one CTA: for slab J = 0.., first the k-chunks past the slab (ascending), then the slab's own chunks in
// DESCENDING order: own chunk `rel` feeds the column blocks <= rel, so after that step block `rel` is complete, and the
// chunk's Phi tile in shared memory holds exactly the amplitudes of that block's columns for the row-dot epilogue.
// Screening: a step (slab, k-chunk kc) is dropped for the whole CTA when no row group of the tile needs it — a row group
// needs it when its Phi block kc is significant (A operand) and it has a significant block inside the slab (it will use
// T there); see k_chunk_masks.  Producer and DMMA warps evaluate the same predicate on the same four masks.
struct RhoMasks {
    unsigned long long m[4];
    bool all;  // no map (more than 64 column blocks): every step is kept
    __device__ __forceinline__ bool keep(int slab, int kc) const {
        if (all) return true;
        bool k = false;
#pragma unroll
        for (int r = 0; r < 4; r++) k = k || (((m[r] >> kc) & 1ull) && ((m[r] >> (4 * slab)) & 0xFull));
        return k;
    }
};

struct RhoStep {
    int slab, i, nbp, stride;  // a CTA visits the slabs slab, slab + stride, ... (stride = number of CTAs sharing a tile)
    __device__ __forceinline__ int nk() const { return nbp / kTileK; }
    __device__ __forceinline__ int first_chunk() const { return slab * (kTileN / kTileK); }
    __device__ __forceinline__ int count() const { return nk() - first_chunk(); }
    __device__ __forceinline__ int nblk() const { return min(kTileN / kTileK, count()); }  // 32-column blocks in the slab
    __device__ __forceinline__ int outer() const { return count() - nblk(); }
    // own-chunk index inside the slab (nblk-1 .. 0), or -1 for a chunk past the slab
    __device__ __forceinline__ int rel() const { return i < outer() ? -1 : nblk() - 1 - (i - outer()); }
    __device__ __forceinline__ int chunk() const { return i < outer() ? first_chunk() + nblk() + i : first_chunk() + rel(); }
    __device__ __forceinline__ void advance() {
        if (++i == count()) {
            i = 0;
            slab += stride;
        }
    }
    __device__ __forceinline__ bool valid(int nslab) const { return slab < nslab; }
    // move to the first kept step at or after the current one
    __device__ __forceinline__ void seek(const RhoMasks& M, int nslab) {
        while (valid(nslab) && !M.keep(slab, chunk())) advance();
    }
};

// grid = (ceil(number of non-zero 32-point chunks / 4), nsplit): a CTA's 128-row tile is made of four non-zero chunks
// (chunk_ids, padded to a multiple of 4 with -1 = unused row group).  Ph: zero-padded [nbp][nbp] density matrix with
// halved 32x32 diagonal blocks (k_pad_P).  rho of skipped chunks stays 0.  With gridDim.y = nsplit > 1 the column slabs
// of a tile are dealt round-robin to nsplit CTAs (finer work items when a rank holds only a few waves of tiles); CTA y
// then writes its partial density to out + y * part_stride and k_rho_combine adds the parts in order.
// 3 warpgroups: two of DMMA warps, one whose first warp is the producer.  384 threads start with 168 registers each;
// the producer warpgroup hands its share back (setmaxnreg.dec) and the DMMA warpgroups grow to 232 (setmaxnreg.inc), so
// the 128-register accumulator tile plus fragments and loop state never spill.
__global__ void __launch_bounds__(kRhoTmaThreads, 1)
k_rho_tma(const double* __restrict__ phi, const double* __restrict__ Ph, const int* __restrict__ chunk_ids,
          const unsigned long long* __restrict__ chunk_mask, double* __restrict__ out, long part_stride, long nloc, int nbp) {
    extern __shared__ __align__(128) double sm[];
    double* red = sm + (size_t)kStages * kRhoStageDoubles;  // [2][128]
    unsigned long long* full = reinterpret_cast<unsigned long long*>(red + 2 * kTileM);
    unsigned long long* empty = full + kStages;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < kStages; s++) {
            mbar_init(full + s, 4);   // one arrive.expect_tx per producer warp
            mbar_init(empty + s, 8);  // one arrival per DMMA warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    const int* my_chunks = chunk_ids + 4 * (size_t)blockIdx.x;  // row group r (32 rows) of the tile = chunk my_chunks[r]
    const int nk = nbp / kTileK;
    const int nslab = (nbp + kTileN - 1) / kTileN;
    const int sub = blockIdx.y, nsplit = gridDim.y;
    RhoMasks M;
    M.all = chunk_mask == nullptr;
#pragma unroll
    for (int r = 0; r < 4; r++) M.m[r] = (!M.all && my_chunks[r] >= 0) ? chunk_mask[4 * (size_t)blockIdx.x + r] : 0ull;
    int total = 0;  // kept steps of this CTA
    for (int J = sub; J < nslab; J += nsplit)
        for (int kc = J * (kTileN / kTileK); kc < nk; kc++) total += M.keep(J, kc) ? 1 : 0;
    double* rho = out + (size_t)sub * part_stride;

    if (warp >= 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
        // ===== producer warps: warp 8+j copies, per stage, the 32 Phi row pieces (256 B each, one per lane) of row group j
        // and 8 of the 32 rows of the Ph block.  A bulk copy costs the issuing warp ~50 clocks, so one warp alone
        // (160 copies per stage) cannot keep up with the short in-slab steps.
        const int j = warp - 8;
        RhoStep ld{sub, 0, nbp, nsplit};
        const double* grp = phi + ((size_t)max(my_chunks[j], 0) * kTileK + lane) * (size_t)nbp;  // unused group: any valid rows
        for (int it = 0; it < total; it++) {
            ld.seek(M, nslab);
            const unsigned stage = (unsigned)it % kStages, round = (unsigned)it / kStages;
            double* st = sm + (size_t)stage * kRhoStageDoubles;
            const int slab = ld.slab * kTileN, kc = ld.chunk() * kTileK;
            const unsigned wb = (unsigned)min(kTileN, nbp - slab) * 8u;
            stress_delay(1, (unsigned)it);
            mbar_wait(empty + stage, (round & 1u) ^ 1u);
            if (lane == 0) mbar_arrive_expect_tx(full + stage, 32u * kTileK * 8u + 8u * wb);
            __syncwarp();
            bulk_copy_g2s(st + (32 * j + lane) * kLdK, grp + kc, kTileK * 8u, full + stage);
            if (lane < 8) {
                const int r = 8 * j + lane;
                bulk_copy_g2s(st + kTileM * kLdK + r * kLdN, Ph + (size_t)(kc + r) * nbp + slab, wb, full + stage);
            }
            ld.advance();
        }
        return;
    }
    // ===== DMMA warps =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");
    const int wm = warp & 3, wn = warp >> 2;
    const int g = lane >> 2, q = lane & 3;
    const int blk0 = rho_block_of(wn, 0), blk1 = rho_block_of(wn, 1);
    const int my_chunk = my_chunks[wm];  // the warp's 32 rows are one row group
    const long p0w = (long)max(my_chunk, 0) * kTileK;
    double rowsum[4] = {0.0, 0.0, 0.0, 0.0};
    double acc[4][8][2];
    RhoStep cs{sub, 0, nbp, nsplit};
    // significant column blocks of this warp's 32 rows (selects, not a dynamically indexed register array)
    const unsigned long long mymask = M.all ? ~0ull : (wm == 0 ? M.m[0] : wm == 1 ? M.m[1] : wm == 2 ? M.m[2] : M.m[3]);
    auto mybit = [&](int b) { return M.all || ((mymask >> b) & 1ull); };
    int cur_slab = -1;
    for (int it = 0; it < total; it++) {
        cs.seek(M, nslab);
        const unsigned stage = (unsigned)it % kStages, round = (unsigned)it / kStages;
        const int nblk = cs.nblk();
        if (cs.slab != cur_slab) {  // first kept step of a slab
            cur_slab = cs.slab;
#pragma unroll
            for (int mt = 0; mt < 4; mt++)
#pragma unroll
                for (int nt = 0; nt < 8; nt++) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
        }
        // a chunk past the slab feeds every column block, the slab's own chunk `rel` the blocks <= rel; a warp skips the
        // step when its rows' amplitudes in k-block `chunk` are insignificant, and a column block whose amplitudes are
        // (T there is only ever multiplied by them)
        const int rel = cs.rel();
        const bool ka = mybit(cs.chunk());
        const int sb = cs.slab * (kTileN / kTileK);
        const bool h0 = ka && blk0 < nblk && (rel < 0 || blk0 <= rel) && mybit(sb + blk0);
        const bool h1 = ka && blk1 < nblk && (rel < 0 || blk1 <= rel) && mybit(sb + blk1);
        mbar_wait(full + stage, round & 1u);
        stress_delay(2, (unsigned)it);
        const double* st = sm + (size_t)stage * kRhoStageDoubles;
        if (h0 && h1)
            rho_mma_stage<true, true>(st, acc, wm, blk0 * 32, blk1 * 32, lane);
        else if (h0)
            rho_mma_stage<true, false>(st, acc, wm, blk0 * 32, blk1 * 32, lane);
        else if (h1)
            rho_mma_stage<false, true>(st, acc, wm, blk0 * 32, blk1 * 32, lane);
        if (rel >= 0 && (rel == blk0 || rel == blk1) && mybit(sb + rel)) {
            // block `rel` is complete: rowsum += T''[p][n] * Phi[p][n] with Phi[p][slab + 32 rel ..] = this stage's Phi tile
            const double* As = st + (wm * 32 + g) * kLdK + q * 2;
            if (rel == blk0) {
#pragma unroll
                for (int mt = 0; mt < 4; mt++)
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const double2 f = *reinterpret_cast<const double2*>(As + mt * 8 * kLdK + j * 8);
                        rowsum[mt] = fma(acc[mt][j][0], f.x, rowsum[mt]);
                        rowsum[mt] = fma(acc[mt][j][1], f.y, rowsum[mt]);
                    }
            } else {
#pragma unroll
                for (int mt = 0; mt < 4; mt++)
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const double2 f = *reinterpret_cast<const double2*>(As + mt * 8 * kLdK + j * 8);
                        rowsum[mt] = fma(acc[mt][4 + j][0], f.x, rowsum[mt]);
                        rowsum[mt] = fma(acc[mt][4 + j][1], f.y, rowsum[mt]);
                    }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + stage);
        cs.advance();
    }
    // reduce over the 4 lanes of a quad, then over the two N-warps
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
        const int c = my_chunks[tid >> 5];
        const long p = (long)c * kTileK + (tid & 31);
        if (c >= 0 && p < nloc) rho[p] = 4.0 * (red[tid] + red[kTileM + tid]);
    }
}

// rho = part_0 + part_1 (+ part_2): the partial densities of a tile's nsplit CTAs, added in a fixed order.
__global__ void k_rho_combine(const double* __restrict__ part, long part_stride, int nsplit, long nloc, double* __restrict__ rho) {
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nloc) return;
    double v = part[p];
    for (int s = 1; s < nsplit; s++) v += part[(size_t)s * part_stride + p];
    rho[p] = v;
}

// =========================================================================================================
// C_z = Phi^T diag(d_z) Phi  (upper-triangular 128x128 tile pairs, split over point ranges)
// =========================================================================================================
constexpr int kConMaskOff = 2 * kTileK * kLdN + kTileK;  // 8-byte slot after the weights: the staged chunk's block map
constexpr int kConFlagOff = kConMaskOff + 1;              // 8-byte slot: kConLast | kConEmpty
constexpr int kConStageDoubles = kConMaskOff + 2;         // (every stage stays 16-byte aligned for the bulk copies)
// Stage flags.  The producers mark the last stage of every (block, segment) group, so the DMMA warps need no chunk count
// of their own (counting meant a scan of the ownership hash and the screening map per group, ~10 us of pipeline stall at
// every switch of a multi-segment CTA); a group without any owned chunk is a data-less marker stage.
constexpr unsigned long long kConLast = 1ull, kConEmpty = 2ull;

// Warp tiling of the 128 x (128|64) output tile: 8 warps stacked along M, each owning MT = 2 row tiles (16 rows) and
// the whole width (NT = 16 column tiles, 8 for an edge tile).  Per k4-step a warp then scales only 2 A fragments by the
// point weight (the DMULs compete with DMMA for the FP64 pipe) and issues 32 DMMAs from 2 + 16 fragment loads.
template <int MT, int NT>
__device__ __forceinline__ void con_mma_stage(const double* st, double (&acc)[32][2], int warp, int lane) {
    const double* As = st;
    const double* Bs = st + kTileK * kLdN;
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


// The same when some of the tile's 32-column blocks are insignificant for this chunk (screening, k_chunk_masks): bbits bit
// g = column tiles 4g .. 4g+3 are needed.  Warp-uniform branches around groups of static DMMAs.
template <int MT, int NT>
__device__ __forceinline__ void con_mma_stage_masked(const double* st, double (&acc)[32][2], int warp, int lane, unsigned bbits) {
    const double* As = st;
    const double* Bs = st + kTileK * kLdN;
    const double* ds = st + 2 * kTileK * kLdN;
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int kk = 0; kk < kTileK; kk += 4) {
        double a[MT];
        const double dv = ds[kk + q];
#pragma unroll
        for (int mt = 0; mt < MT; mt++) a[mt] = As[(kk + q) * kLdN + warp * (MT * 8) + mt * 8 + g] * dv;
#pragma unroll
        for (int grp = 0; grp < NT / 4; grp++)
            if ((bbits >> grp) & 1u) {
                double b[4];
#pragma unroll
                for (int t = 0; t < 4; t++) b[t] = Bs[(kk + q) * kLdN + (grp * 4 + t) * 8 + g];
#pragma unroll
                for (int mt = 0; mt < MT; mt++)
#pragma unroll
                    for (int t = 0; t < 4; t++) dmma884(acc[mt * NT + grp * 4 + t][0], acc[mt * NT + grp * 4 + t][1], a[mt], b[t]);
            }
    }
}

// Diagonal tile pair (ti == tj), full 128 wide: only the 8x8 DMMA tiles on or above the diagonal are needed.  Warp W
// owns row tile W (against column tiles W..15) and row tile 15-W (against column tiles 15-W..15): 17 DMMAs per
// k4-step for every warp (136 of the 256 tiles) with a static register layout per W:
//   acc[c - W]                 row tile W,    column tile c = W..15
//   acc[16 - W + c - (15 - W)] row tile 15-W, column tile c = 15-W..15
// The tiles below the diagonal are never written; k_contract_reduce never reads them.
template <int W>
__device__ __forceinline__ void con_mma_stage_tri(const double* st, double (&acc)[32][2], int lane) {
    const double* As = st;
    const double* ds = st + 2 * kTileK * kLdN;
    const int g = lane >> 2, q = lane & 3;
    constexpr int NB = 16 - W;  // column tiles W..15
#pragma unroll
    for (int kk = 0; kk < kTileK; kk += 4) {
        double b[NB];
        const double dv = ds[kk + q];
        const double a0 = As[(kk + q) * kLdN + W * 8 + g] * dv;
        const double a1 = As[(kk + q) * kLdN + (15 - W) * 8 + g] * dv;
#pragma unroll
        for (int c = 0; c < NB; c++) b[c] = As[(kk + q) * kLdN + (W + c) * 8 + g];
#pragma unroll
        for (int c = 0; c < NB; c++) dmma884(acc[c][0], acc[c][1], a0, b[c]);
#pragma unroll
        for (int c = 15 - W; c < 16; c++) dmma884(acc[NB + c - (15 - W)][0], acc[NB + c - (15 - W)][1], a1, b[c - W]);
    }
}


// 64-wide diagonal edge tile with only some column blocks significant (bits 0, 1)
__device__ __forceinline__ void con_mma_stage_diag_edge_half(const double* st, double (&acc)[32][2], int warp, int lane, unsigned bits) {
    const double* As = st;
    const double* ds = st + 2 * kTileK * kLdN;
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int kk = 0; kk < kTileK; kk += 4) {
        const double dv = ds[kk + q];
        const double a0 = As[(kk + q) * kLdN + warp * 8 + g] * dv;
#pragma unroll
        for (int nt = 0; nt < 8; nt++)
            if ((bits >> (nt >> 2)) & 1u) dmma884(acc[nt][0], acc[nt][1], a0, As[(kk + q) * kLdN + nt * 8 + g]);
    }
}

// 64-wide diagonal edge tile: warp w owns row tile w against the 8 column tiles (the tile is 1 of ~28 pairs).
__device__ __forceinline__ void con_mma_stage_diag_edge(const double* st, double (&acc)[32][2], int warp, int lane) {
    const double* As = st;
    const double* ds = st + 2 * kTileK * kLdN;
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int kk = 0; kk < kTileK; kk += 4) {
        double b[8];
        const double dv = ds[kk + q];
        const double a0 = As[(kk + q) * kLdN + warp * 8 + g] * dv;
#pragma unroll
        for (int nt = 0; nt < 8; nt++) b[nt] = As[(kk + q) * kLdN + nt * 8 + g];
#pragma unroll
        for (int nt = 0; nt < 8; nt++) dmma884(acc[nt][0], acc[nt][1], a0, b[nt]);
    }
}


// Work decomposition ("stream-K", L2-friendly).  For ONE chunk the (matrix z, tile pair) items, weighted by their DMMA
// cost, are laid end to end and cut into one equal share per CTA (one CTA per SM); a share is 1-3 segments
// (z, pair, [tb, te)) where tb/te are 31-bit fixed-point FRACTIONS of the item's chunks.  Which chunks a fraction range
// means is decided by a low-discrepancy hash: chunk position x belongs to the segment with tb <= u(x) < te,
// u(x) = (x * 2654435769 mod 2^32) >> 1 (golden-ratio sequence).  Every CTA walks its chunks in ascending x, and since
// every fraction range receives its chunks evenly spread over x, all CTAs advance through the Phi rows at the same pace:
// a row is fetched from HBM by the first CTA that needs it and its other ~13 uses (7 tile pairs x 2 matrices) hit the
// L2 a few microseconds later.  A CTA whose share crosses an item boundary (2-3 segments) alternates between its
// segments once per block of `bc` chunks (~40 MB of Phi rows, L2-resident) and parks the inactive accumulators in the
// segment's partial tile.  One partial tile per segment; k_contract_reduce adds an item's partial tiles in a fixed
// order.  The schedule is built on the host (dftgrid_api.cu).
struct ConSeg {
    int z, pair;
    unsigned tb, te;  // [tb, te) in units of 2^-31
};

// Screening (k_chunk_masks): a chunk whose amplitudes are insignificant in ALL of tile i's column blocks, or in all of
// tile j's, contributes nothing to the pair and is not staged at all: ownership = hash range AND both tiles significant.
// mi / mj: the tiles' block bits inside a chunk map; cm: the chunk's map (all ones without a map).  The chunk positions x
// are an interleaved order of the active chunks (dftgrid_api.cu), so that any window of positions mixes chunks of all atoms
// and every tile pair pays its average cost per window: the CTAs keep sweeping the Phi rows in step (L2 sharing).
__device__ __forceinline__ bool con_owns(const ConSeg& sg, int x, unsigned long long cm, unsigned long long mi, unsigned long long mj) {
    const unsigned u = ((unsigned)x * 2654435769u) >> 1;
    return u >= sg.tb && u < sg.te && (cm & mi) != 0ull && (cm & mj) != 0ull;
}
__device__ __forceinline__ unsigned long long con_tile_bits(int t) { return 0xFull << (4 * t); }

// Soft lockstep of the sweep (optional): the A producer of every single-segment CTA counts the windows of `wc` chunk
// positions it has finished in prog[] and does not enter window w before `quorum` such CTAs have finished window
// w - lead, so the CTAs that share Phi rows through the L2 stay within lead + 1 windows of each other.  Waiting never
// lengthens the critical path by itself (the slowest CTA never waits); the spin is bounded (~2 ms) and falls back to free
// running, so a grid that is not fully co-resident cannot deadlock.  prog[] is zeroed before every launch.
struct ConSync {
    unsigned* prog;
    int wc, lead, quorum;
};

constexpr int kConTmaThreads = kDenseThreads + 128;  // 2 DMMA warpgroups + 1 producer warpgroup: warp 8 operand A, warp 9 operand B and the weights
constexpr size_t kConTmaSmemBytes = (size_t)kStages * kConStageDoubles * sizeof(double) + 2 * kStages * sizeof(unsigned long long);

// The chunk loop of one segment for one DMMA warp; `op` consumes one staged chunk.
// The stages of one (block, segment) group for one DMMA warp, up to and including the one flagged kConLast.
template <class Op>
__device__ __forceinline__ void con_run_group(const double* sm, unsigned long long* full, unsigned long long* empty, unsigned& n, int lane, Op&& op) {
    for (;;) {
        const unsigned stage = n % kStages, round = n / kStages;
        mbar_wait(full + stage, round & 1u);
        stress_delay(2, n);
        const double* st = sm + (size_t)stage * kConStageDoubles;
        const unsigned long long flags = *reinterpret_cast<const unsigned long long*>(st + kConFlagOff);
        if (!(flags & kConEmpty)) op(st);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + stage);
        n++;
        if (flags & kConLast) break;
    }
}

// Accumulator tile layouts of the four segment kinds: how the DMMA stage, the store to and the reload from the
// segment's partial tile address the 32 accumulator pairs.
// abits / bbits: significance of the four 32-column blocks of tile i (the A rows; warp w's rows lie in block w / 2) and of
// tile j for the staged chunk.
struct ConModeFull {
    static __device__ __forceinline__ void mma(const double* st, double (&acc)[32][2], int warp, int lane, unsigned abits, unsigned bbits) {
        if (!((abits >> (warp >> 1)) & 1u)) return;
        if (bbits == 0xFu)
            con_mma_stage<2, 16>(st, acc, warp, lane);
        else
            con_mma_stage_masked<2, 16>(st, acc, warp, lane, bbits);
    }
    template <bool LOAD>
    static __device__ __forceinline__ void io(double* out, double (&acc)[32][2], int warp, int lane) {
        const int g = lane >> 2, q = lane & 3;
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
            for (int nt = 0; nt < 16; nt++) {
                double2* p = reinterpret_cast<double2*>(out + (warp * 16 + mt * 8 + g) * kTileN + nt * 8 + q * 2);
                if (LOAD) {
                    const double2 v = *p;
                    acc[mt * 16 + nt][0] = v.x;
                    acc[mt * 16 + nt][1] = v.y;
                } else {
                    *p = make_double2(acc[mt * 16 + nt][0], acc[mt * 16 + nt][1]);
                }
            }
    }
};
struct ConModeNarrow {
    static __device__ __forceinline__ void mma(const double* st, double (&acc)[32][2], int warp, int lane, unsigned abits, unsigned bbits) {
        if (!((abits >> (warp >> 1)) & 1u)) return;
        if ((bbits & 3u) == 3u)
            con_mma_stage<2, 8>(st, acc, warp, lane);
        else
            con_mma_stage_masked<2, 8>(st, acc, warp, lane, bbits);
    }
    template <bool LOAD>
    static __device__ __forceinline__ void io(double* out, double (&acc)[32][2], int warp, int lane) {
        const int g = lane >> 2, q = lane & 3;
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
            for (int nt = 0; nt < 8; nt++) {
                double2* p = reinterpret_cast<double2*>(out + (warp * 16 + mt * 8 + g) * kTileN + nt * 8 + q * 2);
                if (LOAD) {
                    const double2 v = *p;
                    acc[mt * 8 + nt][0] = v.x;
                    acc[mt * 8 + nt][1] = v.y;
                } else {
                    *p = make_double2(acc[mt * 8 + nt][0], acc[mt * 8 + nt][1]);
                }
            }
    }
};
// 32-wide edge tile (nbp = 128 k + 32, e.g. nb = 524): 4 column tiles instead of 8
struct ConModeNarrow32 {
    static __device__ __forceinline__ void mma(const double* st, double (&acc)[32][2], int warp, int lane, unsigned abits, unsigned bbits) {
        if (!((abits >> (warp >> 1)) & 1u) || !(bbits & 1u)) return;
        con_mma_stage<2, 4>(st, acc, warp, lane);
    }
    template <bool LOAD>
    static __device__ __forceinline__ void io(double* out, double (&acc)[32][2], int warp, int lane) {
        const int g = lane >> 2, q = lane & 3;
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
            for (int nt = 0; nt < 4; nt++) {
                double2* p = reinterpret_cast<double2*>(out + (warp * 16 + mt * 8 + g) * kTileN + nt * 8 + q * 2);
                if (LOAD) {
                    const double2 v = *p;
                    acc[mt * 4 + nt][0] = v.x;
                    acc[mt * 4 + nt][1] = v.y;
                } else {
                    *p = make_double2(acc[mt * 4 + nt][0], acc[mt * 4 + nt][1]);
                }
            }
    }
};
// 32-wide diagonal edge tile: warps 0-3 own row tile w against the 4 column tiles, warps 4-7 only follow the pipeline
struct ConModeDiagEdge32 {
    static __device__ __forceinline__ void mma(const double* st, double (&acc)[32][2], int warp, int lane, unsigned abits, unsigned) {
        if (warp >= 4 || !(abits & 1u)) return;
        const double* As = st;
        const double* ds = st + 2 * kTileK * kLdN;
        const int g = lane >> 2, q = lane & 3;
#pragma unroll
        for (int kk = 0; kk < kTileK; kk += 4) {
            double b[4];
            const double dv = ds[kk + q];
            const double a0 = As[(kk + q) * kLdN + warp * 8 + g] * dv;
#pragma unroll
            for (int nt = 0; nt < 4; nt++) b[nt] = As[(kk + q) * kLdN + nt * 8 + g];
#pragma unroll
            for (int nt = 0; nt < 4; nt++) dmma884(acc[nt][0], acc[nt][1], a0, b[nt]);
        }
    }
    template <bool LOAD>
    static __device__ __forceinline__ void io(double* out, double (&acc)[32][2], int warp, int lane) {
        if (warp >= 4) return;
        const int g = lane >> 2, q = lane & 3;
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
            double2* p = reinterpret_cast<double2*>(out + (warp * 8 + g) * kTileN + nt * 8 + q * 2);
            if (LOAD) {
                const double2 v = *p;
                acc[nt][0] = v.x;
                acc[nt][1] = v.y;
            } else {
                *p = make_double2(acc[nt][0], acc[nt][1]);
            }
        }
    }
};
struct ConModeDiagEdge {
    static __device__ __forceinline__ void mma(const double* st, double (&acc)[32][2], int warp, int lane, unsigned abits, unsigned) {
        if (!((abits >> (warp >> 2)) & 1u)) return;  // row tile w lies in block w / 4
        if ((abits & 3u) == 3u)
            con_mma_stage_diag_edge(st, acc, warp, lane);
        else
            con_mma_stage_diag_edge_half(st, acc, warp, lane, abits);
    }
    template <bool LOAD>
    static __device__ __forceinline__ void io(double* out, double (&acc)[32][2], int warp, int lane) {
        const int g = lane >> 2, q = lane & 3;
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            double2* p = reinterpret_cast<double2*>(out + (warp * 8 + g) * kTileN + nt * 8 + q * 2);
            if (LOAD) {
                const double2 v = *p;
                acc[nt][0] = v.x;
                acc[nt][1] = v.y;
            } else {
                *p = make_double2(acc[nt][0], acc[nt][1]);
            }
        }
    }
};
template <int W>
struct ConModeTri {
    static __device__ __forceinline__ void mma(const double* st, double (&acc)[32][2], int, int lane, unsigned abits, unsigned) {
        // a staged diagonal chunk always runs the full triangular stage: the per-tile predicates of a masked variant
        // (17 differently conditioned DMMAs per k4-step) cost more than the skipped DMMAs save (measured: the diagonal
        // items became 2x stragglers); diagonal tiles still skip the chunks in which the whole tile is insignificant
        (void)abits;
        con_mma_stage_tri<W>(st, acc, lane);
    }
    template <bool LOAD>
    static __device__ __forceinline__ void io(double* out, double (&acc)[32][2], int, int lane) {
        const int g = lane >> 2, q = lane & 3;
        constexpr int NB = 16 - W;
#pragma unroll
        for (int c = 0; c < NB; c++) {
            double2* p = reinterpret_cast<double2*>(out + (W * 8 + g) * kTileN + (W + c) * 8 + q * 2);
            if (LOAD) {
                const double2 v = *p;
                acc[c][0] = v.x;
                acc[c][1] = v.y;
            } else {
                *p = make_double2(acc[c][0], acc[c][1]);
            }
        }
#pragma unroll
        for (int c = 15 - W; c < 16; c++) {
            double2* p = reinterpret_cast<double2*>(out + ((15 - W) * 8 + g) * kTileN + c * 8 + q * 2);
            if (LOAD) {
                const double2 v = *p;
                acc[NB + c - (15 - W)][0] = v.x;
                acc[NB + c - (15 - W)][1] = v.y;
            } else {
                *p = make_double2(acc[NB + c - (15 - W)][0], acc[NB + c - (15 - W)][1]);
            }
        }
    }
};

// Blocks [b_begin, b_end) of one segment for one DMMA warp: accumulators start from zero (`fresh`) or from the
// segment's partial tile, and are written back to it at the end.
template <class Mode>
__device__ __forceinline__ void con_segment_blocks_mode(const double* sm, unsigned long long* full, unsigned long long* empty, unsigned& n, int warp,
                                                        int lane, const ConSeg& sg, double* out, int nchunk, int bc, int b_begin, int b_end,
                                                        bool fresh, const unsigned long long* __restrict__ chunk_mask, int ti, int tj) {
    double acc[32][2];
    if (fresh) {
#pragma unroll
        for (int t = 0; t < 32; t++) acc[t][0] = acc[t][1] = 0.0;
    } else {
        Mode::template io<true>(out, acc, warp, lane);
    }
    for (int b = b_begin; b < b_end; b++)  // one group per block; the producers flag each group's last stage
        con_run_group(sm, full, empty, n, lane, [&](const double* st) {
            const unsigned long long cm = *reinterpret_cast<const unsigned long long*>(st + kConMaskOff);
            Mode::mma(st, acc, warp, lane, (unsigned)(cm >> (4 * ti)) & 0xFu, (unsigned)(cm >> (4 * tj)) & 0xFu);
        });
    Mode::template io<false>(out, acc, warp, lane);
}

__device__ __forceinline__ void con_segment_blocks(const double* sm, unsigned long long* full, unsigned long long* empty, unsigned& n, int warp, int lane,
                                                   const ConSeg sg, const int* __restrict__ pair_ij, double* out, int nbp, int nchunk, int bc,
                                                   int b_begin, int b_end, bool fresh, const unsigned long long* __restrict__ chunk_mask) {
    const int ti = pair_ij[2 * sg.pair], tj = pair_ij[2 * sg.pair + 1];
    const bool diag = ti == tj;
    const int wj = min(kTileN, nbp - tj * kTileN);
    const bool narrow = wj <= 64, narrow32 = wj <= 32;
#define DFG_SEG_ARGS sm, full, empty, n, warp, lane, sg, out, nchunk, bc, b_begin, b_end, fresh, chunk_mask, ti, tj
    if (!diag) {
        if (narrow32)
            con_segment_blocks_mode<ConModeNarrow32>(DFG_SEG_ARGS);
        else if (narrow)
            con_segment_blocks_mode<ConModeNarrow>(DFG_SEG_ARGS);
        else
            con_segment_blocks_mode<ConModeFull>(DFG_SEG_ARGS);
    } else if (narrow32) {
        con_segment_blocks_mode<ConModeDiagEdge32>(DFG_SEG_ARGS);
    } else if (narrow) {
        con_segment_blocks_mode<ConModeDiagEdge>(DFG_SEG_ARGS);
    } else {
        switch (warp) {
            case 0: con_segment_blocks_mode<ConModeTri<0>>(DFG_SEG_ARGS); break;
            case 1: con_segment_blocks_mode<ConModeTri<1>>(DFG_SEG_ARGS); break;
            case 2: con_segment_blocks_mode<ConModeTri<2>>(DFG_SEG_ARGS); break;
            case 3: con_segment_blocks_mode<ConModeTri<3>>(DFG_SEG_ARGS); break;
            case 4: con_segment_blocks_mode<ConModeTri<4>>(DFG_SEG_ARGS); break;
            case 5: con_segment_blocks_mode<ConModeTri<5>>(DFG_SEG_ARGS); break;
            case 6: con_segment_blocks_mode<ConModeTri<6>>(DFG_SEG_ARGS); break;
            default: con_segment_blocks_mode<ConModeTri<7>>(DFG_SEG_ARGS); break;
        }
    }
#undef DFG_SEG_ARGS
}

// grid = number of CTAs in the schedule.  d0/d1: per-point weights of matrix 0/1 (zero-padded past the shard).
// phi must be readable for whole 32-row chunks (rows past nloc are zero).  partial: [nseg][128*128].
// nchunk = number of non-zero chunks (length of chunk_ids), bc = chunks per L2 block.
__global__ void __launch_bounds__(kConTmaThreads, 1)
k_contract_tma(const double* __restrict__ phi, const double* __restrict__ d0, const double* __restrict__ d1, const int* __restrict__ chunk_ids,
               const unsigned long long* __restrict__ chunk_mask, const int* __restrict__ pair_ij, const ConSeg* __restrict__ segs,
               const int* __restrict__ cta_seg_off, double* __restrict__ partial, int nbp, int nchunk, int bc,
               unsigned long long* __restrict__ dbg_times = nullptr, ConSync sync = ConSync{nullptr, 1, 0, 0}) {
    extern __shared__ __align__(128) double sm[];
    unsigned long long* full = reinterpret_cast<unsigned long long*>(sm + (size_t)kStages * kConStageDoubles);
    unsigned long long* empty = full + kStages;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < kStages; s++) {
            mbar_init(full + s, 2);   // one arrive.expect_tx per producer warp
            mbar_init(empty + s, 8);  // one arrival per DMMA warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    const int s_begin = cta_seg_off[blockIdx.x], s_end = cta_seg_off[blockIdx.x + 1];
    const int nblock = bc > 0 ? (nchunk + bc - 1) / bc : 0;
    unsigned n = 0;  // running stage counter, continues across segments and blocks
    if (warp >= 8) {
        // ===== producer warps: one Phi row per lane and stage; block-major, like the consumers.  Warp 8 copies the A
        // operand (tile i's columns), warp 9 the B operand (tile j's, absent on the diagonal) and the weights: a bulk copy
        // costs its issuing warp ~50 clocks, and with one warp issuing all 65 the narrow edge tiles (32 or 64 columns,
        // a quarter or half of the DMMA work for the same number of copies) were producer-bound =====
        // 384 threads start with 168 registers each; this warpgroup hands registers back, the DMMA warpgroups grow (as in k_rho_tma)
        asm volatile("setmaxnreg.dec.sync.aligned.u32 72;\n");
        if (warp > 9) return;
        const bool prodA = warp == 8;
        bool throttle = prodA && sync.prog != nullptr && s_end - s_begin == 1;
        int cur_w = 0;
        for (int b = 0; b < nblock; b++) {
            for (int sidx = s_begin; sidx < s_end; sidx++) {
                const ConSeg sg = segs[sidx];
                const int ti = pair_ij[2 * sg.pair], tj = pair_ij[2 * sg.pair + 1];
                const int ci = ti * kTileM, cj = tj * kTileN;
                const bool diag = ti == tj;
                const double* d = sg.z == 0 ? d0 : d1;
                const unsigned wi = (unsigned)min(kTileM, nbp - ci) * 8u, wj = (unsigned)min(kTileN, nbp - cj) * 8u;  // valid row bytes
                const unsigned bytes = prodA ? kTileK * wi : kTileK * (diag ? 0u : wj) + kTileK * 8u;
                const int x0 = min(b * bc, nchunk), x1 = min((b + 1) * bc, nchunk);
                const unsigned long long mi = con_tile_bits(ti), mj = con_tile_bits(tj);
                // one stage: wait for the slot, publish map + flags, copy (a marker stage carries no data)
                auto emit = [&](size_t row0, unsigned long long cmx, unsigned long long flags) {
                    const unsigned stage = n % kStages, round = n / kStages;
                    double* st = sm + (size_t)stage * kConStageDoubles;
                    const bool data = !(flags & kConEmpty);
                    stress_delay(1, n);
                    mbar_wait(empty + stage, (round & 1u) ^ 1u);
                    if (lane == 0) {
                        if (prodA) {  // plain stores, released by the arrive below
                            *reinterpret_cast<unsigned long long*>(st + kConMaskOff) = cmx;
                            *reinterpret_cast<unsigned long long*>(st + kConFlagOff) = flags;
                        }
                        mbar_arrive_expect_tx(full + stage, data ? bytes : 0u);
                    }
                    __syncwarp();
                    if (data) {
                        const double* row = phi + (row0 + lane) * (size_t)nbp;
                        if (prodA) {
                            bulk_copy_g2s(st + lane * kLdN, row + ci, wi, full + stage);
                        } else {
                            if (!diag) bulk_copy_g2s(st + kTileK * kLdN + lane * kLdN, row + cj, wj, full + stage);
                            if (lane == 0) bulk_copy_g2s(st + 2 * kTileK * kLdN, d + row0, kTileK * 8u, full + stage);
                        }
                    }
                    n++;
                };
                // the group's owned chunks are emitted one behind the scan, so that the last one can carry kConLast
                bool have = false;
                size_t pend_row0 = 0;
                unsigned long long pend_cm = 0ull;
                for (int base = x0; base < x1; base += 32) {
                    if (throttle) {
                        const int w = base / sync.wc;
                        if (w != cur_w) {  // entering a new window: the earlier ones are done; wait for the pack
                            int gave_up = 0;
                            if (lane == 0) {
                                for (int v = cur_w; v < w; v++) atomicAdd(sync.prog + v, 1u);
                                if (w >= sync.lead) {
                                    const volatile unsigned* flag = sync.prog + (w - sync.lead);
                                    unsigned spins = 0;
                                    while (*flag < (unsigned)sync.quorum) {
                                        __nanosleep(100);
                                        if (++spins > 20000u) {  // ~2 ms: give up the lockstep, never the run
                                            gave_up = 1;
                                            break;
                                        }
                                    }
                                }
                            }
                            gave_up = __shfl_sync(0xffffffffu, gave_up, 0);
                            cur_w = w;
                            if (gave_up) {  // keep the counts moving for the others, stop waiting
                                throttle = false;
                                if (lane == 0)
                                    for (int v = w; v < (nchunk + sync.wc - 1) / sync.wc; v++) atomicAdd(sync.prog + v, 1u);
                            }
                        }
                    }
                    const int x = base + lane;
                    const unsigned long long cm = (x < x1 && chunk_mask) ? chunk_mask[x] : ~0ull;
                    unsigned mask = __ballot_sync(0xffffffffu, x < x1 && con_owns(sg, x, cm, mi, mj));
                    const int my_chunk = x < x1 ? chunk_ids[x] : 0;
                    while (mask) {
                        const int src = __ffs(mask) - 1;
                        mask &= mask - 1u;
                        const size_t row0 = (size_t)__shfl_sync(0xffffffffu, my_chunk, src) * kTileK;
                        const unsigned long long cmx = __shfl_sync(0xffffffffu, cm, src);
                        if (have) emit(pend_row0, pend_cm, 0ull);
                        pend_row0 = row0;
                        pend_cm = cmx;
                        have = true;
                    }
                }
                emit(pend_row0, pend_cm, have ? kConLast : (kConLast | kConEmpty));
            }
        }
        return;
    }
    // ===== DMMA warps =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;\n");
    unsigned long long dbg_t0 = 0;
    if (dbg_times && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(dbg_t0));
    // A CTA with a single segment keeps its accumulators in registers while the blocks go by.  A CTA whose share crosses
    // an item boundary (2-3 segments) visits its segments in turn inside every block and parks the accumulators of the
    // inactive ones in their partial tiles (L2-resident) in between.
    const int nseg = s_end - s_begin;
    if (nseg == 1) {
        con_segment_blocks(sm, full, empty, n, warp, lane, segs[s_begin], pair_ij, partial + (size_t)s_begin * (size_t)(kTileM * kTileN), nbp, nchunk, bc,
                           0, nblock, true, chunk_mask);
    } else {
        for (int b = 0; b < nblock; b++)
            for (int sidx = s_begin; sidx < s_end; sidx++)
                con_segment_blocks(sm, full, empty, n, warp, lane, segs[sidx], pair_ij, partial + (size_t)sidx * (size_t)(kTileM * kTileN), nbp, nchunk,
                                   bc, b, b + 1, b == 0, chunk_mask);
        if (nblock == 0)  // empty shard: the reduction still reads every segment's tile
            for (int sidx = s_begin; sidx < s_end; sidx++)
                con_segment_blocks(sm, full, empty, n, warp, lane, segs[sidx], pair_ij, partial + (size_t)sidx * (size_t)(kTileM * kTileN), nbp, nchunk,
                                   bc, 0, 0, true, chunk_mask);
    }
    if (dbg_times && tid == 0) {  // developer instrumentation (DFTGRID_DEBUG_CTA_TIMES): wall time of this CTA's DMMA warps
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t1));
        dbg_times[3 * blockIdx.x] = dbg_t0;
        dbg_times[3 * blockIdx.x + 1] = t1;
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;\n" : "=r"(smid));
        dbg_times[3 * blockIdx.x + 2] = (unsigned long long)n | ((unsigned long long)smid << 32);  // stages consumed | SM id
    }
}

// out0/out1: nb x nb (symmetric => row/column-major agnostic).  Each thread sums one element over the item's partial
// tiles [item_slot_off[item], item_slot_off[item+1]) in order.  grid = (npairs, 2, kReduceSplit): the tile's elements
// are split over blockIdx.z so that a small basis (one or two tile pairs) still fills the machine.
constexpr int kReduceSplit = 16;
__global__ void k_contract_reduce(const double* __restrict__ partial, const int* __restrict__ pair_ij, const int* __restrict__ item_slot_off,
                                  int npairs, int nb, int nbp, double scale0, double scale1, double* __restrict__ out0, double* __restrict__ out1) {
    const int pair = blockIdx.x, z = blockIdx.y;
    const int ti = pair_ij[2 * pair], tj = pair_ij[2 * pair + 1];
    const int item = z * npairs + pair;
    const int k0 = item_slot_off[item], k1 = item_slot_off[item + 1];
    double* out = z == 0 ? out0 : out1;
    const double scale = z == 0 ? scale0 : scale1;
    constexpr int per_z = kTileM * kTileN / kReduceSplit;
    for (int e = blockIdx.z * per_z + threadIdx.x; e < (blockIdx.z + 1) * per_z; e += blockDim.x) {
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
