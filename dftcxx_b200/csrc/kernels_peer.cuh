// Cross-GPU sum of the [J | XC] contraction results over peer memory (NVLink / NVSwitch P2P), replacing
// k_contract_reduce + ncclAllReduce when every rank's exchange buffer has been mapped (dftgrid_peer_connect).
//
// Every rank owns one exchange buffer (cudaMalloc, exported with cudaIpcGetMemHandle):
//   header  : ready          (u64) number of contraction epochs whose contribution is published
//             error          (u64) set when a spin-wait timed out (the host turns it into an error)
//             ticket         (u32 x 2) last-CTA election counters of the two kernels
//             consumed_by[r] (u64) number of epochs rank r has finished summing — written BY rank r INTO this buffer, so
//                            the owner polls only its own memory (also at teardown, before it frees the buffer)
//   contrib : [2][2 nb^2] doubles, indexed by epoch parity; [J | XC] of this rank's points, full matrices.
//
// k_contract_reduce_publish : fixed-order sum of the stream-K partial tiles (as k_contract_reduce) into contrib[e];
//                             the last CTA to finish publishes ready = epoch with a system-scope release.
// k_peer_sum                : waits until every rank has published the epoch, then out[i] = sum_r contrib_r[e][i] in
//                             rank order (identical bits on every rank), reading the peers' buffers directly over
//                             NVLink; the last CTA writes consumed_by[this rank] = epoch into every rank's header.
// A contribution buffer is reused two epochs later; the publisher first waits until every peer has consumed epoch-2.
// All spin-waits are bounded (kPeerSpinLimit clocks): a lost peer raises the error flag instead of hanging the GPU.
#pragma once
#include "common.cuh"
#include "kernels_dense.cuh"

namespace dfg {

constexpr int kPeerMaxRanks = 16;
constexpr long long kPeerSpinLimit = 20000000000LL;  // ~10 s at 2 GHz
constexpr size_t kPeerHeaderBytes = 256;

struct PeerHeader {
    unsigned long long ready, error;
    unsigned int ticket[2];
    unsigned long long consumed_by[kPeerMaxRanks];
};
static_assert(sizeof(PeerHeader) <= kPeerHeaderBytes, "exchange-buffer header too large");

struct PeerSet {
    int nranks, rank;
    unsigned char* base[kPeerMaxRanks];  // every rank's exchange buffer (own one included), rank order
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ PeerHeader* peer_header(const PeerSet& ps, int r) { return reinterpret_cast<PeerHeader*>(ps.base[r]); }
__device__ __forceinline__ double* peer_contrib(const PeerSet& ps, int r, unsigned long long epoch, size_t n) {
    return reinterpret_cast<double*>(ps.base[r] + kPeerHeaderBytes) + (epoch & 1ull) * n;
}

// Thread 0 of the CTA waits until every rank has published `target`: its `ready` counter (read from the peer's buffer),
// or its consumption of THIS rank's contributions (read from this rank's own header); returns false on timeout.
__device__ __forceinline__ bool peer_wait_all(const PeerSet& ps, bool consumed_field, unsigned long long target) {
    const long long t0 = clock64();
    for (int r = 0; r < ps.nranks; r++) {
        const unsigned long long* f = consumed_field ? &peer_header(ps, ps.rank)->consumed_by[r] : &peer_header(ps, r)->ready;
        while (ld_acquire_sys(f) < target) {
            if (clock64() - t0 > kPeerSpinLimit) return false;
            __nanosleep(200);
        }
    }
    return true;
}

// grid = (npairs, 2, kReduceSplit) as k_contract_reduce.
__global__ void k_contract_reduce_publish(const double* __restrict__ partial, const int* __restrict__ pair_ij, const int* __restrict__ item_slot_off,
                                          int npairs, int nb, int nbp, double scale_xc, double scale_j, PeerSet ps, unsigned long long epoch) {
    __shared__ bool ok;
    PeerHeader* me = peer_header(ps, ps.rank);
    if (threadIdx.x == 0) {
        // the buffer of this parity was last read at epoch-2: every peer must be done with it
        ok = epoch <= 2 || peer_wait_all(ps, true, epoch - 2);
        if (!ok) me->error = 1ull;
    }
    __syncthreads();
    const size_t nb2 = (size_t)nb * nb;
    double* contrib = peer_contrib(ps, ps.rank, epoch, 2 * nb2);
    const int pair = blockIdx.x, z = blockIdx.y;
    const int ti = pair_ij[2 * pair], tj = pair_ij[2 * pair + 1];
    const int item = z * npairs + pair;
    const int k0 = item_slot_off[item], k1 = item_slot_off[item + 1];
    double* out = z == 0 ? contrib + nb2 : contrib;  // res layout [J | XC]; item z = 0 is XC
    const double scale = z == 0 ? scale_xc : scale_j;
    constexpr int per_z = kTileM * kTileN / kReduceSplit;
    for (int e = blockIdx.z * per_z + threadIdx.x; e < (blockIdx.z + 1) * per_z; e += blockDim.x) {
        const int r = e / kTileN, c = e % kTileN;
        const int gi = ti * kTileM + r, gj = tj * kTileN + c;
        if (gi >= nb || gj >= nb) continue;
        if (ti == tj && gj < gi) continue;
        double s = 0.0;
        for (int k = k0; k < k1; k++) s += partial[(size_t)k * (kTileM * kTileN) + e];
        s *= scale;
        out[(size_t)gi * nb + gj] = s;
        out[(size_t)gj * nb + gi] = s;
    }
    // publish once every CTA of this grid has written its tile
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned total = gridDim.x * gridDim.y * gridDim.z;
        if (atomicAdd(&me->ticket[0], 1u) == total - 1) {
            me->ticket[0] = 0u;
            __threadfence_system();
            st_release_sys(&me->ready, epoch);
        }
    }
}

// out = sum over ranks (ascending) of contrib_r for the nmat symmetric nb x nb matrices stored back to back: only the
// upper triangles travel over NVLink, the lower ones are mirrored locally.  Any grid size; every CTA waits for the
// publications.
__global__ void k_peer_sum(PeerSet ps, unsigned long long epoch, int nb, int nmat, double* __restrict__ out) {
    __shared__ bool ok;
    PeerHeader* me = peer_header(ps, ps.rank);
    if (threadIdx.x == 0) {
        ok = peer_wait_all(ps, false, epoch);
        if (!ok) me->error = 1ull;
    }
    __syncthreads();
    if (ok) {
        const size_t nb2 = (size_t)nb * nb, n = nb2 * nmat;
        const double* src[kPeerMaxRanks];
        for (int r = 0; r < ps.nranks; r++) src[r] = peer_contrib(ps, r, epoch, n);
        for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
            const size_t m = t / nb2, e = t - m * nb2;
            const int i = (int)(e / nb), j = (int)(e - (size_t)i * nb);
            if (j < i) continue;
            double s = 0.0;
            for (int r = 0; r < ps.nranks; r++) s += __ldcg(src[r] + t);  // L2-coherent loads: never a stale L1 line
            out[t] = s;
            out[m * nb2 + (size_t)j * nb + i] = s;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(&me->ticket[1], 1u) == gridDim.x - 1) {
            me->ticket[1] = 0u;
            __threadfence_system();
            // tell every owner (this rank included) that its epoch-`epoch` contribution has been read
            for (int r = 0; r < ps.nranks; r++) st_release_sys(&peer_header(ps, r)->consumed_by[ps.rank], epoch);
        }
    }
}

}  // namespace dfg
