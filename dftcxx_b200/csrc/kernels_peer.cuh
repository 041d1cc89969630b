// Cross-GPU sum of the contraction results over peer memory (NVLink / NVSwitch P2P), replacing
// k_contract_reduce + ncclAllReduce when every rank's exchange buffer has been mapped (dftgrid_peer_connect for one
// process per GPU, cudaDeviceEnablePeerAccess inside a single-process multi-GPU handle).
//
// Every rank owns one exchange buffer (cudaMalloc; exported with cudaIpcGetMemHandle in the multi-process case):
//   header  : ready          (u64) number of contraction epochs whose contribution is published
//             error          (u64) set when a spin-wait timed out (the host turns it into an error; the handle is then dead)
//             ticket         (u32 x 2) last-CTA election counters of the two kernels
//             epoch          (u64) this rank's epoch counter, advanced on the device by k_peer_begin so that a captured
//                            CUDA graph of the iteration can be replayed (no epoch in the kernel arguments)
//             spin_limit_ns  (u64) bound of every spin-wait in nanoseconds of %globaltimer (dftgrid_peer_set_timeout)
//             consumed_by[r] (u64) number of epochs rank r has finished summing — written BY rank r INTO this buffer, so
//                            the owner polls only its own memory (also at teardown, before it frees the buffer)
//   contrib : [2][slot] doubles, indexed by epoch parity; slot = [matrix 0 | (matrix 1) | tail]: this rank's partial
//             [J | XC] (two-matrix iteration) or [F | per-shell E_J integrand sums] (fused Fock build), full matrices.
//
// k_peer_begin              : epoch += 1 (one thread).
// k_contract_reduce_publish : fixed-order sum of the stream-K partial tiles (as k_contract_reduce) into contrib[e], plus
//                             a plain copy of the tail; the last CTA to finish publishes ready = epoch (system-scope release).
// k_peer_sum                : waits until every rank has published the epoch, then out[i] = sum_r contrib_r[e][i] in
//                             rank order (identical bits on every rank), reading the peers' buffers directly over
//                             NVLink; the last CTA writes consumed_by[this rank] = epoch into every rank's header.
// A contribution buffer is reused two epochs later; the publisher first waits until every peer has consumed epoch-2.
// All spin-waits are bounded (spin_limit_ns, default 60 s of wall time): a lost peer raises the error flag instead of
// hanging the GPU; after a failed wait nothing is written or published, so the peers time out too.
#pragma once
#include "common.cuh"
#include "kernels_dense.cuh"

namespace dfg {

constexpr int kPeerMaxRanks = 16;
constexpr unsigned long long kPeerDefaultTimeoutNs = 60000000000ull;  // 60 s
constexpr size_t kPeerHeaderBytes = 256;

struct PeerHeader {
    unsigned long long ready, error;
    unsigned int ticket[2];
    unsigned long long epoch, spin_limit_ns;
    unsigned long long consumed_by[kPeerMaxRanks];
};
static_assert(sizeof(PeerHeader) <= kPeerHeaderBytes, "exchange-buffer header too large");

struct PeerSet {
    int nranks, rank;
    size_t slot;                         // doubles per epoch-parity slot of the contribution area
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
__device__ __forceinline__ double* peer_contrib(const PeerSet& ps, int r, unsigned long long epoch) {
    return reinterpret_cast<double*>(ps.base[r] + kPeerHeaderBytes) + (epoch & 1ull) * ps.slot;
}
__device__ __forceinline__ unsigned long long peer_time_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
    return t;
}

// Thread 0 of the CTA waits until every rank has published `target`: its `ready` counter (read from the peer's buffer),
// or its consumption of THIS rank's contributions (read from this rank's own header); returns false on timeout.
__device__ __forceinline__ bool peer_wait_all(const PeerSet& ps, bool consumed_field, unsigned long long target) {
    const unsigned long long t0 = peer_time_ns(), limit = peer_header(ps, ps.rank)->spin_limit_ns;
    for (int r = 0; r < ps.nranks; r++) {
        const unsigned long long* f = consumed_field ? &peer_header(ps, ps.rank)->consumed_by[r] : &peer_header(ps, r)->ready;
        while (ld_acquire_sys(f) < target) {
            if (peer_time_ns() - t0 > limit) return false;
            __nanosleep(200);
        }
    }
    return true;
}

// epoch += 1 on the device (one thread), first kernel of every contraction of a peer-connected handle.
__global__ void k_peer_begin(PeerSet ps) { peer_header(ps, ps.rank)->epoch += 1ull; }

// grid = (npairs, nz, kReduceSplit) as k_contract_reduce.  Matrix z is scaled by scale{z} and written at off{z} (doubles)
// of this epoch's slot; `ntail` plain doubles are copied from tail_src to tail_off.
__global__ void k_contract_reduce_publish(const double* __restrict__ partial, const int* __restrict__ pair_ij, const int* __restrict__ item_slot_off,
                                          int npairs, int nb, int nbp, double scale0, double scale1, size_t off0, size_t off1,
                                          const double* __restrict__ tail_src, int ntail, size_t tail_off, PeerSet ps) {
    __shared__ bool ok;
    PeerHeader* me = peer_header(ps, ps.rank);
    const unsigned long long epoch = me->epoch;
    if (threadIdx.x == 0) {
        // the buffer of this parity was last read at epoch-2: every peer must be done with it
        ok = epoch <= 2 || peer_wait_all(ps, true, epoch - 2);
        if (!ok) me->error = 1ull;
    }
    __syncthreads();
    if (ok) {
        double* contrib = peer_contrib(ps, ps.rank, epoch);
        const int pair = blockIdx.x, z = blockIdx.y;
        const int ti = pair_ij[2 * pair], tj = pair_ij[2 * pair + 1];
        const int item = z * npairs + pair;
        const int k0 = item_slot_off[item], k1 = item_slot_off[item + 1];
        double* out = contrib + (z == 0 ? off0 : off1);
        const double scale = z == 0 ? scale0 : scale1;
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
        if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)
            for (int t = threadIdx.x; t < ntail; t += blockDim.x) contrib[tail_off + t] = tail_src[t];
    }
    // publish once every CTA of this grid has written its tile (never after a failed wait: the peers then time out too)
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned total = gridDim.x * gridDim.y * gridDim.z;
        if (atomicAdd(&me->ticket[0], 1u) == total - 1) {
            me->ticket[0] = 0u;
            __threadfence_system();
            if (*reinterpret_cast<volatile unsigned long long*>(&me->error) == 0ull) st_release_sys(&me->ready, epoch);
        }
    }
}

// out = sum over ranks (ascending) of contrib_r: nmat symmetric nb x nb matrices stored back to back (only the upper
// triangles travel over NVLink, the lower ones are mirrored locally) followed by ntail plain doubles.  Any grid size;
// every CTA waits for the publications.
__global__ void k_peer_sum(PeerSet ps, int nb, int nmat, int ntail, double* __restrict__ out) {
    __shared__ bool ok;
    PeerHeader* me = peer_header(ps, ps.rank);
    const unsigned long long epoch = me->epoch;
    if (threadIdx.x == 0) {
        ok = peer_wait_all(ps, false, epoch);
        if (!ok) me->error = 1ull;
    }
    __syncthreads();
    if (ok) {
        const size_t nb2 = (size_t)nb * nb, n = nb2 * nmat;
        const double* src[kPeerMaxRanks];
        for (int r = 0; r < ps.nranks; r++) src[r] = peer_contrib(ps, r, epoch);
        for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n + (size_t)ntail; t += (size_t)gridDim.x * blockDim.x) {
            if (t < n) {
                const size_t m = t / nb2, e = t - m * nb2;
                const int i = (int)(e / nb), j = (int)(e - (size_t)i * nb);
                if (j < i) continue;
                double s = 0.0;
                for (int r = 0; r < ps.nranks; r++) s += __ldcg(src[r] + t);  // L2-coherent loads: never a stale L1 line
                out[t] = s;
                out[m * nb2 + (size_t)j * nb + i] = s;
            } else {
                double s = 0.0;
                for (int r = 0; r < ps.nranks; r++) s += __ldcg(src[r] + t);
                out[t] = s;
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(&me->ticket[1], 1u) == gridDim.x - 1) {
            me->ticket[1] = 0u;
            __threadfence_system();
            // tell every owner (this rank included) that its epoch-`epoch` contribution has been read
            if (*reinterpret_cast<volatile unsigned long long*>(&me->error) == 0ull)
                for (int r = 0; r < ps.nranks; r++) st_release_sys(&peer_header(ps, r)->consumed_by[ps.rank], epoch);
        }
    }
}

}  // namespace dfg
