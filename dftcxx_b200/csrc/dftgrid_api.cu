// C ABI (include/dftgrid.h) and host orchestration of the B200 grid engine.
// All compute runs in the hand-written sm_100a kernels of kernels_*.cuh; there is no CPU path.
#include "../../include/dftgrid.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "common.cuh"
#include "host_tables.h"
#include "kernels_dense.cuh"
#include "kernels_grid.cuh"
#include "kernels_hartree.cuh"
#include "kernels_peer.cuh"
#include "nccl_dyn.h"

using namespace dfg;

namespace {

thread_local std::string g_error;

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define CK(call)                                                                                                   \
    do {                                                                                                           \
        cudaError_t e_ = (call);                                                                                   \
        if (e_ != cudaSuccess)                                                                                     \
            throw CudaError(std::string(#call) + " failed: " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" +   \
                            std::to_string(__LINE__) + ")");                                                      \
    } while (0)

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    void alloc(size_t count) {
        release();
        n = count;
        if (count) CK(cudaMalloc(&p, count * sizeof(T)));
    }
    void upload(const std::vector<T>& v, cudaStream_t s) {
        alloc(v.size());
        if (!v.empty()) CK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    void zero(cudaStream_t s) {
        if (n) CK(cudaMemsetAsync(p, 0, n * sizeof(T), s));
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
};

}  // namespace

struct dftgrid {
    int device = 0, rank = 0, nranks = 1;
    cudaStream_t stream = nullptr;
    bool built = false, have_density = false, have_potential = false, contract_valid = false, timed_iter = false;
    long launches = 0;

    // host description
    int natoms = 0, nbf = 0, nbp = 0, nprim = 0;
    std::vector<int> Z;
    std::vector<double> atom_xyz;
    double zsum = 0.0;
    dftgrid_params prm{};
    GridShape g{};
    int leb_off = 0;

    // basis (host, column order)
    std::vector<double> center_xyz, exp_alpha, prim_coeff, prim_norm;
    std::vector<int> bf_center, bf_prim_off, center_exp_off, prim_exp, prim_lmn;

    // device: static tables
    DevBuf<double> d_atom_xyz, d_Rdist, d_rtab, d_wrad, d_leb, d_Y, d_Yt, d_pre, d_lu, d_xs;
    DevBuf<int> d_perm, d_lo, d_hi;
    DevBuf<double> d_spA, d_spCp, d_spDen, d_spH, d_spRh;
    SplineDev spline{};
    LdaConstants lda{};
    DevBuf<double> d_center_xyz, d_exp_alpha, d_prim_coeff, d_prim_norm;
    DevBuf<int> d_bf_center, d_bf_prim_off, d_center_exp_off, d_prim_exp, d_prim_lmn;
    DevBuf<PhiPrim> d_prims;
    DevBuf<PhiShell> d_shells;
    DevBuf<int> d_pass_rng;

    // device: per point
    DevBuf<double> d_x, d_y, d_z, d_w, d_wb, d_rho, d_dxc, d_exw, d_V, d_Vown, d_dJ, d_phi;
    // device: per iteration
    DevBuf<double> d_P, d_Praw, d_shell_raw, d_shell2, d_qatom, d_scalars, d_rho_lm, d_U_lm, d_work, d_coef, d_partial, d_res;
    DevBuf<int> d_pairs, d_cta_off, d_item_off, d_chunk_ids;
    DevBuf<double> d_rho_part;  // partial densities when a tile's slabs are split over several CTAs
    int rho_split = 1;
    size_t rho_part_stride = 0;
    int con_bc = 1;
    long n_active_chunks = 0;
    DevBuf<ConSeg> d_segs;
    int npairs = 0, nsplit = 1, con_ctas = 1, interp_chunks = 1;
    DevBuf<double> d_Vpart;
    // binned interpolation (see kernels_hartree.cuh): pairs sorted by (source atom, spline interval) at build time
    bool binned = false;
    int bin_R = 2, bin_nkeys = 0;
    long bin_nitems = 0;
    DevBuf<int> d_binoff, d_item_key, d_pair_point, d_slot_of;
    DevBuf<double> d_pair_out;

    // pinned staging
    double* h_P = nullptr;
    double* h_res = nullptr;

    // comm
    NcclComm comm = nullptr;
    // peer-memory reduction of [J | XC] (kernels_peer.cuh)
    unsigned char* xbuf = nullptr;        // this rank's exchange buffer
    size_t xbuf_bytes = 0;
    PeerSet peers{};
    std::vector<void*> peer_mapped;       // cudaIpcOpenMemHandle results to close
    bool peer_ready = false;
    unsigned long long peer_epoch = 0;

    // CUDA graph of one whole iteration (single-GPU handles): the 15 launches of a small molecule's iteration are
    // launch-latency bound, one graph launch replaces them from the second call of dftgrid_iteration_device on
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    bool capturing = false, graph_failed = false;
    long eager_iterations = 0, graph_launches_per_iter = 0;

    // timing
    cudaEvent_t ev[16]{};
    cudaEvent_t ev_sw[2]{};
    bool ev_build = false, ev_iter = false;
    double t_ms[DFTGRID_T_COUNT]{};

    ~dftgrid() {
        if (graph_exec) cudaGraphExecDestroy(graph_exec);
        if (graph) cudaGraphDestroy(graph);
        for (void* m : peer_mapped) cudaIpcCloseMemHandle(m);
        if (xbuf) cudaFree(xbuf);
        if (comm && nccl_api().ok) nccl_api().CommDestroy(comm);
        for (auto& e : ev)
            if (e) cudaEventDestroy(e);
        for (auto& e : ev_sw)
            if (e) cudaEventDestroy(e);
        if (h_P) cudaFreeHost(h_P);
        if (h_res) cudaFreeHost(h_res);
        if (stream) cudaStreamDestroy(stream);
    }
};

namespace {

void allreduce(dftgrid* h, double* buf, size_t count) {
    if (h->nranks == 1) return;
    if (!h->comm) {
        // developer switch: time one shard's kernels on a single GPU (results are then partial sums, not the molecule's)
        if (std::getenv("DFTGRID_DEBUG_SKIP_COMM")) return;
        throw std::runtime_error("nranks > 1 but dftgrid_comm_init was not called");
    }
    int rc = nccl_api().AllReduce(buf, buf, count, kNcclFloat64, kNcclSum, h->comm, h->stream);
    if (rc != 0) throw std::runtime_error(std::string("ncclAllReduce failed: ") + nccl_api().GetErrorString(rc));
}

void prepare_basis(dftgrid* h, const dftgrid_system* s) {
    // unique centres (exact coordinate match) and, per centre, the distinct exponents
    h->nbf = s->nbf;
    h->nbp = round_up(std::max(s->nbf, 1), kNbAlign);
    h->nprim = s->nprim;
    std::vector<std::vector<double>> cexp;
    h->bf_center.resize(s->nbf);
    h->bf_prim_off.assign(s->nbf + 1, 0);
    std::vector<int> prim_center(s->nprim), prim_local(s->nprim);
    int k = 0;
    for (int b = 0; b < s->nbf; b++) {
        const double* c = s->bf_center + 3 * b;
        int ci = -1;
        // consecutive CGFs normally share a centre: test the most recent ones first
        for (int t = (int)h->center_xyz.size() / 3 - 1; t >= 0; t--)
            if (h->center_xyz[3 * t] == c[0] && h->center_xyz[3 * t + 1] == c[1] && h->center_xyz[3 * t + 2] == c[2]) {
                ci = t;
                break;
            }
        if (ci < 0) {
            ci = (int)h->center_xyz.size() / 3;
            h->center_xyz.insert(h->center_xyz.end(), c, c + 3);
            cexp.emplace_back();
        }
        h->bf_center[b] = ci;
        if (s->bf_nprim[b] < 0) throw std::runtime_error("negative primitive count");
        for (int j = 0; j < s->bf_nprim[b]; j++, k++) {
            if (k >= s->nprim) throw std::runtime_error("sum(bf_nprim) exceeds nprim");
            const int l = s->lmn[3 * k], m = s->lmn[3 * k + 1], n = s->lmn[3 * k + 2];
            if (l < 0 || m < 0 || n < 0 || l + m + n > 2) throw std::runtime_error("Undefined orbital type (l+m+n > 2)");
            auto& ev = cexp[ci];
            int u = -1;
            for (size_t t = 0; t < ev.size(); t++)
                if (ev[t] == s->alpha[k]) u = (int)t;
            if (u < 0) {
                u = (int)ev.size();
                ev.push_back(s->alpha[k]);
            }
            prim_center[k] = ci;
            prim_local[k] = u;
            h->prim_coeff.push_back(s->coeff[k]);
            h->prim_norm.push_back(s->norm[k]);
            h->prim_lmn.push_back(l | (m << 4) | (n << 8));
        }
        h->bf_prim_off[b + 1] = k;
    }
    if (k != s->nprim) throw std::runtime_error("sum(bf_nprim) != nprim");
    h->center_exp_off.assign(cexp.size() + 1, 0);
    for (size_t c = 0; c < cexp.size(); c++) {
        if ((int)cexp[c].size() > kPhiMaxExp) throw std::runtime_error("too many distinct exponents on one centre");
        h->center_exp_off[c + 1] = h->center_exp_off[c] + (int)cexp[c].size();
        h->exp_alpha.insert(h->exp_alpha.end(), cexp[c].begin(), cexp[c].end());
    }
    h->prim_exp.resize(s->nprim);
    for (int t = 0; t < s->nprim; t++) h->prim_exp[t] = h->center_exp_off[prim_center[t]] + prim_local[t];
}

// Phase-timer events.  While the iteration is being captured into a CUDA graph the record becomes an external event
// node, so the events are really recorded at every replay and cudaEventElapsedTime keeps working.
void record(dftgrid* h, int i) {
    if (h->capturing)
        CK(cudaEventRecordWithFlags(h->ev[i], h->stream, cudaEventRecordExternal));
    else
        CK(cudaEventRecord(h->ev[i], h->stream));
}

float elapsed(dftgrid* h, int a, int b) {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->ev[a], h->ev[b]));
    return ms;
}

// Stream-K schedule of the [XC | J] contraction (see kernels_dense.cuh): the items' costs for ONE chunk are laid end to
// end and cut into equal shares, one per CTA (one CTA per SM); a share is 1-3 segments given as fixed-point fractions
// of the item's chunks (which chunks those are is decided on the device by a low-discrepancy hash).
struct ContractSchedule {
    std::vector<int> pairs;     // [npairs][2] upper-triangular tile pairs
    std::vector<ConSeg> segs;   // CTA after CTA; an item's segments are consecutive
    std::vector<int> cta_off;   // [nctas+1] into segs
    std::vector<int> item_off;  // [nitems+1] into segs, item = z * npairs + pair
    int npairs = 0, bc = 1;
};

// Pure host arithmetic (no device): also exported as dftgrid_debug_contract_schedule for the CPU test-suite.
void compute_contract_schedule(int nbp, long nchunk, int nsm, ContractSchedule& S) {
    S = ContractSchedule();
    const int nt = (nbp + kTileM - 1) / kTileM;
    for (int i = 0; i < nt; i++)
        for (int j = i; j < nt; j++) {
            S.pairs.push_back(i);
            S.pairs.push_back(j);
        }
    S.npairs = (int)S.pairs.size() / 2;
    const int npairs = S.npairs, nitems = 2 * npairs;
    std::vector<double> cost(nitems);
    double W1 = 0.0;  // cost of all items for one chunk
    for (int it = 0; it < nitems; it++) {
        const int ti = S.pairs[2 * (it % npairs)], tj = S.pairs[2 * (it % npairs) + 1];
        // relative cost of one k-chunk of this tile pair (full off-diagonal 128x128 tile = 20), calibrated by sweeps at
        // (H2O)64, (H2O)32 and C40H82: a 64-wide edge tile issues half the DMMAs but pays the same loads (10.5), a
        // diagonal tile issues 17 of 32 DMMAs per warp (11.5), the 64-wide diagonal tile 8 of 32 (6.5; under-estimating
        // it makes its CTA the straggler, 5 costs 14 %)
        const char* nc = std::getenv("DFTGRID_NARROW_COST");
        const char* dc = std::getenv("DFTGRID_DIAG_COST");
        const char* ec = std::getenv("DFTGRID_EDGE_DIAG_COST");
        const bool narrow = std::min(kTileN, nbp - tj * kTileN) <= 64;
        const double c_narrow = nc ? std::atof(nc) : 10.5, c_diag = dc ? std::atof(dc) : 11.5, c_edge_diag = ec ? std::atof(ec) : 6.5;
        cost[it] = ti == tj ? (narrow ? c_edge_diag : c_diag) : (narrow ? c_narrow : 20.0);
        W1 += cost[it];
    }
    // block length (the period at which a CTA with several segments alternates between them): ~120 MB of Phi rows.
    // Measured at (H2O)64: DRAM reads 15.1 GB without blocks, 9.8 GB with 80-160 MB blocks, 10.3 GB at 40 MB where the
    // accumulator parking starts to cost time; at least 64 chunks; one block when the shard is smaller.
    {
        const double chunk_bytes = (double)kTileK * nbp * sizeof(double);
        double l2_mb = 120.0;
        if (const char* e = std::getenv("DFTGRID_L2_BLOCK_MB")) l2_mb = std::atof(e);  // developer sweep; <= 0: one block
        long bc = l2_mb > 0 ? (long)(l2_mb * 1e6 / chunk_bytes) : nchunk;
        bc = std::max<long>(bc, 64);
        if (bc >= nchunk) bc = std::max<long>(nchunk, 1);
        const long nblock = (nchunk + bc - 1) / bc;
        if (nblock > 0) bc = (nchunk + nblock - 1) / nblock;  // equal blocks
        S.bc = (int)bc;
    }
    const double W = W1 * (double)nchunk;
    const int G = (int)std::max<long>(1, std::min<long>(nsm, (long)(W / 40.0) > 0 ? (long)(W / 40.0) : 1));
    S.cta_off.assign(1, 0);
    S.item_off.assign(nitems + 1, 0);
    // cumulative cost positions: item `it` occupies [start, start + cost[it]) of [0, W1)
    {
        int it = 0;
        double start = 0.0;
        for (int c = 0; c < G; c++) {
            const double lo = W1 * c / G, hi = c == G - 1 ? W1 : W1 * (c + 1) / G;
            while (it < nitems && start + cost[it] <= lo) {  // items that end before this share
                start += cost[it];
                it++;
            }
            int j = it;
            double sj = start;
            // position x inside item j as a 31-bit fixed-point fraction of the item's chunks; the same value for the
            // share that ends at x and the share that starts there
            auto frac = [&](double x, double s0, double cj) -> unsigned {
                if (x <= s0) return 0u;
                if (x >= s0 + cj) return 0x80000000u;
                return (unsigned)((x - s0) / cj * 2147483648.0);
            };
            while (j < nitems && sj < hi) {
                const unsigned tb = frac(lo, sj, cost[j]), te = frac(hi, sj, cost[j]);
                if (te > tb) S.segs.push_back(ConSeg{j / npairs, j % npairs, tb, te});
                sj += cost[j];
                j++;
            }
            S.cta_off.push_back((int)S.segs.size());
        }
    }
    // A boundary W1*c/G is the same double for the share ending there and the share starting there, and the item starts
    // are exact sums of small numbers, so te of one segment == tb of the next bit for bit.  Segments are emitted CTA
    // after CTA in item-major order, hence an item's segments are consecutive: item_off indexes `segs` directly.
    {
        int pos = 0;
        for (int it = 0; it < nitems; it++) {
            S.item_off[it] = pos;
            while (pos < (int)S.segs.size() && S.segs[pos].z * npairs + S.segs[pos].pair == it) pos++;
        }
        S.item_off[nitems] = pos;
        if (pos != (int)S.segs.size()) throw std::runtime_error("internal error: contraction segments are not item-major");
    }
}

void build_contract_schedule(dftgrid* h, long nchunk, int nsm) {
    cudaStream_t st = h->stream;
    ContractSchedule S;
    compute_contract_schedule(h->nbp, nchunk, nsm, S);
    h->npairs = S.npairs;
    h->con_bc = S.bc;
    h->d_pairs.upload(S.pairs, st);
    h->con_ctas = (int)S.cta_off.size() - 1;
    h->nsplit = (int)S.segs.size();
    h->d_segs.alloc(S.segs.size());
    CK(cudaMemcpyAsync(h->d_segs.p, S.segs.data(), S.segs.size() * sizeof(ConSeg), cudaMemcpyHostToDevice, st));
    h->d_cta_off.upload(S.cta_off, st);
    h->d_item_off.upload(S.item_off, st);
    h->d_partial.alloc((size_t)S.segs.size() * kTileM * kTileN);
    CK(cudaStreamSynchronize(st));  // S goes out of scope
}

// Lists of the 32-point chunks / 128-point tiles of Phi that hold any non-zero amplitude (k_chunk_flags), and the
// contraction schedule over the non-zero chunks.
void build_active_lists(dftgrid* h, int nsm) {
    const GridShape& g = h->g;
    cudaStream_t st = h->stream;
    const long nchunk = (g.nloc + kTileK - 1) / kTileK;
    std::vector<int> flags((size_t)nchunk, 1);
    if (nchunk > 0 && !std::getenv("DFTGRID_NO_ZERO_SKIP")) {  // developer A/B switch
        DevBuf<int> d_flags;
        d_flags.alloc((size_t)nchunk);
        k_chunk_flags<<<(unsigned)((nchunk * 32 + 255) / 256), 256, 0, st>>>(h->d_phi.p, nchunk, h->nbp, d_flags.p);
        h->launches++;
        CK(cudaMemcpyAsync(flags.data(), d_flags.p, (size_t)nchunk * sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    std::vector<int> chunk_ids;
    for (long c = 0; c < nchunk; c++)
        if (flags[c]) chunk_ids.push_back((int)c);
    h->n_active_chunks = (long)chunk_ids.size();
    while (chunk_ids.size() % 4 != 0 || chunk_ids.empty()) chunk_ids.push_back(-1);  // k_rho_tma reads groups of four
    h->d_chunk_ids.upload(chunk_ids, st);
    CK(cudaStreamSynchronize(st));
    {
        // k_rho_tma work items: one CTA per 128-point tile when that gives many waves over the SMs; with few waves
        // (sharded grids, small molecules) the tail wave costs up to 1/waves, so a tile's column slabs are dealt to 2 or 3 CTAs
        const double waves = (double)((h->n_active_chunks + 3) / 4) / (double)nsm;
        const int nslab = (h->nbp + kTileN - 1) / kTileN;
        int split = waves >= 40.0 ? 1 : (waves >= 12.0 ? 2 : 3);  // measured at (H2O)64: 28 waves 11.07 -> 10.98 ms (2 CTAs), 3.5 waves 1.56 -> 1.46 ms (3 CTAs)
        if (const char* e = std::getenv("DFTGRID_RHO_SPLIT")) split = std::atoi(e);  // developer A/B switch
        h->rho_split = std::max(1, std::min(split, std::min(nslab, 3)));
        h->rho_part_stride = (size_t)g.nloc + 64;
        if (h->rho_split > 1) {
            h->d_rho_part.alloc((size_t)h->rho_split * h->rho_part_stride);
            h->d_rho_part.zero(st);  // skipped (all-zero) chunks are never written
        }
    }
    build_contract_schedule(h, h->n_active_chunks, nsm);
}

// Sort the (local point, source atom) pairs of the cross-atom interpolation into (atom, spline interval) bins.
// Geometry only, so it is done once per grid.  Falls back to the point-parallel kernels when lmax is not one of the
// presets, the pair count does not fit 32-bit slots, or the lists do not fit in free device memory.
void build_pair_bins(dftgrid* h) {
    const GridShape& g = h->g;
    cudaStream_t st = h->stream;
    h->binned = false;
    if (std::getenv("DFTGRID_INTERP_POINTWISE")) return;  // developer A/B switch
    if (!(g.lmax == 5 || g.lmax == 8 || g.lmax == 10 || g.lmax == 11)) return;
    if (g.nloc == 0 || g.natoms < 2) return;
    if (const char* r = std::getenv("DFTGRID_INTERP_R")) h->bin_R = std::atoi(r);
    if (h->bin_R < 2 || h->bin_R > 4) h->bin_R = 2;
    const int unit = 32 * h->bin_R;
    const int nkeys = g.natoms * g.nrad;
    const double pairs_max = (double)g.nloc * (g.natoms - 1) + (double)nkeys * unit;
    if (pairs_max >= 2.0e9) return;
    size_t mem_free = 0, mem_total = 0;
    CK(cudaMemGetInfo(&mem_free, &mem_total));
    if (pairs_max * 12.0 + (double)g.nloc * g.natoms * 4.0 > 0.5 * (double)mem_free) return;

    DevBuf<int> d_counts, d_cursor;
    d_counts.alloc(nkeys);
    d_counts.zero(st);
    const unsigned blocks = (unsigned)((g.nloc + 255) / 256);
    const size_t smem = (size_t)g.nrad * sizeof(double);
    k_bin_pairs<<<blocks, 256, smem, st>>>(g, h->d_atom_xyz.p, h->d_x.p, h->d_y.p, h->d_z.p, h->d_xs.p, d_counts.p, nullptr, nullptr, nullptr, nullptr);
    std::vector<int> counts(nkeys), binoff(nkeys + 1, 0);
    CK(cudaMemcpyAsync(counts.data(), d_counts.p, nkeys * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    long total = 0;
    for (int i = 0; i < nkeys; i++) {
        binoff[i] = (int)total;
        total += ((long)counts[i] + unit - 1) / unit * unit;
    }
    if (total >= 2147483647L) return;
    binoff[nkeys] = (int)total;
    h->bin_nkeys = nkeys;
    h->bin_nitems = total / unit;
    h->d_binoff.upload(binoff, st);
    d_cursor.alloc(nkeys);
    d_cursor.zero(st);
    h->d_pair_point.alloc((size_t)total + 1);
    CK(cudaMemsetAsync(h->d_pair_point.p, 0xFF, ((size_t)total + 1) * sizeof(int), st));
    h->d_slot_of.alloc((size_t)g.natoms * g.nloc);
    h->d_pair_out.alloc((size_t)total + 1);
    k_bin_pairs<<<blocks, 256, smem, st>>>(g, h->d_atom_xyz.p, h->d_x.p, h->d_y.p, h->d_z.p, h->d_xs.p, d_counts.p, h->d_binoff.p, d_cursor.p,
                                           h->d_pair_point.p, h->d_slot_of.p);
    h->d_item_key.alloc((size_t)h->bin_nitems + 1);
    if (h->bin_nitems > 0)
        k_item_keys<<<(unsigned)((h->bin_nitems + 255) / 256), 256, 0, st>>>(h->d_binoff.p, nkeys, unit, h->bin_nitems, h->d_item_key.p);
    h->launches += 3;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));  // binoff (host vector) and the scratch buffers go out of scope
    h->binned = true;
}

void do_build(dftgrid* h) {
    const dftgrid_params& prm = h->prm;
    if (prm.lebedev_order < 0 || prm.lebedev_order > 10) throw std::runtime_error("lebedev_order must be 0..10");
    if (prm.lmax < 0 || prm.lmax > kMaxL) throw std::runtime_error("lmax out of range");
    cudaStream_t st = h->stream;
    GridShape& g = h->g;
    g.natoms = h->natoms;
    g.nrad = prm.radial_points;
    g.nang = kLebedevCounts[prm.lebedev_order];
    g.lmax = prm.lmax;
    g.nlm = (prm.lmax + 1) * (prm.lmax + 1);
    g.npts = (long)g.natoms * g.nrad * g.nang;
    const long nshell = (long)g.natoms * g.nrad;
    dftgrid_shard_range(nshell, h->rank, h->nranks, &g.shell0, &g.nshell_loc);
    g.nloc = g.nshell_loc * g.nang;
    h->leb_off = lebedev_offset(prm.lebedev_order);

    // ---- host tables
    std::vector<double> r, wr, Y, pre;
    make_radial(g.nrad, r, wr);
    make_ylm_table(h->leb_off, g.nang, g.lmax, Y, pre);
    std::vector<double> Yt((size_t)g.nang * g.nlm);
    for (int j = 0; j < g.nang; j++)
        for (int lm = 0; lm < g.nlm; lm++) Yt[(size_t)lm * g.nang + j] = Y[(size_t)j * g.nlm + lm];
    PoissonLU lu;
    make_poisson_lu(g.nrad, g.lmax, r, lu);
    SplineSystem sp;
    make_spline_system(g.nrad, r, sp);
    make_lda_constants(h->lda);
    std::vector<double> leb((size_t)g.nang * 4);
    for (int j = 0; j < g.nang; j++)
        for (int c = 0; c < 4; c++) leb[4 * j + c] = kLebedevTable[h->leb_off + j][c];
    std::vector<double> Rdist((size_t)g.natoms * g.natoms, 0.0);
    for (int a = 0; a < g.natoms; a++)
        for (int b = 0; b < g.natoms; b++) {
            // (p2 - p1).norm() of src/moleculargrid.cpp:291-293; symmetric in (a,b) bit for bit
            const double dx = h->atom_xyz[3 * b] - h->atom_xyz[3 * a], dy = h->atom_xyz[3 * b + 1] - h->atom_xyz[3 * a + 1],
                         dz = h->atom_xyz[3 * b + 2] - h->atom_xyz[3 * a + 2];
            Rdist[(size_t)a * g.natoms + b] = std::sqrt(dx * dx + dy * dy + dz * dz);
        }

    h->d_atom_xyz.upload(h->atom_xyz, st);
    h->d_Rdist.upload(Rdist, st);
    h->d_rtab.upload(r, st);
    h->d_wrad.upload(wr, st);
    h->d_leb.upload(leb, st);
    h->d_Y.upload(Y, st);
    h->d_Yt.upload(Yt, st);
    h->d_pre.upload(pre, st);
    h->d_lu.upload(lu.lu, st);
    h->d_perm.upload(lu.perm, st);
    h->d_lo.upload(lu.lo, st);
    h->d_hi.upload(lu.hi, st);
    h->d_xs.upload(sp.x, st);
    h->d_spA.upload(sp.A, st);
    h->d_spCp.upload(sp.Cp, st);
    h->d_spDen.upload(sp.den, st);
    h->d_spH.upload(sp.h, st);
    h->d_spRh.upload(sp.rh, st);
    h->spline = SplineDev{h->d_xs.p, h->d_spA.p, h->d_spCp.p, h->d_spDen.p, h->d_spH.p, h->d_spRh.p,
                          sp.first_w[0], sp.first_w[1], sp.last_w[0], sp.last_w[1]};
    h->d_center_xyz.upload(h->center_xyz, st);
    h->d_exp_alpha.upload(h->exp_alpha, st);
    h->d_prim_coeff.upload(h->prim_coeff, st);
    h->d_prim_norm.upload(h->prim_norm, st);
    h->d_bf_center.upload(h->bf_center, st);
    h->d_bf_prim_off.upload(h->bf_prim_off, st);
    h->d_center_exp_off.upload(h->center_exp_off, st);
    h->d_prim_exp.upload(h->prim_exp, st);
    h->d_prim_lmn.upload(h->prim_lmn, st);

    // ---- per-point storage
    const size_t nl = (size_t)g.nloc, nlp = nl + 64;
    h->d_x.alloc(nlp);
    h->d_y.alloc(nlp);
    h->d_z.alloc(nlp);
    h->d_w.alloc(nlp);
    h->d_wb.alloc(nlp);
    h->d_rho.alloc(nlp);
    h->d_dxc.alloc(nlp);
    h->d_exw.alloc(nlp);
    h->d_V.alloc(nlp);
    h->d_Vown.alloc(nlp);
    h->d_dJ.alloc(nlp);
    h->d_dxc.zero(st);
    h->d_dJ.zero(st);
    h->d_rho.zero(st);  // tiles of exact-zero amplitudes are skipped by k_rho_tma: their density stays 0
    // whole 128-row tiles / 32-row chunks must be readable by the bulk-copy producers: rows past nloc are zero
    h->d_phi.alloc((nl + kTileM - 1) / kTileM * kTileM * (size_t)h->nbp + 64);
    h->d_phi.zero(st);
    const size_t nsys = (size_t)g.natoms * g.nlm;
    h->d_P.alloc((size_t)h->nbp * h->nbp);
    h->d_P.zero(st);
    h->d_Praw.alloc((size_t)h->nbf * h->nbf);
    h->d_shell_raw.alloc((size_t)nshell);
    h->d_shell2.alloc((size_t)nshell * 2 + (size_t)nshell * g.nlm);  // [shell sums (2 per shell) | rho_lm] contiguous: one collective
    h->d_qatom.alloc(g.natoms);
    h->d_scalars.alloc(4);
    h->d_U_lm.alloc((size_t)nshell * g.nlm);
    h->d_work.alloc(std::max((size_t)(g.nrad + 2) * nsys, (size_t)2 * g.nrad * nsys));
    h->d_coef.alloc((size_t)g.natoms * g.nrad * g.nlm * 4);
    h->d_res.alloc((size_t)2 * h->nbf * h->nbf + 2);

    int nsm = 148;
    CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device));
    {
        // source-atom chunks of the interpolation kernel: enough CTAs for >= ~8 full waves (6 CTAs of 128 threads per SM)
        const long ctas = (g.nloc + 127) / 128, wave = 6L * nsm;
        long chunks = ctas > 0 ? (8 * wave + ctas - 1) / ctas : 1;
        h->interp_chunks = (int)std::max<long>(1, std::min<long>(chunks, std::min<long>(g.natoms, 32)));  // point-parallel fallback only
    }
    if (!h->h_P) CK(cudaMallocHost(&h->h_P, sizeof(double) * std::max<size_t>(1, (size_t)h->nbf * h->nbf)));
    if (!h->h_res) CK(cudaMallocHost(&h->h_res, sizeof(double) * ((size_t)2 * h->nbf * h->nbf + 2)));

    CK(cudaFuncSetAttribute(k_phi, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)(((size_t)kPhiPts * (kPhiCols + 1) + (size_t)kPhiMaxExp * kPhiPts) * sizeof(double))));
    CK(cudaFuncSetAttribute(k_rho_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRhoTmaSmemBytes));
    CK(cudaFuncSetAttribute(k_contract_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kConTmaSmemBytes));
    const size_t becke_smem = (size_t)kBeckeWarps * 2 * g.natoms * sizeof(double);
    if (becke_smem > 200 * 1024) throw std::runtime_error("too many atoms for the Becke kernel's shared-memory layout");
    CK(cudaFuncSetAttribute(k_becke, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(becke_smem, 1024)));

    // ---- kernels
    record(h, 0);
    if (g.nloc > 0) {
        k_points<<<(unsigned)((g.nloc + 255) / 256), 256, 0, st>>>(g, h->d_atom_xyz.p, h->d_rtab.p, h->d_wrad.p, h->d_leb.p,
                                                                  h->d_x.p, h->d_y.p, h->d_z.p, h->d_w.p);
        h->launches++;
    }
    record(h, 1);
    if (g.nloc > 0) {
        k_becke<<<(unsigned)((g.nloc + kBeckeWarps - 1) / kBeckeWarps), kBeckeWarps * 32, becke_smem, st>>>(
            g, h->d_atom_xyz.p, nullptr, h->d_Rdist.p, h->d_x.p, h->d_y.p, h->d_z.p, h->d_w.p, h->d_wb.p);
        h->launches++;
    }
    record(h, 2);
    if (g.nloc > 0) {
        // group the columns into shells (see kernels_grid.cuh): S / P / D runs that share their primitives, else generic
        std::vector<PhiShell> shells;
        std::vector<PhiPrim> prims;
        auto slot_of = [&](int k, int centre) { return h->prim_exp[k] - h->center_exp_off[centre]; };
        auto same_prims = [&](int b0, int b1) {  // same centre, exponents and coefficients
            if (h->bf_center[b0] != h->bf_center[b1]) return false;
            const int n0 = h->bf_prim_off[b0 + 1] - h->bf_prim_off[b0], n1 = h->bf_prim_off[b1 + 1] - h->bf_prim_off[b1];
            if (n0 != n1) return false;
            for (int k = 0; k < n0; k++) {
                const int k0 = h->bf_prim_off[b0] + k, k1 = h->bf_prim_off[b1] + k;
                if (h->prim_exp[k0] != h->prim_exp[k1] || h->prim_coeff[k0] != h->prim_coeff[k1]) return false;
            }
            return true;
        };
        auto all_lmn = [&](int b, int code) {
            for (int k = h->bf_prim_off[b]; k < h->bf_prim_off[b + 1]; k++)
                if (h->prim_lmn[k] != code) return false;
            return true;
        };
        auto same_norms = [&](int b0, int b1) {
            const int n = h->bf_prim_off[b0 + 1] - h->bf_prim_off[b0];
            for (int k = 0; k < n; k++)
                if (h->prim_norm[h->bf_prim_off[b0] + k] != h->prim_norm[h->bf_prim_off[b1] + k]) return false;
            return true;
        };
        const int LX = 1, LY = 1 << 4, LZ = 1 << 8;
        for (int b = 0; b < h->nbf;) {
            const int centre = h->bf_center[b], k0 = h->bf_prim_off[b], np = h->bf_prim_off[b + 1] - k0;
            int type = kShellGeneric, ncol = 1;
            if (all_lmn(b, 0)) {
                type = kShellS;
            } else if (b + 2 < h->nbf && all_lmn(b, LX) && all_lmn(b + 1, LY) && all_lmn(b + 2, LZ) && same_prims(b, b + 1) && same_prims(b, b + 2) &&
                       same_norms(b, b + 1) && same_norms(b, b + 2)) {
                type = kShellP;
                ncol = 3;
            } else if (b + 5 < h->nbf && all_lmn(b, 2 * LX) && all_lmn(b + 1, LX + LY) && all_lmn(b + 2, LX + LZ) && all_lmn(b + 3, 2 * LY) &&
                       all_lmn(b + 4, LY + LZ) && all_lmn(b + 5, 2 * LZ) && same_prims(b, b + 1) && same_prims(b, b + 2) && same_prims(b, b + 3) &&
                       same_prims(b, b + 4) && same_prims(b, b + 5) && same_norms(b, b + 3) && same_norms(b, b + 5) && same_norms(b + 1, b + 2) &&
                       same_norms(b + 1, b + 4)) {
                type = kShellD;
                ncol = 6;
            }
            shells.push_back(PhiShell{type, b, centre, (int)prims.size(), np, ncol, 0, 0});
            for (int k = 0; k < np; k++) {
                const double nb_ = type == kShellD ? h->prim_norm[h->bf_prim_off[b + 1] + k] : 0.0;
                prims.push_back(PhiPrim{h->prim_coeff[k0 + k], h->prim_norm[k0 + k], nb_, slot_of(k0 + k, centre), h->prim_lmn[k0 + k]});
            }
            b += ncol;
        }
        const int npass = h->nbp / kPhiCols;
        std::vector<int> pass_rng(2 * (size_t)npass, 0);
        for (int pass = 0; pass < npass; pass++) {
            const int c0 = pass * kPhiCols, c1 = c0 + kPhiCols;
            int first = (int)shells.size(), last = first;
            for (int si = 0; si < (int)shells.size(); si++)
                if (shells[si].col < c1 && shells[si].col + shells[si].ncol > c0) {
                    first = std::min(first, si);
                    last = si + 1;
                }
            if (first > last) first = last;
            pass_rng[2 * pass] = first;
            pass_rng[2 * pass + 1] = last;
        }
        h->d_shells.alloc(shells.size());
        CK(cudaMemcpyAsync(h->d_shells.p, shells.data(), shells.size() * sizeof(PhiShell), cudaMemcpyHostToDevice, st));
        h->d_prims.alloc(prims.size());
        CK(cudaMemcpyAsync(h->d_prims.p, prims.data(), prims.size() * sizeof(PhiPrim), cudaMemcpyHostToDevice, st));
        h->d_pass_rng.upload(pass_rng, st);
        PhiBasis B{h->nbf, h->nbp, (int)shells.size(), h->d_shells.p, h->d_prims.p, h->d_center_exp_off.p, h->d_exp_alpha.p, h->d_center_xyz.p,
                   h->d_pass_rng.p};
        // the exponential table is sized for the molecule's largest centre (rounded up to the group of four), not for
        // kPhiMaxExp: with 6-31G (10 exponents on O) four CTAs fit an SM instead of three
        int max_exp = 1;
        for (size_t c = 0; c + 1 < h->center_exp_off.size(); c++) max_exp = std::max(max_exp, h->center_exp_off[c + 1] - h->center_exp_off[c]);
        max_exp = std::min(kPhiMaxExp, (max_exp + 3) / 4 * 4);
        const size_t smem = ((size_t)kPhiPts * (kPhiCols + 1) + (size_t)max_exp * kPhiPts) * sizeof(double);
        k_phi<<<(unsigned)((g.nloc + kPhiPts - 1) / kPhiPts), kPhiPts, smem, st>>>(g.nloc, B, h->d_x.p, h->d_y.p, h->d_z.p, h->d_phi.p);
        h->launches++;
    }
    record(h, 3);
    build_active_lists(h, nsm);
    build_pair_bins(h);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    h->t_ms[DFTGRID_T_POINTS] = elapsed(h, 0, 1);
    h->t_ms[DFTGRID_T_BECKE] = elapsed(h, 1, 2);
    h->t_ms[DFTGRID_T_PHI] = elapsed(h, 2, 3);
    h->built = true;
}

// rho = 2 phi^T P phi, rescale to sum(Z), LDA pointwise, charge / E_xc sums
void run_density(dftgrid* h) {
    cudaStream_t st = h->stream;
    const GridShape& g = h->g;
    const long nshell = (long)g.natoms * g.nrad;
    record(h, 4);
    if (g.nloc > 0) {
        if (h->n_active_chunks > 0) {
            const unsigned tiles = (unsigned)((h->n_active_chunks + 3) / 4);
            if (h->rho_split == 1) {
                k_rho_tma<<<tiles, kRhoTmaThreads, kRhoTmaSmemBytes, st>>>(h->d_phi.p, h->d_P.p, h->d_chunk_ids.p, h->d_rho.p, 0, g.nloc, h->nbp);
            } else {
                k_rho_tma<<<dim3(tiles, h->rho_split), kRhoTmaThreads, kRhoTmaSmemBytes, st>>>(h->d_phi.p, h->d_P.p, h->d_chunk_ids.p, h->d_rho_part.p,
                                                                                          (long)h->rho_part_stride, g.nloc, h->nbp);
                k_rho_combine<<<(unsigned)((g.nloc + 255) / 256), 256, 0, st>>>(h->d_rho_part.p, (long)h->rho_part_stride, h->rho_split, g.nloc, h->d_rho.p);
                h->launches++;
            }
        }
        h->launches++;
    }
    record(h, 5);
    const unsigned sblocks = (unsigned)((g.nshell_loc * 32 + 255) / 256);
    if (h->nranks > 1) h->d_shell_raw.zero(st);
    if (g.nloc > 0) {
        k_shell_sum<<<sblocks, 256, 0, st>>>(g, h->d_w.p, h->d_rho.p, h->d_shell_raw.p, 1, 0);
        h->launches++;
    }
    allreduce(h, h->d_shell_raw.p, (size_t)nshell);
    k_totals<<<1, 256, 0, st>>>(g, h->d_shell_raw.p, 1, h->zsum, 0, h->d_qatom.p, h->d_scalars.p);
    h->launches++;
    if (h->nranks > 1) h->d_shell2.zero(st);
    if (g.nloc > 0) {
        k_scale_xc<<<(unsigned)((g.nloc + 255) / 256), 256, 0, st>>>(g.nloc, h->lda, h->d_scalars.p, h->d_w.p, h->d_rho.p, h->d_dxc.p, h->d_exw.p);
        k_shell_sum<<<sblocks, 256, 0, st>>>(g, h->d_w.p, h->d_rho.p, h->d_shell2.p, 2, 0);
        k_shell_sum<<<sblocks, 256, 0, st>>>(g, h->d_w.p, h->d_exw.p, h->d_shell2.p, 2, 1);
        h->launches += 3;
    }
    record(h, 6);
    // Ylm projection of the rescaled density; shares the collective with the shell sums
    double* rho_lm = h->d_shell2.p + (size_t)nshell * 2;
    if (g.nloc > 0) {
        const int threads = round_up(g.nlm, 32);
        k_rho_lm<<<(unsigned)g.nshell_loc, threads, 3 * g.nang * sizeof(double), st>>>(g, h->d_rho.p, h->d_wb.p, h->d_leb.p, h->d_Y.p, rho_lm);
        h->launches++;
    }
    record(h, 7);
    allreduce(h, h->d_shell2.p, (size_t)nshell * 2 + (size_t)nshell * g.nlm);
    k_totals<<<1, 256, 0, st>>>(g, h->d_shell2.p, 2, h->zsum, 1, h->d_qatom.p, h->d_scalars.p);
    h->launches++;
    record(h, 8);
    h->have_density = true;
    h->have_potential = false;
    h->contract_valid = false;
    h->timed_iter = false;
}

// Hartree potential on every local point (rho_lm -> U_lm -> splines -> V)
void run_potential(dftgrid* h) {
    cudaStream_t st = h->stream;
    const GridShape& g = h->g;
    const long nshell = (long)g.natoms * g.nrad;
    const long nsys = (long)g.natoms * g.nlm;
    double* rho_lm = h->d_shell2.p + (size_t)nshell * 2;
    record(h, 9);
    // per-thread scratch vectors in shared memory when they fit the default 48 KB (radial grids up to ~90 nodes)
    const size_t sm_poisson = (size_t)(g.nrad + 2) * 64 * sizeof(double), sm_spline = (size_t)2 * g.nrad * 64 * sizeof(double);
    const bool smp = sm_poisson <= 48 * 1024, sms = sm_spline <= 48 * 1024;
    k_poisson<<<(unsigned)((nsys + 63) / 64), 64, smp ? sm_poisson : 0, st>>>(g, g.nrad + 2, h->d_lu.p, h->d_perm.p, h->d_lo.p, h->d_hi.p,
                                                                              h->d_rtab.p, rho_lm, h->d_qatom.p, h->d_work.p, h->d_U_lm.p, smp);
    k_spline<<<(unsigned)((nsys + 63) / 64), 64, sms ? sm_spline : 0, st>>>(g, h->spline, h->d_U_lm.p, h->d_pre.p, h->d_work.p, h->d_coef.p, sms);
    h->launches += 2;
    if (g.nloc > 0) {
        k_v_own<<<(unsigned)g.nshell_loc, 128, g.nlm * sizeof(double), st>>>(g, h->d_rtab.p, h->d_leb.p, h->d_Yt.p, h->d_U_lm.p, h->d_Vown.p);
        h->launches++;
    }
    record(h, 10);
    if (g.nloc > 0 && h->binned) {
        if (h->bin_nitems > 0) {
            const unsigned bx = (unsigned)((h->bin_nitems + kBinWarps - 1) / kBinWarps);
            const size_t smem = (size_t)kBinWarps * g.nlm * 4 * sizeof(double);
#define DFG_BIN_ARGS g, h->d_atom_xyz.p, h->d_x.p, h->d_y.p, h->d_z.p, h->d_xs.p, h->d_coef.p, h->d_item_key.p, h->d_pair_point.p, h->bin_nitems, h->d_pair_out.p
#define DFG_BIN_CASE(LL)                                                                                        \
    case LL:                                                                                                    \
        if (h->bin_R == 4)                                                                                      \
            k_interp_bin<LL, 4, 3><<<bx, kBinWarps * 32, smem, st>>>(DFG_BIN_ARGS);                            \
        else if (h->bin_R == 3)                                                                                 \
            k_interp_bin<LL, 3, 4><<<bx, kBinWarps * 32, smem, st>>>(DFG_BIN_ARGS);                            \
        else if (five)                                                                                          \
            k_interp_bin<LL, 2, 5><<<bx, kBinWarps * 32, smem, st>>>(DFG_BIN_ARGS);                            \
        else                                                                                                    \
            k_interp_bin<LL, 2, 6><<<bx, kBinWarps * 32, smem, st>>>(DFG_BIN_ARGS);                            \
        break;
            static const bool five = std::getenv("DFTGRID_INTERP_MINB5") != nullptr;  // developer A/B switch
            switch (g.lmax) {
                DFG_BIN_CASE(5)
                DFG_BIN_CASE(8)
                DFG_BIN_CASE(10)
                DFG_BIN_CASE(11)
                default: throw std::runtime_error("binned interpolation: unsupported lmax");
            }
#undef DFG_BIN_CASE
#undef DFG_BIN_ARGS
            h->launches++;
        }
        k_finish_binned<<<(unsigned)((g.nloc + 255) / 256), 256, 0, st>>>(g, h->d_slot_of.p, h->d_pair_out.p, h->d_Vown.p, h->d_w.p, h->d_V.p, h->d_dJ.p);
        h->launches++;
    } else if (g.nloc > 0) {
        if (h->d_Vpart.n == 0) h->d_Vpart.alloc((size_t)h->interp_chunks * ((size_t)g.nloc + 64));
        const size_t smem_g = ((size_t)g.nrad + (size_t)(g.lmax + 1) * (g.lmax + 1) + 2 * g.lmax + 2) * sizeof(double);
        const size_t smem = smem_g + (size_t)4 * 2 * g.nlm * 4 * sizeof(double);  // + per-warp staging rows of the unrolled kernels
        const unsigned bx = (unsigned)((g.nloc + 127) / 128);
        const dim3 blocks(bx, h->interp_chunks);
#define DFG_INTERP_ARGS g, h->d_atom_xyz.p, h->d_x.p, h->d_y.p, h->d_z.p, h->d_w.p, h->d_Vown.p, h->d_xs.p, h->d_pre.p, h->d_coef.p, h->d_Vpart.p
        switch (g.lmax) {  // unrolled kernels for the reference's grid presets (src/settings.cpp:158-187), generic otherwise
            case 5: k_interp_t<5, 6><<<blocks, 128, smem, st>>>(DFG_INTERP_ARGS); break;
            case 8: k_interp_t<8, 6><<<blocks, 128, smem, st>>>(DFG_INTERP_ARGS); break;
            case 10: k_interp_t<10, 6><<<blocks, 128, smem, st>>>(DFG_INTERP_ARGS); break;
            case 11: k_interp_t<11, 6><<<blocks, 128, smem, st>>>(DFG_INTERP_ARGS); break;
            default: k_interp<<<blocks, 128, smem_g, st>>>(DFG_INTERP_ARGS); break;
        }
#undef DFG_INTERP_ARGS
        k_finish_potential<<<(unsigned)((g.nloc + 255) / 256), 256, 0, st>>>(g.nloc, h->interp_chunks, h->d_Vpart.p, h->d_w.p, h->d_V.p, h->d_dJ.p);
        h->launches += 2;
    }
    record(h, 11);
    h->have_potential = true;
}

// [XC | J] = Phi^T diag(d) Phi, results laid out as res = [J (nb^2) | XC (nb^2) | exc | nel]
void run_contract(dftgrid* h) {
    cudaStream_t st = h->stream;
    const size_t nb2 = (size_t)h->nbf * h->nbf;
    record(h, 12);
    k_contract_tma<<<h->con_ctas, kConTmaThreads, kConTmaSmemBytes, st>>>(h->d_phi.p, h->d_dxc.p, h->d_dJ.p, h->d_chunk_ids.p, h->d_pairs.p,
                                                                          h->d_segs.p, h->d_cta_off.p, h->d_partial.p, h->nbp, (int)h->n_active_chunks, h->con_bc);
    if (h->peer_ready) {
        // split-K reduction straight into this rank's exchange buffer, then the cross-rank sum over peer memory
        h->peer_epoch++;
        k_contract_reduce_publish<<<dim3(h->npairs, 2, kReduceSplit), 256, 0, st>>>(h->d_partial.p, h->d_pairs.p, h->d_item_off.p, h->npairs, h->nbf, h->nbp,
                                                                     1.0, 0.5, h->peers, h->peer_epoch);
    } else {
        k_contract_reduce<<<dim3(h->npairs, 2, kReduceSplit), 256, 0, st>>>(h->d_partial.p, h->d_pairs.p, h->d_item_off.p, h->npairs, h->nbf, h->nbp, 1.0, 0.5,
                                                             h->d_res.p + nb2, h->d_res.p);
    }
    h->launches += 2;
    // exc and nel come from the already-reduced shell sums (identical on every rank): appended after the reduced block
    CK(cudaMemcpyAsync(h->d_res.p + 2 * nb2, h->d_scalars.p + 2, sizeof(double), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(h->d_res.p + 2 * nb2 + 1, h->d_scalars.p + 1, sizeof(double), cudaMemcpyDeviceToDevice, st));
    record(h, 13);
    if (h->peer_ready) {
        const size_t n = 2 * nb2;
        const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, 296);
        k_peer_sum<<<blocks, 256, 0, st>>>(h->peers, h->peer_epoch, h->nbf, 2, h->d_res.p);
        h->launches++;
    } else {
        allreduce(h, h->d_res.p, 2 * nb2);
    }
    record(h, 14);
    h->contract_valid = h->have_potential;
    h->timed_iter = h->have_potential;
}

// Zero-padded copy of the density matrix for k_rho_tma with the 32x32 diagonal blocks halved (exact; see
// kernels_dense.cuh: the blocks above the diagonal are visited once and stand for both triangles).
__global__ void k_pad_P(const double* __restrict__ Praw, double* __restrict__ P, int nb, int nbp) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)nb * nb) return;
    const int i = (int)(t / nb), j = (int)(t % nb);
    P[(size_t)i * nbp + j] = Praw[t] * (i / kTileK == j / kTileK ? 0.5 : 1.0);
}

// True when the caller's host buffer is page-locked (cudaHostAlloc / cudaHostRegister, e.g. a pinned torch tensor):
// the DMA engine can then read or write it directly and the bounce through the handle's own pinned staging is skipped.
bool is_pinned_host(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

// Returns true when the copy reads the caller's buffer asynchronously (the caller's buffer must stay untouched until
// the stream has been synchronised).
bool upload_P(dftgrid* h, const double* P) {
    if (!h->built) throw std::runtime_error("dftgrid_build has not been called");
    const size_t nb2 = (size_t)h->nbf * h->nbf;
    const bool direct = is_pinned_host(P);
    if (!direct) std::memcpy(h->h_P, P, nb2 * sizeof(double));
    CK(cudaMemcpyAsync(h->d_Praw.p, direct ? P : h->h_P, nb2 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    // P is symmetric, so Eigen's column-major and our row-major padded copy coincide
    k_pad_P<<<(unsigned)((nb2 + 255) / 256), 256, 0, h->stream>>>(h->d_Praw.p, h->d_P.p, h->nbf, h->nbp);
    h->launches++;
    return direct;
}

void finish_timings(dftgrid* h) {
    h->t_ms[DFTGRID_T_RHO] = elapsed(h, 4, 5);
    h->t_ms[DFTGRID_T_XCPOINT] = elapsed(h, 5, 6) + elapsed(h, 7, 8);
    h->t_ms[DFTGRID_T_RHOLM] = elapsed(h, 6, 7);
    h->t_ms[DFTGRID_T_POISSON] = elapsed(h, 9, 10);
    h->t_ms[DFTGRID_T_INTERP] = elapsed(h, 10, 11);
    h->t_ms[DFTGRID_T_CONTRACT] = elapsed(h, 12, 13);
    h->t_ms[DFTGRID_T_COMM] = elapsed(h, 13, 14);
    h->t_ms[DFTGRID_T_TOTAL] = elapsed(h, 4, 14);
}

template <typename F>
int guarded(F&& f) {
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    } catch (...) {
        g_error = "unknown error";
        return 2;
    }
}

void use_device(dftgrid* h) { CK(cudaSetDevice(h->device)); }

// After a stream synchronisation: a bounded spin-wait of the peer-memory reduction that gave up leaves a flag behind.
void check_peer_error(dftgrid* h) {
    if (!h->peer_ready) return;
    unsigned long long err = 0;
    CK(cudaMemcpy(&err, h->xbuf + offsetof(PeerHeader, error), sizeof err, cudaMemcpyDeviceToHost));
    if (err) throw std::runtime_error("peer-memory reduction timed out waiting for another rank");
}

template <typename T>
void download(dftgrid* h, const DevBuf<T>& b, T* out, size_t count) {
    use_device(h);
    CK(cudaMemcpyAsync(out, b.p, count * sizeof(T), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
}

}  // namespace

extern "C" {

const char* dftgrid_last_error(void) { return g_error.c_str(); }
int dftgrid_abi_version(void) { return 1; }

int dftgrid_create(dftgrid_t** out, const dftgrid_system* sys, const dftgrid_params* prm, int device, int rank, int nranks) {
    return guarded([&] {
        if (!out || !sys || !prm) throw std::runtime_error("null argument");
        *out = nullptr;
        if (sys->natoms <= 0 || sys->nbf <= 0 || sys->nprim <= 0) throw std::runtime_error("empty system");
        if (nranks < 1 || rank < 0 || rank >= nranks) throw std::runtime_error("bad rank/nranks");
        if (prm->radial_points < 6) throw std::runtime_error("radial_points must be at least 6");
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            throw std::runtime_error(std::string("no CUDA device available (this library has no CPU fallback): ") + cudaGetErrorString(e));
        if (device < 0 || device >= ndev) throw std::runtime_error("bad CUDA device ordinal");
        std::unique_ptr<dftgrid> h(new dftgrid());
        h->device = device;
        h->rank = rank;
        h->nranks = nranks;
        h->prm = *prm;
        h->natoms = sys->natoms;
        h->Z.assign(sys->Z, sys->Z + sys->natoms);
        h->atom_xyz.assign(sys->xyz, sys->xyz + 3 * sys->natoms);
        h->zsum = 0.0;
        for (int a = 0; a < sys->natoms; a++) h->zsum += (double)sys->Z[a];  // charge += get_atomic_charge(i), src/moleculargrid.cpp:138
        prepare_basis(h.get(), sys);
        CK(cudaSetDevice(device));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10) throw std::runtime_error("this library is built for sm_100a (B200) only");
        CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        for (auto& ev : h->ev) CK(cudaEventCreate(&ev));
        for (auto& ev : h->ev_sw) CK(cudaEventCreate(&ev));
        *out = h.release();
    });
}

void dftgrid_destroy(dftgrid_t* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->peer_ready && h->peer_epoch > 0) {
        // another rank may still be summing this rank's exchange buffer (its k_peer_sum of the last epoch): wait, bounded,
        // until every peer has written consumed_by[peer] >= the last epoch into THIS rank's header, then free it
        const auto t0 = std::chrono::steady_clock::now();
        for (int r = 0; r < h->peers.nranks; r++) {
            if (r == h->rank) continue;
            unsigned long long consumed = 0;
            do {
                if (cudaMemcpy(&consumed, h->xbuf + offsetof(PeerHeader, consumed_by) + r * sizeof(unsigned long long), sizeof consumed,
                               cudaMemcpyDeviceToHost) != cudaSuccess) {
                    cudaGetLastError();
                    break;
                }
            } while (consumed < h->peer_epoch && std::chrono::steady_clock::now() - t0 < std::chrono::seconds(5));
        }
    }
    delete h;
}

int dftgrid_shard_range(long nshell_total, int rank, int nranks, long* first_shell, long* nshell) {
    if (nranks < 1 || rank < 0 || rank >= nranks || nshell_total < 0 || !first_shell || !nshell) {
        g_error = "dftgrid_shard_range: bad arguments";
        return 1;
    }
    *first_shell = nshell_total * rank / nranks;
    *nshell = nshell_total * (rank + 1) / nranks - *first_shell;
    return 0;
}

int dftgrid_comm_unique_id(void* id128) {
    return guarded([&] {
        if (!nccl_api().load()) throw std::runtime_error("cannot load libnccl.so.2");
        NcclUniqueId id;
        int rc = nccl_api().GetUniqueId(&id);
        if (rc != 0) throw std::runtime_error(std::string("ncclGetUniqueId: ") + nccl_api().GetErrorString(rc));
        std::memcpy(id128, &id, sizeof id);
    });
}

int dftgrid_comm_init(dftgrid_t* h, const void* id128) {
    return guarded([&] {
        if (!nccl_api().load()) throw std::runtime_error("cannot load libnccl.so.2");
        use_device(h);
        NcclUniqueId id;
        std::memcpy(&id, id128, sizeof id);
        int rc = nccl_api().CommInitRank(&h->comm, h->nranks, id, h->rank);
        if (rc != 0) throw std::runtime_error(std::string("ncclCommInitRank: ") + nccl_api().GetErrorString(rc));
    });
}

int dftgrid_peer_export(dftgrid_t* h, void* handle64) {
    return guarded([&] {
        use_device(h);
        if (!handle64) throw std::runtime_error("null argument");
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
        if (h->nranks < 2 || h->nranks > kPeerMaxRanks) throw std::runtime_error("peer reduction needs 2..16 ranks");
        if (!h->xbuf) {
            const size_t nb2 = (size_t)h->nbf * h->nbf;
            h->xbuf_bytes = kPeerHeaderBytes + 2 * (2 * nb2) * sizeof(double);
            CK(cudaMalloc(&h->xbuf, h->xbuf_bytes));
            CK(cudaMemset(h->xbuf, 0, h->xbuf_bytes));
        }
        cudaIpcMemHandle_t hd;
        CK(cudaIpcGetMemHandle(&hd, h->xbuf));
        std::memcpy(handle64, &hd, sizeof hd);
    });
}

int dftgrid_peer_connect(dftgrid_t* h, const void* handles) {
    return guarded([&] {
        use_device(h);
        if (!handles) throw std::runtime_error("null argument");
        if (!h->xbuf) throw std::runtime_error("dftgrid_peer_export has not been called");
        if (h->peer_ready) return;
        PeerSet ps{};
        ps.nranks = h->nranks;
        ps.rank = h->rank;
        for (int r = 0; r < h->nranks; r++) {
            if (r == h->rank) {
                ps.base[r] = h->xbuf;
                continue;
            }
            cudaIpcMemHandle_t hd;
            std::memcpy(&hd, (const char*)handles + (size_t)r * sizeof hd, sizeof hd);
            void* m = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&m, hd, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                cudaGetLastError();
                for (void* q : h->peer_mapped) cudaIpcCloseMemHandle(q);
                h->peer_mapped.clear();
                throw std::runtime_error(std::string("cudaIpcOpenMemHandle failed (no P2P path between the ranks' GPUs?): ") + cudaGetErrorString(e));
            }
            h->peer_mapped.push_back(m);
            ps.base[r] = (unsigned char*)m;
        }
        h->peers = ps;
        h->peer_ready = true;
    });
}

int dftgrid_peer_active(const dftgrid_t* h) { return h->peer_ready ? 1 : 0; }

int dftgrid_peer_disable(dftgrid_t* h) {
    return guarded([&] {
        use_device(h);
        CK(cudaStreamSynchronize(h->stream));
        h->peer_ready = false;  // back to ncclAllReduce; the mappings stay open until the handle is destroyed
    });
}

int dftgrid_build(dftgrid_t* h) {
    return guarded([&] {
        use_device(h);
        do_build(h);
    });
}

long dftgrid_npoints(const dftgrid_t* h) { return h->g.npts; }
long dftgrid_npoints_local(const dftgrid_t* h) { return h->g.nloc; }
long dftgrid_point_offset(const dftgrid_t* h) { return h->g.shell0 * h->g.nang; }
int dftgrid_nbf(const dftgrid_t* h) { return h->nbf; }
int dftgrid_nlm(const dftgrid_t* h) { return h->g.nlm; }

int dftgrid_upload_density(dftgrid_t* h, const double* P) {
    return guarded([&] {
        use_device(h);
        if (upload_P(h, P)) CK(cudaStreamSynchronize(h->stream));  // the caller may reuse P as soon as this returns
    });
}

int dftgrid_set_density(dftgrid_t* h, const double* P) {
    return guarded([&] {
        use_device(h);
        upload_P(h, P);
        run_density(h);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(h->stream));
    });
}

int dftgrid_hartree_J(dftgrid_t* h, double* J) {
    return guarded([&] {
        use_device(h);
        if (!h->have_density) throw std::runtime_error("dftgrid_set_density has not been called");
        const size_t nb2 = (size_t)h->nbf * h->nbf;
        run_potential(h);
        run_contract(h);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(h->h_res, h->d_res.p, nb2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        check_peer_error(h);
        std::memcpy(J, h->h_res, nb2 * sizeof(double));
    });
}

int dftgrid_xc(dftgrid_t* h, double* XC, double* exc) {
    return guarded([&] {
        use_device(h);
        if (!h->have_density) throw std::runtime_error("dftgrid_set_density has not been called");
        const size_t nb2 = (size_t)h->nbf * h->nbf;
        if (!h->contract_valid) {
            // XC asked for before J: the J half of the fused contraction still needs a defined weight vector
            if (!h->have_potential) h->d_dJ.zero(h->stream);
            run_contract(h);
        }
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(h->h_res, h->d_res.p + nb2, (nb2 + 1) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        check_peer_error(h);
        if (XC) std::memcpy(XC, h->h_res, nb2 * sizeof(double));
        if (exc) *exc = h->h_res[nb2];
    });
}

int dftgrid_electron_count(dftgrid_t* h, double* nelec) {
    return guarded([&] {
        use_device(h);
        if (!h->have_density) throw std::runtime_error("dftgrid_set_density has not been called");
        CK(cudaMemcpyAsync(h->h_res, h->d_scalars.p + 1, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        *nelec = h->h_res[0];
    });
}

int dftgrid_iteration_device(dftgrid_t* h) {
    return guarded([&] {
        use_device(h);
        if (!h->built) throw std::runtime_error("dftgrid_build has not been called");
        static const bool no_graph = std::getenv("DFTGRID_NO_GRAPH") != nullptr;  // developer A/B switch
        const bool want_graph = h->nranks == 1 && !no_graph && !h->graph_failed;
        if (want_graph && h->graph_exec) {
            CK(cudaGraphLaunch(h->graph_exec, h->stream));
            h->launches += h->graph_launches_per_iter;
            h->have_density = h->have_potential = h->contract_valid = h->timed_iter = true;
            return;
        }
        if (want_graph && h->eager_iterations >= 1) {
            // the first iteration ran eagerly (every lazily sized buffer exists now): capture the second one
            const long l0 = h->launches;
            bool ok = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
            if (ok) {
                h->capturing = true;
                try {
                    run_density(h);
                    run_potential(h);
                    run_contract(h);
                } catch (...) {
                    ok = false;
                }
                h->capturing = false;
                cudaGraph_t gcap = nullptr;
                if (cudaStreamEndCapture(h->stream, &gcap) != cudaSuccess || !gcap) ok = false;
                if (ok && cudaGraphInstantiate(&h->graph_exec, gcap, 0) != cudaSuccess) ok = false;
                if (ok) {
                    h->graph = gcap;
                    h->graph_launches_per_iter = h->launches - l0;
                    h->launches = l0;
                } else if (gcap) {
                    cudaGraphDestroy(gcap);
                }
            }
            if (!ok) {
                cudaGetLastError();
                h->graph_exec = nullptr;
                h->graph_failed = true;  // stay on eager launches
                h->launches = l0;
            } else {
                CK(cudaGraphLaunch(h->graph_exec, h->stream));
                h->launches += h->graph_launches_per_iter;
                h->have_density = h->have_potential = h->contract_valid = h->timed_iter = true;
                return;
            }
        }
        run_density(h);
        run_potential(h);
        run_contract(h);
        h->eager_iterations++;
        CK(cudaGetLastError());
    });
}

int dftgrid_download_results(dftgrid_t* h, double* J, double* XC, double* exc, double* nelec) {
    return guarded([&] {
        use_device(h);
        const size_t nb2 = (size_t)h->nbf * h->nbf;
        // page-locked caller buffers receive their matrix straight from the DMA engine; others go through h_res
        const bool dj = J && is_pinned_host(J), dx = XC && is_pinned_host(XC);
        if (dj) CK(cudaMemcpyAsync(J, h->d_res.p, nb2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        if (J && !dj) CK(cudaMemcpyAsync(h->h_res, h->d_res.p, nb2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        if (dx) CK(cudaMemcpyAsync(XC, h->d_res.p + nb2, nb2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        if (XC && !dx) CK(cudaMemcpyAsync(h->h_res + nb2, h->d_res.p + nb2, nb2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(h->h_res + 2 * nb2, h->d_res.p + 2 * nb2, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        check_peer_error(h);
        if (J && !dj) std::memcpy(J, h->h_res, nb2 * sizeof(double));
        if (XC && !dx) std::memcpy(XC, h->h_res + nb2, nb2 * sizeof(double));
        if (exc) *exc = h->h_res[2 * nb2];
        if (nelec) *nelec = h->h_res[2 * nb2 + 1];
    });
}

int dftgrid_iteration(dftgrid_t* h, const double* P, double* J, double* XC, double* exc, double* nelec) {
    // P is consumed by the time the download below has synchronised the stream, so a pinned P is read in place
    int rc = guarded([&] {
        use_device(h);
        upload_P(h, P);
    });
    if (rc) return rc;
    rc = dftgrid_iteration_device(h);
    if (rc) return rc;
    return dftgrid_download_results(h, J, XC, exc, nelec);
}

int dftgrid_synchronize(dftgrid_t* h) {
    return guarded([&] {
        use_device(h);
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaGetLastError());
    });
}

int dftgrid_get_positions(dftgrid_t* h, double* xyz) {
    return guarded([&] {
        const size_t n = (size_t)h->g.nloc;
        std::vector<double> x(n), y(n), z(n);
        download(h, h->d_x, x.data(), n);
        download(h, h->d_y, y.data(), n);
        download(h, h->d_z, z.data(), n);
        for (size_t i = 0; i < n; i++) {
            xyz[3 * i] = x[i];
            xyz[3 * i + 1] = y[i];
            xyz[3 * i + 2] = z[i];
        }
    });
}
int dftgrid_get_weights(dftgrid_t* h, double* w) { return guarded([&] { download(h, h->d_w, w, (size_t)h->g.nloc); }); }
int dftgrid_get_becke_weights(dftgrid_t* h, double* wb) { return guarded([&] { download(h, h->d_wb, wb, (size_t)h->g.nloc); }); }
int dftgrid_get_densities(dftgrid_t* h, double* rho) { return guarded([&] { download(h, h->d_rho, rho, (size_t)h->g.nloc); }); }
int dftgrid_get_potential(dftgrid_t* h, double* V) { return guarded([&] { download(h, h->d_V, V, (size_t)h->g.nloc); }); }
int dftgrid_get_amplitudes(dftgrid_t* h, double* phi) {
    return guarded([&] {
        use_device(h);
        const size_t n = (size_t)h->g.nloc;
        CK(cudaMemcpy2DAsync(phi, (size_t)h->nbf * sizeof(double), h->d_phi.p, (size_t)h->nbp * sizeof(double),
                             (size_t)h->nbf * sizeof(double), n, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    });
}
int dftgrid_get_rho_lm(dftgrid_t* h, double* out) {
    return guarded([&] {
        use_device(h);
        const size_t nshell = (size_t)h->g.natoms * h->g.nrad;
        CK(cudaMemcpyAsync(out, h->d_shell2.p + nshell * 2, nshell * h->g.nlm * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    });
}
int dftgrid_get_U_lm(dftgrid_t* h, double* out) {
    return guarded([&] { download(h, h->d_U_lm, out, (size_t)h->g.natoms * h->g.nrad * h->g.nlm); });
}

int dftgrid_last_timings(dftgrid_t* h, double* out, int n) {
    return guarded([&] {
        use_device(h);
        CK(cudaStreamSynchronize(h->stream));
        if (h->timed_iter) finish_timings(h);
        for (int i = 0; i < n && i < DFTGRID_T_COUNT; i++) out[i] = h->t_ms[i];
    });
}

int dftgrid_timer_start(dftgrid_t* h) {
    return guarded([&] {
        use_device(h);
        CK(cudaEventRecord(h->ev_sw[0], h->stream));
    });
}

int dftgrid_timer_stop(dftgrid_t* h, double* ms) {
    return guarded([&] {
        use_device(h);
        CK(cudaEventRecord(h->ev_sw[1], h->stream));
        CK(cudaEventSynchronize(h->ev_sw[1]));
        float t = 0.f;
        CK(cudaEventElapsedTime(&t, h->ev_sw[0], h->ev_sw[1]));
        *ms = t;
    });
}

long dftgrid_launch_count(const dftgrid_t* h) { return h->launches; }

int dftgrid_debug_contract_schedule(int nbp, long nchunk, int nsm, int max_segs, int* segs_out, int* cta_off_out, int* nctas, int* nsegs,
                                    int* block_chunks) {
    return guarded([&] {
        if (nbp <= 0 || nbp % kNbAlign != 0 || nchunk < 0 || nsm <= 0 || !segs_out || !cta_off_out || !nctas || !nsegs || !block_chunks)
            throw std::runtime_error("dftgrid_debug_contract_schedule: bad arguments");
        ContractSchedule S;
        compute_contract_schedule(nbp, nchunk, nsm, S);
        if ((int)S.segs.size() > max_segs) throw std::runtime_error("dftgrid_debug_contract_schedule: max_segs too small");
        for (size_t i = 0; i < S.segs.size(); i++) {
            segs_out[4 * i + 0] = S.segs[i].z;
            segs_out[4 * i + 1] = S.segs[i].pair;
            std::memcpy(&segs_out[4 * i + 2], &S.segs[i].tb, sizeof(unsigned));
            std::memcpy(&segs_out[4 * i + 3], &S.segs[i].te, sizeof(unsigned));
        }
        for (size_t i = 0; i < S.cta_off.size(); i++) cta_off_out[i] = S.cta_off[i];
        *nctas = (int)S.cta_off.size() - 1;
        *nsegs = (int)S.segs.size();
        *block_chunks = S.bc;
    });
}

}  // extern "C"
