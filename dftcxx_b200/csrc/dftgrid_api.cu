// C ABI (include/dftgrid.h) and host orchestration of the B200 grid engine.
// All compute runs in the hand-written sm_100a kernels of kernels_*.cuh; there is no CPU path.
#include "../../include/dftgrid.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <stdexcept>
#include <string>
#include <vector>

#include "common.cuh"
#include "host_tables.h"
#include "kernels_dense.cuh"
#include "kernels_grid.cuh"
#include "kernels_hartree.cuh"
#include "kernels_integrals.cuh"
#include "kernels_peer.cuh"
#include "kernels_rect.cuh"
#include "kernels_scf.cuh"
#include "nccl_dyn.h"

using namespace dfg;

namespace {

thread_local std::string g_error;

// Developer switches (A/B measurements, tools/dev_*; listed at the end of DESIGN.md) are honoured only when DFTGRID_DEVELOPER
// is set: a stray environment variable cannot change what a production run does.
const char* dev_env(const char* name) { return std::getenv("DFTGRID_DEVELOPER") ? std::getenv(name) : nullptr; }

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

namespace {
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

}  // namespace

struct dftgrid_group;
struct dftgrid {
    int device = 0, rank = 0, nranks = 1;
    cudaStream_t stream = nullptr;
    bool built = false, have_density = false, have_potential = false, contract_valid = false, timed_iter = false;
    int fock_valid = -1;  // contraction mode whose result d_fres holds for the current density, or -1
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
    DevBuf<double> d_atom_xyz, d_Rdist, d_rtab, d_wrad, d_leb, d_Y, d_Yt, d_pre, d_pre_scaled, d_lu, d_xs;
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
    DevBuf<double> d_x, d_y, d_z, d_w, d_wb, d_rho, d_dxc, d_exw, d_V, d_Vown, d_dJ, d_dF, d_phi;
    // device: per iteration
    DevBuf<double> d_P, d_Praw, d_shell_raw, d_shell2, d_qatom, d_qatom2, d_scalars, d_rho_lm, d_U_lm, d_work, d_coef, d_partial, d_res;
    DevBuf<int> d_pairs, d_chunk_ids;
    DevBuf<unsigned long long> d_chunk_mask;  // screening map of the active chunks (k_chunk_masks), aligned with d_chunk_ids
    DevBuf<unsigned long long> d_dbg_times;   // DFTGRID_DEBUG_CTA_TIMES
    DevBuf<int> d_con_chunk_ids;              // the active chunks in the contraction's shuffled sweep order, and their maps
    DevBuf<unsigned long long> d_con_chunk_mask;
    bool screened = false, calibrated = false;
    double screen_work_fraction = 1.0;
    ContractSchedule host_sched;  // host copy of the fused schedule (calibration)
    // stream-K schedules of the contraction: [0] two matrices (XC, J), [1] one matrix (fused Fock build)
    struct DevSchedule {
        DevBuf<ConSeg> segs;
        DevBuf<int> cta_off, item_off;
        int ctas = 1, nsegs = 1, n_single = 0;  // n_single: CTAs with exactly one segment (the ones that keep the soft lockstep)
    } sched[2];
    DevBuf<unsigned> d_con_sync;  // window counters of the contraction's soft lockstep (kernels_dense.cuh: ConSync)
    int con_sync_wc = 0, con_sync_lead = 2, con_sync_pct = 100;
    DevBuf<double> d_fres;  // fused build: [F (nb^2) | per-shell sums of w V rho (natoms*nrad) | e_j, exc, nel, pad]
    DevBuf<double> d_rho_part;  // partial densities when a tile's slabs are split over several CTAs
    int rho_split = 1;
    size_t rho_part_stride = 0;
    int con_bc = 1;
    long n_active_chunks = 0;
    int npairs = 0, interp_chunks = 1;
    DevBuf<double> d_Vpart;
    // binned interpolation (see kernels_hartree.cuh): pairs sorted by (source atom, spline interval) at build time
    bool binned = false;
    int bin_R = 2, bin_nkeys = 0;  // pairs per lane of k_interp_bin (3 with the precomputed pair geometry)
    long bin_nitems = 0;
    DevBuf<int> d_binoff, d_item_key, d_pair_point, d_slot_of;
    DevBuf<double> d_pair_out;
    DevBuf<double> d_pair_geo;  // [6][slots]: 1/r, t, cos / sin(theta), cos / sin(phi) of every pair (k_pair_geometry), when it fits

    // pinned staging
    double* h_P = nullptr;
    double* h_res = nullptr;

    // comm
    NcclComm comm = nullptr;
    // peer-memory reduction of [J | XC] (kernels_peer.cuh)
    unsigned char* xbuf = nullptr;        // this rank's exchange buffer
    size_t xbuf_bytes = 0, xslot = 0;
    PeerSet peers{};
    std::vector<void*> peer_mapped;       // cudaIpcOpenMemHandle results to close
    bool peer_ready = false, peer_used = false, peer_local = false;  // peer_local: single-process group (no IPC mappings)
    unsigned long long peer_epoch = 0;

    // CUDA graph of one whole iteration (single-GPU handles): the 15 launches of a small molecule's iteration are
    // launch-latency bound, one graph launch replaces them from the second call of dftgrid_iteration_device on
    struct IterGraph {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        bool failed = false;
        long eager = 0, launches_per_iter = 0;
        void reset() {
            if (exec) cudaGraphExecDestroy(exec);
            if (graph) cudaGraphDestroy(graph);
            graph = nullptr;
            exec = nullptr;
            failed = false;
            eager = 0;
        }
    } graphs[3];  // per contraction mode (kModePair, kModeFock, kModeFockJ)
    bool capturing = false;

    // single-process multi-GPU: a group handle holds no device state of its own, only its rank handles
    dftgrid_group* group = nullptr;

    // device-resident SCF algebra (kernels_scf.cuh): H, X, X^T, work matrices [np][np], purification state
    struct ScfState {
        bool ready = false;
        int np = 0, nocc = 0;
        long steps = 0;
        double alpha = 0.5;
        DevBuf<double> H, X, Xt, F, T1, Fp, D, D2, D3, lo, hi, diag, rows, eone;
        DevBuf<PmState> pm;
        PmState* h_pm = nullptr;  // pinned
        double* h_out = nullptr;  // pinned [8]
        cudaEvent_t ev[3]{};
    } scf;

    // timing
    cudaEvent_t ev[16]{};
    cudaEvent_t ev_sw[2]{};
    bool ev_build = false, ev_iter = false;
    double t_ms[DFTGRID_T_COUNT]{};

    ~dftgrid() {
        for (auto& G : graphs) G.reset();
        for (void* m : peer_mapped) cudaIpcCloseMemHandle(m);
        if (xbuf) cudaFree(xbuf);
        if (comm && nccl_api().ok) nccl_api().CommDestroy(comm);
        for (auto& e : ev)
            if (e) cudaEventDestroy(e);
        for (auto& e : ev_sw)
            if (e) cudaEventDestroy(e);
        if (scf.h_pm) cudaFreeHost(scf.h_pm);
        if (scf.h_out) cudaFreeHost(scf.h_out);
        for (auto& e : scf.ev)
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
        if (dev_env("DFTGRID_DEBUG_SKIP_COMM")) return;
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

// Pure host arithmetic (no device): also exported as dftgrid_debug_contract_schedule for the CPU test-suite.
// item_frac (optional, [npairs]): mean fraction of a full stage's DMMA work that a chunk costs the tile pair under the
// screening map (1 = every block of every chunk significant); the items' costs are scaled by it so that the CTAs' shares
// stay equal in time.
void compute_contract_schedule(int nbp, long nchunk, int nsm, int nz, ContractSchedule& S, const std::vector<double>* item_frac = nullptr,
                               const std::vector<double>* pair_cost = nullptr) {
    S = ContractSchedule();
    const int nt = (nbp + kTileM - 1) / kTileM;
    for (int i = 0; i < nt; i++)
        for (int j = i; j < nt; j++) {
            S.pairs.push_back(i);
            S.pairs.push_back(j);
        }
    S.npairs = (int)S.pairs.size() / 2;
    const int npairs = S.npairs, nitems = nz * npairs;
    std::vector<double> cost(nitems);
    static const char* nc = dev_env("DFTGRID_NARROW_COST");  // developer sweeps of the cost model
    static const char* dc = dev_env("DFTGRID_DIAG_COST");
    static const char* ec = dev_env("DFTGRID_EDGE_DIAG_COST");
    static const char* l2e = dev_env("DFTGRID_L2_BLOCK_MB");
    static const char* n32c = dev_env("DFTGRID_N32_COST");
    static const char* d32c = dev_env("DFTGRID_D32_COST");
    double W1 = 0.0;  // cost of all items for one chunk
    for (int it = 0; it < nitems; it++) {
        const int ti = S.pairs[2 * (it % npairs)], tj = S.pairs[2 * (it % npairs) + 1];
        // relative cost of one k-chunk of this tile pair (full off-diagonal 128x128 tile = 20), calibrated by sweeps at
        // (H2O)64, (H2O)32 and C40H82: a 64-wide edge tile issues half the DMMAs but pays the same loads (10.5), a
        // diagonal tile issues 17 of 32 DMMAs per warp (11.5), the 64-wide diagonal tile 8 of 32 (6.5; under-estimating
        // it makes its CTA the straggler, 5 costs 14 %)
        // a 32-wide edge tile (nb = 524 -> nbp = 544) issues a quarter of the DMMAs for the same A loads (6.5); its
        // diagonal tile keeps only four warps busy with 4 dependent DMMAs per k4-step (5; sweeps at C40H82: 3 costs 35 %)
        const int wj = std::min(kTileN, nbp - tj * kTileN);
        const bool narrow = wj <= 64, narrow32 = wj <= 32;
        const double c_narrow = nc ? std::atof(nc) : 10.5, c_diag = dc ? std::atof(dc) : 11.5, c_edge_diag = ec ? std::atof(ec) : 6.5;
        const double c_n32 = n32c ? std::atof(n32c) : 6.5, c_d32 = d32c ? std::atof(d32c) : 5.0;
        cost[it] = ti == tj ? (narrow32 ? c_d32 : narrow ? c_edge_diag : c_diag) : (narrow32 ? c_n32 : narrow ? c_narrow : 20.0);
        if (item_frac) cost[it] *= std::max(0.02, (*item_frac)[it % npairs]);  // never zero: every item keeps a segment
        if (pair_cost) cost[it] = std::max(1e-3, (*pair_cost)[it % npairs]);      // measured costs (calibrate_contract_costs)
        W1 += cost[it];
    }
    // block length (the period at which a CTA with several segments alternates between them): ~120 MB of Phi rows.
    // Measured at (H2O)64: DRAM reads 15.1 GB without blocks, 9.8 GB with 80-160 MB blocks, 10.3 GB at 40 MB where the
    // accumulator parking starts to cost time; at least 64 chunks; one block when the shard is smaller.
    {
        const double chunk_bytes = (double)kTileK * nbp * sizeof(double);
        double l2_mb = 120.0;
        if (l2e) l2_mb = std::atof(l2e);  // developer sweep; <= 0: one block
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

void build_contract_schedule(dftgrid* h, long nchunk, int nsm, const std::vector<double>* item_frac, const std::vector<double>* pair_cost = nullptr) {
    cudaStream_t st = h->stream;
    size_t max_segs = 0;
    for (int k = 0; k < 2; k++) {
        ContractSchedule S;
        compute_contract_schedule(h->nbp, nchunk, nsm, k == 0 ? 2 : 1, S, item_frac, pair_cost);
        if (k == 1) h->host_sched = S;  // kept for the calibration pass
        h->npairs = S.npairs;
        h->con_bc = S.bc;
        if (k == 0) h->d_pairs.upload(S.pairs, st);
        dftgrid::DevSchedule& D = h->sched[k];
        D.ctas = (int)S.cta_off.size() - 1;
        D.nsegs = (int)S.segs.size();
        D.n_single = 0;
        for (int c = 0; c < D.ctas; c++) D.n_single += S.cta_off[c + 1] - S.cta_off[c] == 1 ? 1 : 0;
        D.segs.alloc(S.segs.size());
        CK(cudaMemcpyAsync(D.segs.p, S.segs.data(), S.segs.size() * sizeof(ConSeg), cudaMemcpyHostToDevice, st));
        D.cta_off.upload(S.cta_off, st);
        D.item_off.upload(S.item_off, st);
        max_segs = std::max(max_segs, S.segs.size());
        if (dev_env("DFTGRID_DEBUG_CTA_TIMES") && k == 1)  // developer instrumentation: the fused schedule's segments per CTA
            for (int c = 0; c + 1 < (int)S.cta_off.size(); c++) {
                std::fprintf(stderr, "[dftgrid] sched cta %3d:", c);
                for (int q = S.cta_off[c]; q < S.cta_off[c + 1]; q++)
                    std::fprintf(stderr, " (pair %d [%d,%d] share %.3f)", S.segs[q].pair, S.pairs[2 * S.segs[q].pair], S.pairs[2 * S.segs[q].pair + 1],
                                 (S.segs[q].te - S.segs[q].tb) / 2147483648.0);
                std::fprintf(stderr, "\n");
            }
        CK(cudaStreamSynchronize(st));  // S goes out of scope
    }
    h->d_partial.alloc(max_segs * kTileM * kTileN);
    {
        // soft lockstep of the sweep (kernels_dense.cuh: ConSync): windows of 24 MB of Phi rows, at most 2 windows of lead
        // (developer switches DFTGRID_CON_SYNC_MB, 0 = off, _LEAD, _PCT); only when the sweep has enough windows.  Measured at
        // (H2O)64 on the calibrated shares: DRAM reads 7.0 -> 4.7 GB per launch (1.26x the algorithmic 3.73 GB), 8.73 -> 8.62 ms
        static const char* mb = dev_env("DFTGRID_CON_SYNC_MB");
        static const char* lead = dev_env("DFTGRID_CON_SYNC_LEAD");
        static const char* pct = dev_env("DFTGRID_CON_SYNC_PCT");
        const double sync_mb = mb ? std::atof(mb) : 24.0;
        const double chunk_bytes = (double)kTileK * h->nbp * sizeof(double);
        h->con_sync_wc = 0;
        if (sync_mb > 0.0) {
            const long wc = std::max<long>(32, (long)(sync_mb * 1e6 / chunk_bytes));
            if (nchunk >= 8 * wc) {
                h->con_sync_wc = (int)wc;
                h->con_sync_lead = lead ? std::max(1, std::atoi(lead)) : 2;
                h->con_sync_pct = pct ? std::min(100, std::max(1, std::atoi(pct))) : 100;
                h->d_con_sync.alloc((size_t)((nchunk + wc - 1) / wc) + 1);
            }
        }
    }
}

// Measured cost model of the screened contraction.  The analytic weights (DMMA counts per tile kind x the map's work
// fractions) miss what the map does to the memory side: a tile pair that is significant for few chunks is read by few
// CTAs, its Phi rows come from DRAM instead of the L2 and its stages run latency-bound (measured on a 1/8 shard of
// (H2O)64: 2.5 us per diagonal stage on near tiles, 3-4.3 us on far ones; CTAs with equal modelled shares differed by
// 60 %).  So the schedule is calibrated on the device it will run on: one instrumented launch of the fused contraction
// (per CTA: wall time and stages consumed), the time per staged chunk of every tile pair from the CTAs that hold a single
// segment, and the stream-K shares are cut again on cost(pair) = P(chunk staged for the pair) x measured time per stage.
// The launch costs as much as one contraction, once per grid.  Results do not depend on it (any schedule sums the same
// partial tiles in a fixed order); DFTGRID_NO_CALIBRATE keeps the analytic model.
void calibrate_contract_costs(dftgrid* h, int nsm, const std::vector<double>& sig_frac, const std::vector<double>& item_frac) {
    if (dev_env("DFTGRID_NO_CALIBRATE") || h->n_active_chunks < 4L * nsm) return;
    cudaStream_t st = h->stream;
    const bool verbose = dev_env("DFTGRID_DEBUG_CTA_TIMES") != nullptr;
    // one instrumented run of the CURRENT fused schedule: kernel span in us, per-CTA (duration us, stages)
    auto measure = [&](std::vector<double>& us, std::vector<double>& stages) -> double {
        const dftgrid::DevSchedule& D = h->sched[1];
        DevBuf<unsigned long long> d_times;
        d_times.alloc(3 * (size_t)D.ctas);
        std::vector<unsigned long long> t(3 * (size_t)D.ctas);
        for (int rep = 0; rep < 2; rep++)  // the second launch runs at steady clocks with the schedule's own L2 pattern
            k_contract_tma<<<D.ctas, kConTmaThreads, kConTmaSmemBytes, st>>>(h->d_phi.p, h->d_dF.p, h->d_dF.p, h->d_con_chunk_ids.p,
                                                                             h->d_con_chunk_mask.p, h->d_pairs.p, D.segs.p, D.cta_off.p, h->d_partial.p,
                                                                             h->nbp, (int)h->n_active_chunks, h->con_bc, d_times.p);
        h->launches += 2;
        CK(cudaMemcpyAsync(t.data(), d_times.p, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        unsigned long long t0 = ~0ull, t1 = 0ull;
        us.assign(D.ctas, 0.0);
        stages.assign(D.ctas, 0.0);
        for (int c = 0; c < D.ctas; c++) {
            t0 = std::min(t0, t[3 * c]);
            t1 = std::max(t1, t[3 * c + 1]);
            us[c] = (double)(t[3 * c + 1] - t[3 * c]) * 1e-3;
            stages[c] = (double)(t[3 * c + 2] & 0xffffffffull);
        }
        return (double)(t1 - t0) * 1e-3;
    };
    auto base_cost = [&](const ContractSchedule& S, int p) {  // the analytic constants of compute_contract_schedule
        const int ti = S.pairs[2 * p], tj = S.pairs[2 * p + 1];
        const int wj = std::min(kTileN, h->nbp - tj * kTileN);
        const bool narrow = wj <= 64, narrow32 = wj <= 32;
        return ti == tj ? (narrow32 ? 5.0 : narrow ? 6.5 : 11.5) : (narrow32 ? 6.5 : narrow ? 10.5 : 20.0);
    };
    // costs from a measurement of the current schedule; empty when nothing reliable was measured
    auto costs_from = [&](const std::vector<double>& us, const std::vector<double>& stages) {
        const ContractSchedule& S = h->host_sched;
        const int npairs = S.npairs;
        std::vector<std::vector<double>> samples(npairs);
        for (int c = 0; c + 1 < (int)S.cta_off.size(); c++)
            if (S.cta_off[c + 1] - S.cta_off[c] == 1 && stages[c] >= 24.0 && us[c] > 0.0) samples[S.segs[S.cta_off[c]].pair].push_back(us[c] / stages[c]);
        std::vector<double> per_stage(npairs, 0.0), analytic(npairs), ratios;
        for (int p = 0; p < npairs; p++) {
            analytic[p] = item_frac[p] / std::max(sig_frac[p], 1e-6);  // mean masked fraction of a staged chunk
            if (!samples[p].empty()) {
                std::sort(samples[p].begin(), samples[p].end());
                per_stage[p] = samples[p][samples[p].size() / 2];
                ratios.push_back(per_stage[p] / (base_cost(S, p) * analytic[p]));
            }
        }
        std::vector<double> pair_cost;
        if (ratios.size() < 2) return pair_cost;
        std::sort(ratios.begin(), ratios.end());
        const double r = ratios[ratios.size() / 2];  // unmeasured pairs: the analytic cost at the median measured / analytic ratio
        pair_cost.resize(npairs);
        for (int p = 0; p < npairs; p++) pair_cost[p] = sig_frac[p] * (per_stage[p] > 0.0 ? per_stage[p] : r * base_cost(S, p) * analytic[p]);
        if (verbose)
            for (int p = 0; p < npairs; p++)
                std::fprintf(stderr, "[dftgrid] calibrated pair %d [%d,%d]: %.2f us/stage (%zu samples), staged %.3f\n", p, S.pairs[2 * p], S.pairs[2 * p + 1],
                             per_stage[p], samples[p].size(), sig_frac[p]);
        return pair_cost;
    };
    // Autotuning loop: the analytic schedule, then up to three schedules cut on what the previous one measured.  The fastest
    // measured-cost schedule is kept unless it is more than 1 % SLOWER than the analytic one: at equal speed the measured
    // costs are still the better choice, because CTAs whose shares match their real pace sweep the Phi rows in step and
    // share them through the L2 (ncu at (H2O)64: 7.0 GB of DRAM reads per launch, L2 hit rate 53 %, against 13.1 GB / 28 %
    // for the analytic shares at 8.4 vs 8.6 ms — profiles/r02b_contract_dram_per_launch.txt).
    std::vector<double> us, stages, best_cost, cur_cost;
    const double t_analytic = measure(us, stages);
    double best = 1.01 * t_analytic;
    if (verbose) std::fprintf(stderr, "[dftgrid] contraction schedule, analytic costs: %.1f us\n", t_analytic);
    for (int round = 0; round < 3; round++) {
        cur_cost = costs_from(us, stages);
        if (cur_cost.empty()) break;
        build_contract_schedule(h, h->n_active_chunks, nsm, nullptr, &cur_cost);
        const double tcur = measure(us, stages);
        if (verbose) std::fprintf(stderr, "[dftgrid] contraction schedule, measured costs (round %d): %.1f us\n", round + 1, tcur);
        if (tcur < best) {
            best = tcur;
            best_cost = cur_cost;
        }
    }
    if (best_cost.empty())
        build_contract_schedule(h, h->n_active_chunks, nsm, &item_frac);
    else
        build_contract_schedule(h, h->n_active_chunks, nsm, nullptr, &best_cost);
    h->calibrated = !best_cost.empty();
}

// Lists of the 32-point chunks / 128-point tiles of Phi that hold any non-zero amplitude (k_chunk_flags), and the
// contraction schedule over the non-zero chunks.
void build_active_lists(dftgrid* h, int nsm) {
    const GridShape& g = h->g;
    cudaStream_t st = h->stream;
    const long nchunk = (g.nloc + kTileK - 1) / kTileK;
    const int nblk = h->nbp / 32;
    // Screening map (k_chunk_masks).  DFTGRID_SCREEN_TAU: threshold on |phi| (default 1e-20; 0 = exact zeros only; negative
    // or DFTGRID_NO_ZERO_SKIP = no skipping at all, developer A/B switches).  More than 64 column blocks: no map.
    double tau = 1e-20;
    if (const char* e = dev_env("DFTGRID_SCREEN_TAU")) tau = std::atof(e);
    const bool screen = nchunk > 0 && tau >= 0.0 && nblk <= 64 && !dev_env("DFTGRID_NO_ZERO_SKIP");
    std::vector<unsigned long long> masks((size_t)nchunk, ~0ull);
    if (screen) {
        DevBuf<unsigned long long> d_masks;
        d_masks.alloc((size_t)nchunk);
        k_chunk_masks<<<(unsigned)((nchunk * 32 + 255) / 256), 256, 0, st>>>(h->d_phi.p, nchunk, h->nbp, tau, d_masks.p);
        h->launches++;
        CK(cudaMemcpyAsync(masks.data(), d_masks.p, (size_t)nchunk * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    std::vector<int> flags((size_t)nchunk, 1);
    for (long c = 0; c < nchunk; c++) flags[c] = masks[c] != 0ull;
    std::vector<int> chunk_ids;
    for (long c = 0; c < nchunk; c++)
        if (flags[c]) chunk_ids.push_back((int)c);
    h->n_active_chunks = (long)chunk_ids.size();
    std::vector<unsigned long long> act_masks;
    for (int c : chunk_ids) act_masks.push_back(masks[c]);
    while (chunk_ids.size() % 4 != 0 || chunk_ids.empty()) {  // k_rho_tma reads groups of four
        chunk_ids.push_back(-1);
        act_masks.push_back(0ull);
    }
    h->d_chunk_ids.upload(chunk_ids, st);
    h->screened = screen;
    if (screen) h->d_chunk_mask.upload(act_masks, st);
    CK(cudaStreamSynchronize(st));
    // The contraction sweeps the active chunks in a pseudo-randomly shuffled order when the map is in use: any window of
    // positions then holds chunks of all atoms and radii, every tile pair pays its AVERAGE cost per window, and the CTAs
    // (whose shares are equalised on those averages) keep passing over the Phi rows in step, which is what lets a row's
    // ~14 uses hit the L2.  (In atom-major order a pair is cheap far from its atoms and expensive near them; the CTAs
    // drift apart by thousands of chunks and DRAM reads triple: measured 6.5 -> 18 GB.)
    std::vector<double> item_frac, sig_frac;
    if (screen && h->n_active_chunks > 0) {
        const long n = h->n_active_chunks;
        // (an integer mixing hash, NOT another golden-ratio sequence: the device's ownership hash of the POSITIONS is one, and
        // the composition of two of them maps runs of neighbouring chunks — one atom's shells, exactly the chunks that are
        // significant for a far tile pair — onto arithmetic progressions of owners: CTAs with equal shares of one item
        // then differed by 60 % in work)
        std::vector<std::pair<unsigned, int>> key((size_t)n);
        for (long k = 0; k < n; k++) {
            unsigned v = (unsigned)k * 0x9E3779B1u + 0x7F4A7C15u;
            v ^= v >> 16;
            v *= 0x85EBCA6Bu;
            v ^= v >> 13;
            v *= 0xC2B2AE35u;
            v ^= v >> 16;
            key[k] = {v, (int)k};
        }
        std::sort(key.begin(), key.end());
        std::vector<int> con_ids((size_t)n);
        std::vector<unsigned long long> con_masks((size_t)n);
        for (long t = 0; t < n; t++) {
            con_ids[t] = chunk_ids[key[t].second];
            con_masks[t] = act_masks[key[t].second];
        }
        h->d_con_chunk_ids.upload(con_ids, st);
        h->d_con_chunk_mask.upload(con_masks, st);
        CK(cudaStreamSynchronize(st));
        // work fraction of every tile pair under the map: a stage costs the pair as much as its busiest DMMA warp
        // (kernels_dense.cuh ConMode*::mma), a chunk with an insignificant tile is not staged at all
        // A staged chunk never costs less than the pipeline's own turnaround (the producers' ~33 bulk copies each, the
        // barrier round trip): a floor on the per-stage fraction, calibrated by sweeps (DFTGRID_STAGE_FLOOR).
        double floor_ = 0.25;
        if (const char* e = dev_env("DFTGRID_STAGE_FLOOR")) floor_ = std::atof(e);
        const int nt = (h->nbp + kTileM - 1) / kTileM;
        for (int ti = 0; ti < nt; ti++)
            for (int tj = ti; tj < nt; tj++) {
                const int nbj = std::min(4, nblk - 4 * tj);
                double acc = 0.0;
                long staged = 0;
                for (long x = 0; x < n; x++) {
                    const unsigned long long cm = act_masks[x];
                    const unsigned ab = (unsigned)(cm >> (4 * ti)) & 0xFu, bb = (unsigned)(cm >> (4 * tj)) & ((1u << nbj) - 1u);
                    if (!ab || !bb) continue;
                    staged++;
                    if (ti != tj) {
                        acc += std::max(floor_, (double)__builtin_popcount(bb) / nbj);
                    } else if (nbj == 4) {  // triangular diagonal tile: staged chunks run the full stage
                        acc += 1.0;
                    } else {
                        acc += std::max(floor_, (double)__builtin_popcount(bb) / nbj);
                    }
                }
                item_frac.push_back(acc / (double)n);
                sig_frac.push_back((double)staged / (double)n);
            }
        double mean = 0.0;
        for (double f : item_frac) mean += f;
        h->screen_work_fraction = item_frac.empty() ? 1.0 : mean / item_frac.size();
    }
    {
        // k_rho_tma work items: one CTA per 128-point tile when that gives many waves over the SMs; with few waves
        // (sharded grids, small molecules) the tail wave costs up to 1/waves, so a tile's column slabs are dealt to 2 or 3 CTAs
        const double waves = (double)((h->n_active_chunks + 3) / 4) / (double)nsm;
        const int nslab = (h->nbp + kTileN - 1) / kTileN;
        int split = waves >= 40.0 ? 1 : (waves >= 12.0 ? 2 : 3);  // measured at (H2O)64: 28 waves 11.07 -> 10.98 ms (2 CTAs), 3.5 waves 1.56 -> 1.46 ms (3 CTAs)
        if (const char* e = dev_env("DFTGRID_RHO_SPLIT")) split = std::atoi(e);  // developer A/B switch
        h->rho_split = std::max(1, std::min(split, std::min(nslab, 3)));
        h->rho_part_stride = (size_t)g.nloc + 64;
        if (h->rho_split > 1) {
            h->d_rho_part.alloc((size_t)h->rho_split * h->rho_part_stride);
            h->d_rho_part.zero(st);  // skipped (all-zero) chunks are never written
        }
    }
    build_contract_schedule(h, h->n_active_chunks, nsm, item_frac.empty() ? nullptr : &item_frac);
    if (!item_frac.empty()) calibrate_contract_costs(h, nsm, sig_frac, item_frac);
    if (dev_env("DFTGRID_DEBUG_CTA_TIMES")) {
        h->d_dbg_times.alloc(3 * (size_t)nsm + 3);
        for (size_t i = 0; i < item_frac.size(); i++) std::fprintf(stderr, "[dftgrid] pair %zu work fraction %.3f\n", i, item_frac[i]);
    }
}

// Sort the (local point, source atom) pairs of the cross-atom interpolation into (atom, spline interval) bins.
// Geometry only, so it is done once per grid.  Falls back to the point-parallel kernels when lmax is not one of the
// presets, the pair count does not fit 32-bit slots, or the lists do not fit in free device memory.
void build_pair_bins(dftgrid* h) {
    const GridShape& g = h->g;
    cudaStream_t st = h->stream;
    h->binned = false;
    if (dev_env("DFTGRID_INTERP_POINTWISE")) return;  // developer A/B switch
    if (!(g.lmax == 5 || g.lmax == 8 || g.lmax == 10 || g.lmax == 11)) return;
    if (g.nloc == 0 || g.natoms < 2) return;
    const int nkeys = g.natoms * g.nrad;
    size_t mem_free = 0, mem_total = 0;
    CK(cudaMemGetInfo(&mem_free, &mem_total));
    // pairs per lane: 3 when the pairs' geometry can be kept (k_pair_geometry: the kernel then has the registers for a third
    // pair per lane, measured 5.73 -> 5.48 ms at (H2O)64), 2 otherwise; DFTGRID_INTERP_R overrides (developer switch)
    static const bool no_geo = dev_env("DFTGRID_NO_PAIR_GEO") != nullptr;  // developer A/B switch
    const double pairs_est = (double)g.nloc * (g.natoms - 1) + (double)nkeys * 96.0;
    const double lists = pairs_est * 12.0 + (double)g.nloc * g.natoms * 4.0;
    const bool want_geo = !no_geo && lists + pairs_est * 48.0 < 0.8 * (double)mem_free;  // nothing large is allocated after this
    h->bin_R = want_geo ? 3 : 2;
    if (const char* r = dev_env("DFTGRID_INTERP_R")) h->bin_R = std::atoi(r);
    if (h->bin_R < 2 || h->bin_R > 4) h->bin_R = 2;
    const int unit = 32 * h->bin_R;
    const double pairs_max = (double)g.nloc * (g.natoms - 1) + (double)nkeys * unit;
    if (pairs_max >= 2.0e9) return;
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
    // the pairs' geometry, once (48 bytes per slot), when it fits comfortably beside what the iteration still allocates
    h->d_pair_geo.release();
    {
        CK(cudaMemGetInfo(&mem_free, &mem_total));
        const double need = 6.0 * 8.0 * (double)total;
        if (want_geo && total > 0 && need < 0.85 * (double)mem_free) {
            h->d_pair_geo.alloc(6 * (size_t)total);
            k_pair_geometry<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(g, h->d_atom_xyz.p, h->d_x.p, h->d_y.p, h->d_z.p, h->d_xs.p, h->d_item_key.p,
                                                                              h->d_pair_point.p, unit, total, h->d_pair_geo.p);
            h->launches++;
        }
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));  // binoff (host vector) and the scratch buffers go out of scope
    h->binned = true;
}

// This rank's contiguous block of the (atom, radial shell) units.  With screening the tensor work of a shell depends on
// how many 32-column blocks of Phi are significant on it (a shell far from most atoms is cheap), so equal shell COUNTS
// (dftgrid_shard_range) leave the ranks up to ~20 % out of balance; the blocks are therefore cut at equal estimated WORK:
//   w(shell) = t_dense (f^2 [rho] + f [contraction]) + t_interp [interpolation],  f = significant blocks / blocks
// where a block counts as significant on shell (a, r) when some basis function f in it has | |R_a - R_f| - r | < reach_f,
// reach_f = the distance beyond which |phi_f| <= tau by the bound sum_k |c_k N_k| d^L exp(-alpha_k d^2).  Pure host
// arithmetic on replicated data: every rank computes the same cuts.  Single rank, or screening off: the plain rule.
void shard_shells(dftgrid* h, const std::vector<double>& rtab, long nshell, long* first, long* count) {
    double tau = 1e-20;
    if (const char* e = dev_env("DFTGRID_SCREEN_TAU")) tau = std::atof(e);
    const int nblk = h->nbp / 32;
    if (h->nranks == 1 || tau < 0.0 || nblk > 64 || dev_env("DFTGRID_NO_ZERO_SKIP") || dev_env("DFTGRID_EQUAL_SHARDS")) {
        dftgrid_shard_range(nshell, h->rank, h->nranks, first, count);
        return;
    }
    const double tau_eff = std::max(tau, 1e-300);
    std::vector<double> reach(h->nbf, 0.0);
    for (int f = 0; f < h->nbf; f++) {
        double rf = 0.25;
        for (double d = 0.25; d <= 80.0; d += 0.25) {
            double b = 0.0;
            for (int k = h->bf_prim_off[f]; k < h->bf_prim_off[f + 1]; k++) {
                const int code = h->prim_lmn[k], L = (code & 15) + ((code >> 4) & 15) + ((code >> 8) & 15);
                b += std::fabs(h->prim_coeff[k] * h->prim_norm[k]) * std::pow(d, L) * std::exp(-h->exp_alpha[h->prim_exp[k]] * d * d);
            }
            if (b > tau_eff) rf = d + 0.25;
        }
        reach[f] = rf;
    }
    const int nrad = h->prm.radial_points, na = h->natoms;
    // per-point cost model (ns, B200): the density kernel executes nb^2 DMMA flops per point on the significant blocks
    // (~ fsig^2), the fused contraction nb (nb + 1) of which screening removes less (masked stages have a floor, diagonal
    // tiles run whole: ~ fsig, measured on the shards of (H2O)64), the interpolation 6.2 FP64 operations per
    // (point, source atom, lm) at ~70 % of the FP64 pipe whatever the map says (5.5 ms per 1.3e10 terms with the pairs'
    // geometry precomputed)
    const int lmax = h->prm.lmax, nlm = (lmax + 1) * (lmax + 1);
    const double t_dense = (double)h->nbf * h->nbf / 37.0e3, t_interp = (double)(na - 1) * nlm * 6.2 / 14.7e3;
    std::vector<double> cum((size_t)nshell + 1, 0.0);
    std::vector<double> dist(h->nbf);
    for (int a = 0; a < na; a++) {
        for (int f = 0; f < h->nbf; f++) {
            const double* c = &h->center_xyz[3 * (size_t)h->bf_center[f]];
            const double dx = c[0] - h->atom_xyz[3 * a], dy = c[1] - h->atom_xyz[3 * a + 1], dz = c[2] - h->atom_xyz[3 * a + 2];
            dist[f] = std::sqrt(dx * dx + dy * dy + dz * dz);
        }
        for (int i = 0; i < nrad; i++) {
            int nsig = 0;
            for (int b = 0; b < nblk; b++) {
                bool sig = false;
                for (int f = 32 * b; f < std::min(h->nbf, 32 * b + 32) && !sig; f++) sig = std::fabs(dist[f] - rtab[i]) < reach[f];
                nsig += sig ? 1 : 0;
            }
            const double fsig = (double)nsig / nblk;
            cum[(size_t)a * nrad + i + 1] = cum[(size_t)a * nrad + i] + t_dense * (fsig * fsig + fsig) + t_interp;
        }
    }
    const double total = cum[nshell];
    auto cut = [&](int r) -> long {  // first shell whose cumulative work reaches r/nranks of the total
        if (r <= 0) return 0;
        if (r >= h->nranks) return nshell;
        const double target = total * r / h->nranks;
        return (long)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());
    };
    const long lo = cut(h->rank), hi = std::max(lo, cut(h->rank + 1));
    *first = std::min(lo, nshell);
    *count = std::min(hi, nshell) - *first;
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
    h->leb_off = lebedev_offset(prm.lebedev_order);
    // developer instrumentation (DFTGRID_DEBUG_BUILD_TIMES): wall clock of the build's host-side milestones
    static const bool build_times = dev_env("DFTGRID_DEBUG_BUILD_TIMES") != nullptr;
    auto t_mark = std::chrono::steady_clock::now();
    auto mark = [&](const char* what) {
        if (!build_times) return;
        CK(cudaStreamSynchronize(st));
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[dftgrid] build rank %d: %-28s %8.1f ms\n", h->rank, what, std::chrono::duration<double, std::milli>(now - t_mark).count());
        t_mark = now;
    };

    // ---- host tables
    std::vector<double> r, wr, Y, pre;
    make_radial(g.nrad, r, wr);
    shard_shells(h, r, nshell, &g.shell0, &g.nshell_loc);
    g.nloc = g.nshell_loc * g.nang;
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

    mark("host tables");
    h->d_atom_xyz.upload(h->atom_xyz, st);
    h->d_Rdist.upload(Rdist, st);
    h->d_rtab.upload(r, st);
    h->d_wrad.upload(wr, st);
    h->d_leb.upload(leb, st);
    h->d_Y.upload(Y, st);
    h->d_Yt.upload(Yt, st);
    h->d_pre.upload(pre, st);
    {
        // prefactors times the scale of k_interp_bin's Legendre recurrence (kernels_hartree.cuh: legendre_scale)
        std::vector<double> pre_s(pre.size());
        for (int l = 0; l <= g.lmax; l++)
            for (int m = 0; m <= g.lmax; m++) pre_s[(size_t)l * (g.lmax + 1) + m] = pre[(size_t)l * (g.lmax + 1) + m] * (m <= l ? legendre_scale(l, m) : 1.0);
        h->d_pre_scaled.upload(pre_s, st);
    }
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

    mark("table uploads");
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
    h->d_dF.alloc(nlp);
    h->d_dF.zero(st);
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
    h->d_qatom2.alloc(g.natoms);
    h->d_fres.alloc((size_t)h->nbf * h->nbf + (size_t)nshell + 4);
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
    if (!h->h_res) CK(cudaMallocHost(&h->h_res, sizeof(double) * ((size_t)2 * h->nbf * h->nbf + 4)));

    CK(cudaFuncSetAttribute(k_phi, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)(((size_t)kPhiPts * (kPhiCols + 1) + (size_t)kPhiMaxExp * kPhiPts) * sizeof(double))));
    CK(cudaFuncSetAttribute(k_rho_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRhoTmaSmemBytes));
    CK(cudaFuncSetAttribute(k_contract_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kConTmaSmemBytes));
    const size_t becke_smem = (size_t)kBeckeWarps * 2 * g.natoms * sizeof(double);
    if (becke_smem > 200 * 1024) throw std::runtime_error("too many atoms for the Becke kernel's shared-memory layout");
    CK(cudaFuncSetAttribute(k_becke, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(becke_smem, 1024)));

    mark("allocations + attributes");
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
            g, h->d_atom_xyz.p, h->d_Rdist.p, h->d_x.p, h->d_y.p, h->d_z.p, h->d_w.p, h->d_wb.p);
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
    mark("points, Becke, Phi");
    build_active_lists(h, nsm);
    mark("block map, schedule, calibration");
    build_pair_bins(h);
    mark("interpolation pair bins");
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
            const unsigned long long* cmask = h->screened ? h->d_chunk_mask.p : nullptr;
            if (h->rho_split == 1) {
                k_rho_tma<<<tiles, kRhoTmaThreads, kRhoTmaSmemBytes, st>>>(h->d_phi.p, h->d_P.p, h->d_chunk_ids.p, cmask, h->d_rho.p, 0, g.nloc, h->nbp);
            } else {
                k_rho_tma<<<dim3(tiles, h->rho_split), kRhoTmaThreads, kRhoTmaSmemBytes, st>>>(h->d_phi.p, h->d_P.p, h->d_chunk_ids.p, cmask, h->d_rho_part.p,
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
    h->fock_valid = -1;
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
    k_spline<<<(unsigned)((nsys + 63) / 64), 64, sms ? sm_spline : 0, st>>>(g, h->spline, h->d_U_lm.p, h->binned ? h->d_pre_scaled.p : h->d_pre.p,
                                                                            h->d_work.p, h->d_coef.p, sms);
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
#define DFG_BIN_ARGS g, h->d_atom_xyz.p, h->d_x.p, h->d_y.p, h->d_z.p, h->d_xs.p, h->d_coef.p, h->d_item_key.p, h->d_pair_point.p, h->bin_nitems, h->d_pair_out.p, h->d_pair_geo.p, (long)h->bin_nitems * 32 * h->bin_R
#define DFG_BIN_LAUNCH(LL, RR, MB)                                                                              \
    if (geo)                                                                                                    \
        k_interp_bin<LL, RR, MB, true><<<bx, kBinWarps * 32, smem, st>>>(DFG_BIN_ARGS);                         \
    else                                                                                                        \
        k_interp_bin<LL, RR, MB, false><<<bx, kBinWarps * 32, smem, st>>>(DFG_BIN_ARGS)
#define DFG_BIN_CASE(LL)                                                                                        \
    case LL:                                                                                                    \
        if (h->bin_R == 4) {                                                                                    \
            DFG_BIN_LAUNCH(LL, 4, 3);                                                                           \
        } else if (h->bin_R == 3) {                                                                             \
            DFG_BIN_LAUNCH(LL, 3, 4);                                                                           \
        } else if (five) {                                                                                      \
            DFG_BIN_LAUNCH(LL, 2, 5);                                                                           \
        } else {                                                                                                \
            DFG_BIN_LAUNCH(LL, 2, 6);                                                                           \
        }                                                                                                       \
        break;
            const bool geo = h->d_pair_geo.p != nullptr;
            static const bool five = dev_env("DFTGRID_INTERP_MINB5") != nullptr;  // developer A/B switch
            switch (g.lmax) {
                DFG_BIN_CASE(5)
                DFG_BIN_CASE(8)
                DFG_BIN_CASE(10)
                DFG_BIN_CASE(11)
                default: throw std::runtime_error("binned interpolation: unsupported lmax");
            }
#undef DFG_BIN_CASE
#undef DFG_BIN_LAUNCH
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

// Contraction modes: kModePair  [XC | J] = Phi^T diag(d) Phi as two matrices, res = [J (nb^2) | XC (nb^2) | exc | nel];
//                    kModeFock  F_grid = 2J + XC as ONE contraction with the summed weight vector (half the DMMA work),
//                    kModeFockJ F_grid = 2J only (the reference's first iteration: XC still zero, src/dft.cpp:219-226);
//                    the two Fock modes write fres = [F (nb^2) | per-shell sums of w V rho | e_j | exc | nel].
enum { kModePair = 0, kModeFock = 1, kModeFockJ = 2 };

void run_contract(dftgrid* h, int mode) {
    cudaStream_t st = h->stream;
    const GridShape& g = h->g;
    const size_t nb2 = (size_t)h->nbf * h->nbf;
    const long nshell = (long)g.natoms * g.nrad;
    const bool fock = mode != kModePair;
    const dftgrid::DevSchedule& D = h->sched[fock ? 1 : 0];
    const int nz = fock ? 1 : 2;
    record(h, 12);
    double* ejshell = h->d_fres.p + nb2;
    if (fock) {
        // summed weight vector and the per-shell sums of the E_J integrand (zero outside this rank's shells)
        if (h->nranks > 1) CK(cudaMemsetAsync(ejshell, 0, (size_t)nshell * sizeof(double), st));
        if (g.nloc > 0) {
            k_fock_weights<<<(unsigned)((g.nloc + 255) / 256), 256, 0, st>>>(g.nloc, h->d_dJ.p, h->d_dxc.p, mode == kModeFock ? 1 : 0, h->d_dF.p);
            k_shell_sum<<<(unsigned)((g.nshell_loc * 32 + 255) / 256), 256, 0, st>>>(g, h->d_dJ.p, h->d_rho.p, ejshell, 1, 0);
            h->launches += 2;
        }
    }
    if (h->peer_ready) {
        k_peer_begin<<<1, 1, 0, st>>>(h->peers);
        h->peer_epoch++;
        h->peer_used = true;
        h->launches++;
    }
    ConSync csync{nullptr, 1, 0, 0};
    if (h->con_sync_wc > 0 && D.n_single >= 16) {
        h->d_con_sync.zero(st);
        csync = ConSync{h->d_con_sync.p, h->con_sync_wc, h->con_sync_lead, std::max(1, D.n_single * h->con_sync_pct / 100)};
    }
    k_contract_tma<<<D.ctas, kConTmaThreads, kConTmaSmemBytes, st>>>(h->d_phi.p, fock ? h->d_dF.p : h->d_dxc.p, h->d_dJ.p,
                                                                     h->screened ? h->d_con_chunk_ids.p : h->d_chunk_ids.p,
                                                                     h->screened ? h->d_con_chunk_mask.p : nullptr, h->d_pairs.p, D.segs.p, D.cta_off.p, h->d_partial.p, h->nbp, (int)h->n_active_chunks, h->con_bc, h->d_dbg_times.p, csync);
    if (h->d_dbg_times.p && !h->capturing) {
        // developer instrumentation: per-CTA wall time of the contraction with the CTA's segments
        std::vector<unsigned long long> t(3 * (size_t)D.ctas);
        CK(cudaMemcpyAsync(t.data(), h->d_dbg_times.p, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        unsigned long long t0 = ~0ull;
        for (int c = 0; c < D.ctas; c++) t0 = std::min(t0, t[3 * c]);
        std::fprintf(stderr, "[dftgrid] contraction CTA times (start us, duration us, stages) mode %d\n", mode);
        for (int c = 0; c < D.ctas; c++)
            std::fprintf(stderr, "  cta %3d  %8.1f %8.1f %6llu sm %llu\n", c, (t[3 * c] - t0) * 1e-3, (t[3 * c + 1] - t[3 * c]) * 1e-3,
                         t[3 * c + 2] & 0xffffffffull, t[3 * c + 2] >> 32);
    }
    double* res = fock ? h->d_fres.p : h->d_res.p;
    if (h->peer_ready) {
        // split-K reduction straight into this rank's exchange buffer, then the cross-rank sum over peer memory
        if (fock)
            k_contract_reduce_publish<<<dim3(h->npairs, 1, kReduceSplit), 256, 0, st>>>(h->d_partial.p, h->d_pairs.p, D.item_off.p, h->npairs, h->nbf, h->nbp,
                                                                                     1.0, 1.0, 0, 0, ejshell, (int)nshell, nb2, h->peers);
        else  // res layout [J | XC]; item z = 0 is XC
            k_contract_reduce_publish<<<dim3(h->npairs, 2, kReduceSplit), 256, 0, st>>>(h->d_partial.p, h->d_pairs.p, D.item_off.p, h->npairs, h->nbf, h->nbp,
                                                                                     1.0, 0.5, nb2, 0, nullptr, 0, 2 * nb2, h->peers);
    } else {
        if (fock)
            k_contract_reduce<<<dim3(h->npairs, 1, kReduceSplit), 256, 0, st>>>(h->d_partial.p, h->d_pairs.p, D.item_off.p, h->npairs, h->nbf, h->nbp, 1.0, 1.0,
                                                                             res, res);
        else
            k_contract_reduce<<<dim3(h->npairs, 2, kReduceSplit), 256, 0, st>>>(h->d_partial.p, h->d_pairs.p, D.item_off.p, h->npairs, h->nbf, h->nbp, 1.0, 0.5,
                                                                             res + nb2, res);
    }
    h->launches += 2;
    record(h, 13);
    if (h->peer_ready) {
        const size_t n = nz * nb2 + (fock ? (size_t)nshell : 0);
        const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, 296);
        k_peer_sum<<<blocks, 256, 0, st>>>(h->peers, h->nbf, nz, fock ? (int)nshell : 0, res);
        h->launches++;
    } else {
        allreduce(h, res, nz * nb2 + (fock ? (size_t)nshell : 0));
    }
    // exc and nel come from the already-reduced shell sums (identical on every rank): appended after the reduced block
    if (fock) {
        k_fock_scalars<<<1, 256, 0, st>>>(g, ejshell, h->d_scalars.p, h->d_qatom2.p, h->d_fres.p + nb2 + nshell);
        h->launches++;
    } else {
        CK(cudaMemcpyAsync(res + 2 * nb2, h->d_scalars.p + 2, sizeof(double), cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(res + 2 * nb2 + 1, h->d_scalars.p + 1, sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    record(h, 14);
    if (fock) {
        h->fock_valid = h->have_potential ? mode : -1;
        h->contract_valid = false;
    } else {
        h->contract_valid = h->have_potential;
        h->fock_valid = -1;
    }
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
// The flag is sticky on purpose: the epochs of the ranks are out of step from then on, the handle must be recreated.
void check_peer_error(dftgrid* h) {
    if (!h->peer_ready) return;
    unsigned long long err = 0;
    CK(cudaMemcpy(&err, h->xbuf + offsetof(PeerHeader, error), sizeof err, cudaMemcpyDeviceToHost));
    if (err) throw std::runtime_error("peer-memory reduction timed out waiting for another rank (the handle is unusable now: destroy and recreate it)");
}

template <typename T>
void download(dftgrid* h, const DevBuf<T>& b, T* out, size_t count) {
    use_device(h);
    CK(cudaMemcpyAsync(out, b.p, count * sizeof(T), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
}

// Rows [r0, r1) of an nb x nb device matrix into the caller's matrix: straight from the DMA engine when the caller's
// buffer is page-locked, through the handle's pinned staging otherwise (`pending` remembers the memcpy still to do after
// the stream has been synchronised).
struct PendingCopy {
    double* dst;
    const double* src;
    size_t bytes;
};
void download_rows(dftgrid* h, const double* dev, double* host, double* stage, int r0, int r1, std::vector<PendingCopy>& pending) {
    if (!host || r1 <= r0) return;
    const size_t off = (size_t)r0 * h->nbf, bytes = (size_t)(r1 - r0) * h->nbf * sizeof(double);
    if (is_pinned_host(host)) {
        CK(cudaMemcpyAsync(host + off, dev + off, bytes, cudaMemcpyDeviceToHost, h->stream));
    } else {
        CK(cudaMemcpyAsync(stage + off, dev + off, bytes, cudaMemcpyDeviceToHost, h->stream));
        pending.push_back(PendingCopy{host + off, stage + off, bytes});
    }
}

void require_density(dftgrid* h) {
    if (!h->have_density) throw std::runtime_error("dftgrid_set_density has not been called");
}

// One whole iteration's kernels in the given contraction mode, eagerly the first time (every lazily sized buffer exists
// afterwards), captured into a CUDA graph the second time and replayed from then on.  Multi-rank handles are captured
// too: the peer-memory kernels read their epoch from device memory and NCCL collectives are capturable.
void run_iteration_device(dftgrid* h, int mode) {
    if (!h->built) throw std::runtime_error("dftgrid_build has not been called");
    static const bool no_graph = dev_env("DFTGRID_NO_GRAPH") != nullptr;  // developer A/B switch
    dftgrid::IterGraph& G = h->graphs[mode];
    const bool want_graph = !no_graph && !G.failed && mode != kModeFockJ;
    auto mark = [&] {
        h->have_density = h->have_potential = h->timed_iter = true;
        h->contract_valid = mode == kModePair;
        h->fock_valid = mode == kModePair ? -1 : mode;
    };
    if (want_graph && G.exec) {
        CK(cudaGraphLaunch(G.exec, h->stream));
        h->launches += G.launches_per_iter;
        if (h->peer_ready) {
            h->peer_epoch++;
            h->peer_used = true;
        }
        mark();
        return;
    }
    if (want_graph && G.eager >= 1) {
        const long l0 = h->launches;
        const unsigned long long e0 = h->peer_epoch;
        bool ok = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
            h->capturing = true;
            try {
                run_density(h);
                run_potential(h);
                run_contract(h, mode);
            } catch (...) {
                ok = false;
            }
            h->capturing = false;
            cudaGraph_t gcap = nullptr;
            if (cudaStreamEndCapture(h->stream, &gcap) != cudaSuccess || !gcap) ok = false;
            if (ok && cudaGraphInstantiate(&G.exec, gcap, 0) != cudaSuccess) ok = false;
            if (ok) {
                G.graph = gcap;
                G.launches_per_iter = h->launches - l0;
            } else if (gcap) {
                cudaGraphDestroy(gcap);
            }
        }
        h->launches = l0;
        h->peer_epoch = e0;  // nothing ran during the capture
        if (!ok) {
            cudaGetLastError();
            G.exec = nullptr;
            G.failed = true;  // stay on eager launches
        } else {
            CK(cudaGraphLaunch(G.exec, h->stream));
            h->launches += G.launches_per_iter;
            if (h->peer_ready) {
                h->peer_epoch++;
                h->peer_used = true;
            }
            mark();
            return;
        }
    }
    run_density(h);
    run_potential(h);
    run_contract(h, mode);
    G.eager++;
    CK(cudaGetLastError());
}


// ---- device-resident SCF step (kernels_scf.cuh) ----------------------------------------------------------------------
template <bool SYM>
void scf_gemm(dftgrid* h, const double* A, const double* B, double* C, const int* skip) {
    const int np = h->scf.np, nt = np / kGemmTile;
    if (SYM)
        k_gemm_sym32<<<nt * (nt + 1), kGemmThreads, kGemmSmemBytes2, h->stream>>>(A, B, C, np, skip);  // sum over ti of 2 (nt - ti) half-width tiles
    else
        k_gemm_nn<false><<<dim3(nt, nt), kGemmThreads, kGemmSmemBytes, h->stream>>>(A, B, C, np, skip);
    h->launches++;
}

void scf_init(dftgrid* h, const double* H, const double* X, int nocc, double alpha) {
    if (!h->built) throw std::runtime_error("dftgrid_build has not been called");
    if (nocc < 0 || nocc > h->nbf) throw std::runtime_error("dftgrid_scf_init: nocc out of range");
    auto& S = h->scf;
    cudaStream_t st = h->stream;
    const int nb = h->nbf, np = round_up(nb, kGemmTile);
    const size_t nb2 = (size_t)nb * nb, np2 = (size_t)np * np;
    S.np = np;
    S.nocc = nocc;
    S.alpha = alpha;
    S.steps = 0;
    S.H.alloc(nb2);
    for (DevBuf<double>* b : {&S.X, &S.Xt, &S.F, &S.T1, &S.Fp, &S.D, &S.D2, &S.D3}) b->alloc(np2);
    for (DevBuf<double>* b : {&S.lo, &S.hi, &S.diag, &S.rows}) b->alloc(nb);
    S.eone.alloc(1);
    S.pm.alloc(1);
    if (!S.h_pm) CK(cudaMallocHost(&S.h_pm, sizeof(PmState)));
    if (!S.h_out) CK(cudaMallocHost(&S.h_out, 8 * sizeof(double)));
    for (auto& e : S.ev)
        if (!e) CK(cudaEventCreate(&e));
    CK(cudaFuncSetAttribute(k_gemm_nn<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmemBytes));
    CK(cudaFuncSetAttribute(k_gemm_sym32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmemBytes2));
    CK(cudaFuncSetAttribute(k_gemm_nn<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmemBytes));
    // H and X arrive row-major nb x nb (H symmetric; X = U s^-1/2 is not): pad X, build X^T once
    CK(cudaMemcpyAsync(S.H.p, H, nb2 * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(S.T1.p, X, nb2 * sizeof(double), cudaMemcpyHostToDevice, st));
    const unsigned eb = (unsigned)((np2 + 255) / 256);
    k_scf_pad_sum<<<eb, 256, 0, st>>>(S.T1.p, nullptr, nb, np, S.X.p);
    k_scf_transpose<<<dim3(np / 32, np / 32), dim3(32, 8), 0, st>>>(S.X.p, np, S.Xt.p);
    h->launches += 2;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    S.ready = true;
}

// out[8] = e_one, e_j, exc, nelec, purification steps, idempotency residual |tr(D - D^2)|, algebra ms, grid ms
void scf_step(dftgrid* h, int include_xc, double* out) {
    auto& S = h->scf;
    if (!S.ready) throw std::runtime_error("dftgrid_scf_init has not been called");
    cudaStream_t st = h->stream;
    const int nb = h->nbf, np = S.np;
    const size_t nb2 = (size_t)nb * nb, np2 = (size_t)np * np;
    const unsigned eb = (unsigned)((np2 + 255) / 256);
    CK(cudaEventRecord(S.ev[0], st));
    // F = H + F_grid of the previous step (none before the first density: core-Hamiltonian guess, src/dft.cpp:176-182)
    k_scf_pad_sum<<<eb, 256, 0, st>>>(S.H.p, S.steps > 0 ? h->d_fres.p : nullptr, nb, np, S.F.p);
    h->launches++;
    scf_gemm<false>(h, S.F.p, S.X.p, S.T1.p, nullptr);   // T1 = F X
    scf_gemm<true>(h, S.Xt.p, S.T1.p, S.Fp.p, nullptr);  // F' = X^T F X
    int pm_iters = 0;
    double pm_err = 0.0;
    if (S.nocc >= nb || S.nocc == 0) {
        if (S.nocc == 0)
            CK(cudaMemsetAsync(S.D.p, 0, np2 * sizeof(double), st));
        else
            k_scf_identity<<<eb, 256, 0, st>>>(nb, np, S.D.p);
        h->launches++;
    } else {
        k_pm_gershgorin<<<(unsigned)((nb * 32 + 255) / 256), 256, 0, st>>>(S.Fp.p, nb, np, S.lo.p, S.hi.p, S.diag.p);
        k_pm_setup<<<1, 256, 0, st>>>(S.lo.p, S.hi.p, S.diag.p, nb, S.nocc, S.pm.p);
        k_pm_init<<<eb, 256, 0, st>>>(S.Fp.p, nb, np, S.pm.p, S.D.p);
        h->launches += 3;
        const int* skip = &S.pm.p->done;
        const int kBatch = 8, kMaxIter = 128;
        S.h_pm->done = 0;
        for (int it = 0; it < kMaxIter && !S.h_pm->done; it += kBatch) {
            for (int b = 0; b < kBatch; b++) {
                scf_gemm<true>(h, S.D.p, S.D.p, S.D2.p, skip);
                scf_gemm<true>(h, S.D.p, S.D2.p, S.D3.p, skip);
                k_pm_coeff<<<1, 256, 0, st>>>(S.D.p, S.D2.p, S.D3.p, nb, np, S.pm.p);
                k_pm_update<<<eb, 256, 0, st>>>(S.D.p, S.D2.p, S.D3.p, np, S.pm.p);
                h->launches += 2;
            }
            CK(cudaMemcpyAsync(S.h_pm, S.pm.p, sizeof(PmState), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
        }
        if (S.h_pm->failed) throw std::runtime_error("density purification: F' is a multiple of the identity");
        if (!S.h_pm->done)
            throw std::runtime_error("density purification did not converge (no gap between the occupied and virtual orbitals?); "
                                     "use the host eigen-solver (scf = host)");
        pm_iters = S.h_pm->iters;
        pm_err = S.h_pm->err;
    }
    scf_gemm<false>(h, S.X.p, S.D.p, S.T1.p, nullptr);   // T2 = X D'
    scf_gemm<true>(h, S.T1.p, S.Xt.p, S.D2.p, nullptr);  // Pnew = X D' X^T
    k_scf_mix<<<(unsigned)((nb2 + 255) / 256), 256, 0, st>>>(S.D2.p, nb, np, S.alpha, S.steps == 0 ? 1 : 0, h->d_Praw.p);
    k_pad_P<<<(unsigned)((nb2 + 255) / 256), 256, 0, st>>>(h->d_Praw.p, h->d_P.p, nb, h->nbp);
    k_scf_rowdot<<<(unsigned)((nb * 32 + 255) / 256), 256, 0, st>>>(h->d_Praw.p, S.H.p, nb, S.rows.p);
    k_scf_trace_finish<<<1, 256, 0, st>>>(S.rows.p, nb, 2.0, S.eone.p);
    h->launches += 4;
    CK(cudaEventRecord(S.ev[1], st));
    run_iteration_device(h, include_xc ? kModeFock : kModeFockJ);
    CK(cudaEventRecord(S.ev[2], st));
    const size_t nshell = (size_t)h->g.natoms * h->g.nrad;
    CK(cudaMemcpyAsync(S.h_out, S.eone.p, sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(S.h_out + 1, h->d_fres.p + nb2 + nshell, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    check_peer_error(h);
    S.steps++;
    float ms_alg = 0.f, ms_grid = 0.f;
    CK(cudaEventElapsedTime(&ms_alg, S.ev[0], S.ev[1]));
    CK(cudaEventElapsedTime(&ms_grid, S.ev[1], S.ev[2]));
    if (out) {
        for (int i = 0; i < 4; i++) out[i] = S.h_out[i];
        out[4] = pm_iters;
        out[5] = pm_err;
        out[6] = ms_alg;
        out[7] = ms_grid;
    }
}

// ---- single-process multi-GPU: a group handle owns one rank handle per device and one worker thread per rank; every
// public entry point forwards to the rank handles on their threads and returns when all of them are done.
struct Group {
    std::vector<dftgrid*> subs;
    std::vector<std::thread> threads;
    std::mutex m;
    std::condition_variable cv_job, cv_done;
    std::function<int(dftgrid*, int)> job;
    unsigned long gen = 0;
    int pending = 0;
    bool stop = false;
    std::vector<int> rc;
    std::vector<std::string> err;

    void worker(int r) {
        cudaSetDevice(subs[r]->device);
        unsigned long seen = 0;
        for (;;) {
            std::function<int(dftgrid*, int)> f;
            {
                std::unique_lock<std::mutex> lk(m);
                cv_job.wait(lk, [&] { return stop || gen != seen; });
                if (stop) return;
                seen = gen;
                f = job;
            }
            int code = 0;
            try {
                code = f(subs[r], r);
            } catch (const std::exception& e) {
                g_error = e.what();
                code = 1;
            } catch (...) {
                g_error = "unknown error";
                code = 2;
            }
            {
                std::lock_guard<std::mutex> lk(m);
                rc[r] = code;
                if (code) err[r] = g_error;
                if (--pending == 0) cv_done.notify_all();
            }
        }
    }
    void start() {
        rc.assign(subs.size(), 0);
        err.assign(subs.size(), std::string());
        for (size_t r = 0; r < subs.size(); r++) threads.emplace_back([this, r] { worker((int)r); });
    }
    int run(std::function<int(dftgrid*, int)> f) {
        std::unique_lock<std::mutex> lk(m);
        job = std::move(f);
        pending = (int)subs.size();
        gen++;
        cv_job.notify_all();
        cv_done.wait(lk, [&] { return pending == 0; });
        for (size_t r = 0; r < subs.size(); r++)
            if (rc[r]) {
                g_error = "rank " + std::to_string(r) + ": " + err[r];
                return rc[r];
            }
        return 0;
    }
    void shutdown() {
        {
            std::lock_guard<std::mutex> lk(m);
            stop = true;
        }
        cv_job.notify_all();
        for (auto& t : threads) t.join();
        threads.clear();
    }
};

inline void row_slice(int nb, int r, int n, int* r0, int* r1) {
    *r0 = (int)((long)nb * r / n);
    *r1 = (int)((long)nb * (r + 1) / n);
}

void alloc_xbuf(dftgrid* h) {
    if (h->xbuf) return;
    const size_t nb2 = (size_t)h->nbf * h->nbf;
    const size_t nshell = (size_t)h->natoms * h->prm.radial_points;
    h->xslot = std::max(2 * nb2, nb2 + nshell);
    h->xbuf_bytes = kPeerHeaderBytes + 2 * h->xslot * sizeof(double);
    CK(cudaMalloc(&h->xbuf, h->xbuf_bytes));
    CK(cudaMemset(h->xbuf, 0, h->xbuf_bytes));
    const unsigned long long lim = kPeerDefaultTimeoutNs;
    CK(cudaMemcpy(h->xbuf + offsetof(PeerHeader, spin_limit_ns), &lim, sizeof lim, cudaMemcpyHostToDevice));
}

}  // namespace

struct dftgrid_group : Group {};

extern "C" {

const char* dftgrid_last_error(void) { return g_error.c_str(); }
int dftgrid_abi_version(void) { return 2; }

#define GROUP_FORWARD(h, expr)                                                        \
    if ((h)->group) return (h)->group->run([&](dftgrid* s, int r) -> int { (void)r; return (expr); })

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

int dftgrid_create_multi(dftgrid_t** out, const dftgrid_system* sys, const dftgrid_params* prm, int ngpus, const int* devices) {
    if (!out) {
        g_error = "null argument";
        return 1;
    }
    *out = nullptr;
    if (ngpus == 1) return dftgrid_create(out, sys, prm, devices ? devices[0] : 0, 0, 1);
    std::unique_ptr<dftgrid> gh(new dftgrid());
    int rc = guarded([&] {
        if (ngpus < 1 || ngpus > kPeerMaxRanks) throw std::runtime_error("ngpus must be 1..16");
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            throw std::runtime_error(std::string("no CUDA device available (this library has no CPU fallback): ") + cudaGetErrorString(e));
        if (ngpus > ndev && !devices) throw std::runtime_error("ngpus exceeds the number of visible CUDA devices (" + std::to_string(ndev) + ")");
        gh->group = new dftgrid_group();
        gh->nranks = ngpus;
        for (int r = 0; r < ngpus; r++) {
            dftgrid_t* s = nullptr;
            if (dftgrid_create(&s, sys, prm, devices ? devices[r] : r, r, ngpus) != 0) throw std::runtime_error(g_error);
            gh->group->subs.push_back(s);
        }
        gh->nbf = gh->group->subs[0]->nbf;
        gh->natoms = gh->group->subs[0]->natoms;
        gh->prm = *prm;
        gh->device = gh->group->subs[0]->device;
        // the small per-shell sums go through NCCL (one communicator per device, created in one call)
        if (!nccl_api().load() || !nccl_api().CommInitAll) throw std::runtime_error("cannot load libnccl.so.2 (needed for ngpus > 1)");
        std::vector<NcclComm> comms(ngpus);
        std::vector<int> devs(ngpus);
        for (int r = 0; r < ngpus; r++) devs[r] = gh->group->subs[r]->device;
        int nrc = nccl_api().CommInitAll(comms.data(), ngpus, devs.data());
        if (nrc != 0) throw std::runtime_error(std::string("ncclCommInitAll: ") + nccl_api().GetErrorString(nrc));
        for (int r = 0; r < ngpus; r++) gh->group->subs[r]->comm = comms[r];
        gh->group->start();
        // peer-memory path for the [J | XC] / F sum when every pair of devices has a P2P route
        bool p2p = true;
        for (int a = 0; a < ngpus && p2p; a++)
            for (int b = 0; b < ngpus && p2p; b++) {
                int can = 1;
                if (a != b && (cudaDeviceCanAccessPeer(&can, devs[a], devs[b]) != cudaSuccess || !can)) p2p = false;
            }
        static const bool no_peer = dev_env("DFTGRID_NO_PEER") != nullptr;  // developer A/B switch
        if (p2p && !no_peer) {
            int prc = gh->group->run([&](dftgrid* s, int r) -> int {
                return guarded([&] {
                    use_device(s);
                    for (int b = 0; b < ngpus; b++)
                        if (b != r) {
                            cudaError_t pe = cudaDeviceEnablePeerAccess(devs[b], 0);
                            if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) CK(pe);
                            cudaGetLastError();
                        }
                    alloc_xbuf(s);
                });
            });
            if (prc != 0) throw std::runtime_error(g_error);
            PeerSet ps{};
            ps.nranks = ngpus;
            ps.slot = gh->group->subs[0]->xslot;
            for (int r = 0; r < ngpus; r++) ps.base[r] = gh->group->subs[r]->xbuf;
            for (int r = 0; r < ngpus; r++) {
                dftgrid* s = gh->group->subs[r];
                s->peers = ps;
                s->peers.rank = r;
                s->peer_ready = true;
                s->peer_local = true;
            }
        }
    });
    if (rc != 0) {
        const std::string keep = g_error;
        dftgrid_destroy(gh.release());
        g_error = keep;
        return rc;
    }
    *out = gh.release();
    return 0;
}

int dftgrid_ngpus(const dftgrid_t* h) { return h->group ? (int)h->group->subs.size() : 1; }

void dftgrid_destroy(dftgrid_t* h) {
    if (!h) return;
    if (h->group) {
        // every rank's stream is drained before any exchange buffer is freed: nobody is still reading a peer
        if (!h->group->threads.empty()) {
            h->group->run([](dftgrid* s, int) -> int {
                cudaSetDevice(s->device);
                cudaStreamSynchronize(s->stream);
                return 0;
            });
            h->group->shutdown();
        }
        for (dftgrid* s : h->group->subs) dftgrid_destroy(s);
        h->group->subs.clear();
        delete h->group;
        h->group = nullptr;
        delete h;
        return;
    }
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->peer_used && !h->peer_local && h->peer_epoch > 0) {
        // one process per GPU: another rank may still be summing this rank's exchange buffer (its k_peer_sum of the last
        // epoch).  Wait, bounded, until every peer has written consumed_by[peer] >= the last epoch into THIS rank's header;
        // if a peer never gets there the buffer is leaked rather than freed under its feet (freeing an exported
        // allocation another process still has mapped is undefined behaviour).
        const auto t0 = std::chrono::steady_clock::now();
        bool all = true;
        for (int r = 0; r < h->peers.nranks; r++) {
            if (r == h->rank) continue;
            unsigned long long consumed = 0;
            do {
                if (cudaMemcpy(&consumed, h->xbuf + offsetof(PeerHeader, consumed_by) + r * sizeof(unsigned long long), sizeof consumed,
                               cudaMemcpyDeviceToHost) != cudaSuccess) {
                    cudaGetLastError();
                    break;
                }
            } while (consumed < h->peer_epoch && std::chrono::steady_clock::now() - t0 < std::chrono::seconds(30));
            if (consumed < h->peer_epoch) all = false;
        }
        if (!all) h->xbuf = nullptr;  // leaked on purpose
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
        if (h->group) throw std::runtime_error("a multi-GPU handle wires its own communicator");
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
        if (h->group) throw std::runtime_error("a multi-GPU handle wires its own peer mappings");
        use_device(h);
        if (!handle64) throw std::runtime_error("null argument");
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
        if (h->nranks < 2 || h->nranks > kPeerMaxRanks) throw std::runtime_error("peer reduction needs 2..16 ranks");
        alloc_xbuf(h);
        cudaIpcMemHandle_t hd;
        CK(cudaIpcGetMemHandle(&hd, h->xbuf));
        std::memcpy(handle64, &hd, sizeof hd);
    });
}

int dftgrid_peer_connect(dftgrid_t* h, const void* handles) {
    return guarded([&] {
        if (h->group) throw std::runtime_error("a multi-GPU handle wires its own peer mappings");
        use_device(h);
        if (!handles) throw std::runtime_error("null argument");
        if (!h->xbuf) throw std::runtime_error("dftgrid_peer_export has not been called");
        if (h->peer_ready) return;
        PeerSet ps{};
        ps.nranks = h->nranks;
        ps.rank = h->rank;
        ps.slot = h->xslot;
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

int dftgrid_peer_active(const dftgrid_t* h) {
    if (h->group) return h->group->subs[0]->peer_ready ? 1 : 0;
    return h->peer_ready ? 1 : 0;
}

int dftgrid_peer_disable(dftgrid_t* h) {
    GROUP_FORWARD(h, dftgrid_peer_disable(s));
    return guarded([&] {
        use_device(h);
        CK(cudaStreamSynchronize(h->stream));
        h->peer_ready = false;  // back to ncclAllReduce; the mappings stay open until the handle is destroyed
        for (auto& G : h->graphs) G.reset();  // captured graphs hold the peer kernels
    });
}

int dftgrid_peer_set_timeout(dftgrid_t* h, double seconds) {
    GROUP_FORWARD(h, dftgrid_peer_set_timeout(s, seconds));
    return guarded([&] {
        use_device(h);
        if (!h->xbuf) throw std::runtime_error("no peer exchange buffer on this handle");
        if (!(seconds > 0.0)) throw std::runtime_error("timeout must be positive");
        const unsigned long long lim = (unsigned long long)(seconds * 1e9);
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaMemcpy(h->xbuf + offsetof(PeerHeader, spin_limit_ns), &lim, sizeof lim, cudaMemcpyHostToDevice));
    });
}

int dftgrid_build(dftgrid_t* h) {
    if (h->group) {
        int rc = h->group->run([&](dftgrid* s, int) -> int { return dftgrid_build(s); });
        if (rc == 0) {
            h->g = h->group->subs[0]->g;
            h->g.shell0 = 0;
            h->g.nshell_loc = (long)h->g.natoms * h->g.nrad;
            h->g.nloc = h->g.npts;
            h->built = true;
        }
        return rc;
    }
    return guarded([&] {
        use_device(h);
        if (h->built) throw std::runtime_error("dftgrid_build has already been called on this handle");
        do_build(h);
    });
}

long dftgrid_npoints(const dftgrid_t* h) { return h->g.npts; }
long dftgrid_npoints_local(const dftgrid_t* h) { return h->g.nloc; }
long dftgrid_point_offset(const dftgrid_t* h) { return h->g.shell0 * h->g.nang; }
int dftgrid_nbf(const dftgrid_t* h) { return h->nbf; }
int dftgrid_nlm(const dftgrid_t* h) { return h->g.nlm; }

int dftgrid_upload_density(dftgrid_t* h, const double* P) {
    GROUP_FORWARD(h, dftgrid_upload_density(s, P));
    return guarded([&] {
        use_device(h);
        if (upload_P(h, P)) CK(cudaStreamSynchronize(h->stream));  // the caller may reuse P as soon as this returns
    });
}

int dftgrid_set_density(dftgrid_t* h, const double* P) {
    GROUP_FORWARD(h, dftgrid_set_density(s, P));
    return guarded([&] {
        use_device(h);
        upload_P(h, P);
        run_density(h);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(h->stream));
    });
}

// rows [r0, r1) of J only (every rank of a group holds the whole matrix; each one delivers its slice)
static int hartree_J_rows(dftgrid_t* h, double* J, int r0, int r1) {
    return guarded([&] {
        use_device(h);
        require_density(h);
        run_potential(h);
        run_contract(h, kModePair);
        CK(cudaGetLastError());
        std::vector<PendingCopy> pend;
        download_rows(h, h->d_res.p, J, h->h_res, r0, r1, pend);
        CK(cudaStreamSynchronize(h->stream));
        check_peer_error(h);
        for (auto& c : pend) std::memcpy(c.dst, c.src, c.bytes);
    });
}

int dftgrid_hartree_J(dftgrid_t* h, double* J) {
    if (h->group) {
        const int n = (int)h->group->subs.size();
        return h->group->run([&](dftgrid* s, int r) -> int {
            int r0, r1;
            row_slice(s->nbf, r, n, &r0, &r1);
            return hartree_J_rows(s, J, r0, r1);
        });
    }
    return hartree_J_rows(h, J, 0, h->nbf);
}

static int xc_rows(dftgrid_t* h, double* XC, double* exc, int r0, int r1) {
    return guarded([&] {
        use_device(h);
        require_density(h);
        const size_t nb2 = (size_t)h->nbf * h->nbf;
        if (!h->contract_valid) {
            // XC asked for before J: the J half of the two-matrix contraction still needs a defined weight vector
            if (!h->have_potential) h->d_dJ.zero(h->stream);
            run_contract(h, kModePair);
        }
        CK(cudaGetLastError());
        std::vector<PendingCopy> pend;
        download_rows(h, h->d_res.p + nb2, XC, h->h_res + nb2, r0, r1, pend);
        CK(cudaMemcpyAsync(h->h_res + 2 * nb2, h->d_res.p + 2 * nb2, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        check_peer_error(h);
        for (auto& c : pend) std::memcpy(c.dst, c.src, c.bytes);
        if (exc) *exc = h->h_res[2 * nb2];
    });
}

int dftgrid_xc(dftgrid_t* h, double* XC, double* exc) {
    if (h->group) {
        const int n = (int)h->group->subs.size();
        return h->group->run([&](dftgrid* s, int r) -> int {
            int r0, r1;
            row_slice(s->nbf, r, n, &r0, &r1);
            return xc_rows(s, XC, r == 0 ? exc : nullptr, r0, r1);
        });
    }
    return xc_rows(h, XC, exc, 0, h->nbf);
}

int dftgrid_electron_count(dftgrid_t* h, double* nelec) {
    if (h->group) return dftgrid_electron_count(h->group->subs[0], nelec);
    return guarded([&] {
        use_device(h);
        require_density(h);
        CK(cudaMemcpyAsync(h->h_res, h->d_scalars.p + 1, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        *nelec = h->h_res[0];
    });
}

int dftgrid_iteration_device(dftgrid_t* h) {
    GROUP_FORWARD(h, dftgrid_iteration_device(s));
    return guarded([&] {
        use_device(h);
        run_iteration_device(h, kModePair);
    });
}

int dftgrid_fock_device(dftgrid_t* h, int include_xc) {
    GROUP_FORWARD(h, dftgrid_fock_device(s, include_xc));
    return guarded([&] {
        use_device(h);
        run_iteration_device(h, include_xc ? kModeFock : kModeFockJ);
    });
}

static int download_results_rows(dftgrid_t* h, double* J, double* XC, double* exc, double* nelec, int r0, int r1) {
    return guarded([&] {
        use_device(h);
        if (!h->contract_valid) throw std::runtime_error("no [J | XC] result on the device for the current density");
        const size_t nb2 = (size_t)h->nbf * h->nbf;
        std::vector<PendingCopy> pend;
        download_rows(h, h->d_res.p, J, h->h_res, r0, r1, pend);
        download_rows(h, h->d_res.p + nb2, XC, h->h_res + nb2, r0, r1, pend);
        CK(cudaMemcpyAsync(h->h_res + 2 * nb2, h->d_res.p + 2 * nb2, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        check_peer_error(h);
        for (auto& c : pend) std::memcpy(c.dst, c.src, c.bytes);
        if (exc) *exc = h->h_res[2 * nb2];
        if (nelec) *nelec = h->h_res[2 * nb2 + 1];
    });
}

int dftgrid_download_results(dftgrid_t* h, double* J, double* XC, double* exc, double* nelec) {
    if (h->group) {
        const int n = (int)h->group->subs.size();
        return h->group->run([&](dftgrid* s, int r) -> int {
            int r0, r1;
            row_slice(s->nbf, r, n, &r0, &r1);
            return download_results_rows(s, J, XC, r == 0 ? exc : nullptr, r == 0 ? nelec : nullptr, r0, r1);
        });
    }
    return download_results_rows(h, J, XC, exc, nelec, 0, h->nbf);
}

static int download_fock_rows(dftgrid_t* h, double* F, double* e_j, double* exc, double* nelec, int r0, int r1) {
    return guarded([&] {
        use_device(h);
        if (h->fock_valid < 0) throw std::runtime_error("no fused Fock result on the device for the current density");
        const size_t nb2 = (size_t)h->nbf * h->nbf;
        const size_t nshell = (size_t)h->g.natoms * h->g.nrad;
        std::vector<PendingCopy> pend;
        download_rows(h, h->d_fres.p, F, h->h_res, r0, r1, pend);
        CK(cudaMemcpyAsync(h->h_res + 2 * nb2, h->d_fres.p + nb2 + nshell, 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        check_peer_error(h);
        for (auto& c : pend) std::memcpy(c.dst, c.src, c.bytes);
        if (e_j) *e_j = h->h_res[2 * nb2];
        if (exc) *exc = h->h_res[2 * nb2 + 1];
        if (nelec) *nelec = h->h_res[2 * nb2 + 2];
    });
}

int dftgrid_download_fock(dftgrid_t* h, double* F, double* e_j, double* exc, double* nelec) {
    if (h->group) {
        const int n = (int)h->group->subs.size();
        return h->group->run([&](dftgrid* s, int r) -> int {
            int r0, r1;
            row_slice(s->nbf, r, n, &r0, &r1);
            return download_fock_rows(s, F, r == 0 ? e_j : nullptr, r == 0 ? exc : nullptr, r == 0 ? nelec : nullptr, r0, r1);
        });
    }
    return download_fock_rows(h, F, e_j, exc, nelec, 0, h->nbf);
}

int dftgrid_iteration(dftgrid_t* h, const double* P, double* J, double* XC, double* exc, double* nelec) {
    if (h->group) {
        const int n = (int)h->group->subs.size();
        return h->group->run([&](dftgrid* s, int r) -> int {
            int r0, r1;
            row_slice(s->nbf, r, n, &r0, &r1);
            int rc = guarded([&] {
                use_device(s);
                upload_P(s, P);
                run_iteration_device(s, kModePair);
            });
            if (rc) return rc;
            return download_results_rows(s, J, XC, r == 0 ? exc : nullptr, r == 0 ? nelec : nullptr, r0, r1);
        });
    }
    // P is consumed by the time the download below has synchronised the stream, so a pinned P is read in place
    int rc = guarded([&] {
        use_device(h);
        upload_P(h, P);
        run_iteration_device(h, kModePair);
    });
    if (rc) return rc;
    return download_results_rows(h, J, XC, exc, nelec, 0, h->nbf);
}

int dftgrid_fock(dftgrid_t* h, const double* P, int include_xc, double* F, double* e_j, double* exc, double* nelec) {
    const int mode = include_xc ? kModeFock : kModeFockJ;
    if (h->group) {
        const int n = (int)h->group->subs.size();
        return h->group->run([&](dftgrid* s, int r) -> int {
            int r0, r1;
            row_slice(s->nbf, r, n, &r0, &r1);
            int rc = guarded([&] {
                use_device(s);
                upload_P(s, P);
                run_iteration_device(s, mode);
            });
            if (rc) return rc;
            return download_fock_rows(s, F, r == 0 ? e_j : nullptr, r == 0 ? exc : nullptr, r == 0 ? nelec : nullptr, r0, r1);
        });
    }
    int rc = guarded([&] {
        use_device(h);
        upload_P(h, P);
        run_iteration_device(h, mode);
    });
    if (rc) return rc;
    return download_fock_rows(h, F, e_j, exc, nelec, 0, h->nbf);
}

int dftgrid_synchronize(dftgrid_t* h) {
    GROUP_FORWARD(h, dftgrid_synchronize(s));
    return guarded([&] {
        use_device(h);
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaGetLastError());
        check_peer_error(h);
    });
}

int dftgrid_get_positions(dftgrid_t* h, double* xyz) {
    GROUP_FORWARD(h, dftgrid_get_positions(s, xyz + 3 * dftgrid_point_offset(s)));
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
int dftgrid_get_weights(dftgrid_t* h, double* w) {
    GROUP_FORWARD(h, dftgrid_get_weights(s, w + dftgrid_point_offset(s)));
    return guarded([&] { download(h, h->d_w, w, (size_t)h->g.nloc); });
}
int dftgrid_get_becke_weights(dftgrid_t* h, double* wb) {
    GROUP_FORWARD(h, dftgrid_get_becke_weights(s, wb + dftgrid_point_offset(s)));
    return guarded([&] { download(h, h->d_wb, wb, (size_t)h->g.nloc); });
}
int dftgrid_get_densities(dftgrid_t* h, double* rho) {
    GROUP_FORWARD(h, dftgrid_get_densities(s, rho + dftgrid_point_offset(s)));
    return guarded([&] { download(h, h->d_rho, rho, (size_t)h->g.nloc); });
}
int dftgrid_get_potential(dftgrid_t* h, double* V) {
    GROUP_FORWARD(h, dftgrid_get_potential(s, V + dftgrid_point_offset(s)));
    return guarded([&] { download(h, h->d_V, V, (size_t)h->g.nloc); });
}
int dftgrid_get_amplitudes(dftgrid_t* h, double* phi) {
    GROUP_FORWARD(h, dftgrid_get_amplitudes(s, phi + (size_t)dftgrid_point_offset(s) * s->nbf));
    return guarded([&] {
        use_device(h);
        const size_t n = (size_t)h->g.nloc;
        if (n == 0) return;
        CK(cudaMemcpy2DAsync(phi, (size_t)h->nbf * sizeof(double), h->d_phi.p, (size_t)h->nbp * sizeof(double),
                             (size_t)h->nbf * sizeof(double), n, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    });
}
// DFT::construct_matrices' integral loop (src/dft.cpp:185-198): S, T, V over the handle's basis and nuclei
int dftgrid_one_electron(dftgrid_t* h, double* S, double* T, double* V) {
    if (h->group) return dftgrid_one_electron(h->group->subs[0], S, T, V);
    return guarded([&] {
        use_device(h);
        if (!h->built) throw std::runtime_error("dftgrid_build must be called before dftgrid_one_electron");
        const size_t nb2 = (size_t)h->nbf * h->nbf;
        DevBuf<double> dZ, dS;
        std::vector<double> zq(h->Z.begin(), h->Z.end());
        dZ.upload(zq, h->stream);
        dS.alloc(3 * nb2);
        const long npair = (long)h->nbf * (h->nbf + 1) / 2;
        k_one_electron<<<(unsigned)((npair + kIntWarps - 1) / kIntWarps), kIntWarps * 32, 0, h->stream>>>(
            h->nbf, h->natoms, h->d_bf_center.p, h->d_bf_prim_off.p, h->d_center_xyz.p, h->d_prim_exp.p, h->d_exp_alpha.p, h->d_prim_coeff.p,
            h->d_prim_norm.p, h->d_prim_lmn.p, h->d_atom_xyz.p, dZ.p, dS.p, dS.p + nb2, dS.p + 2 * nb2);
        h->launches++;
        CK(cudaGetLastError());
        if (S) CK(cudaMemcpyAsync(S, dS.p, nb2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        if (T) CK(cudaMemcpyAsync(T, dS.p + nb2, nb2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        if (V) CK(cudaMemcpyAsync(V, dS.p + 2 * nb2, nb2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    });
}

// RectangularGrid::build_grid + set_density (src/rectangulargrid.cpp:34-80): density and density gradient on a dp^3 box
static void launch_rect(dftgrid* h, int PT, const double* dP, double size, int dp, long npts, double* dpos, double* drho, double* dgrad) {
    const size_t smem = (size_t)4 * PT * h->nbf * sizeof(double);
    const unsigned grid = (unsigned)((npts + PT - 1) / PT);
#define DFG_RECT_LAUNCH(N)                                                                                                          \
    CK(cudaFuncSetAttribute(k_rect_density<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                              \
    k_rect_density<N><<<grid, kRectThreads, smem, h->stream>>>(h->nbf, h->d_bf_center.p, h->d_bf_prim_off.p, h->d_center_xyz.p,        \
                                                               h->d_prim_exp.p, h->d_exp_alpha.p, h->d_prim_coeff.p, h->d_prim_norm.p, \
                                                               h->d_prim_lmn.p, dP, size, dp, npts, dpos, drho, dgrad)
    if (PT == 4) {
        DFG_RECT_LAUNCH(4);
    } else if (PT == 2) {
        DFG_RECT_LAUNCH(2);
    } else {
        DFG_RECT_LAUNCH(1);
    }
#undef DFG_RECT_LAUNCH
    h->launches++;
}

int dftgrid_rectangular_density(dftgrid_t* h, double size, int dp, const double* P, double* pos, double* rho, double* grad) {
    if (h->group) return dftgrid_rectangular_density(h->group->subs[0], size, dp, P, pos, rho, grad);
    return guarded([&] {
        use_device(h);
        if (!h->built) throw std::runtime_error("dftgrid_build must be called before dftgrid_rectangular_density");
        if (!P) throw std::runtime_error("null density matrix");
        if (dp < 2 || dp > 1024 || !(size > 0.0)) throw std::runtime_error("rectangular grid: need size > 0 and 2 <= dp <= 1024");
        const long npts = (long)dp * dp * dp;
        const size_t nb2 = (size_t)h->nbf * h->nbf;
        DevBuf<double> dP, dpos, drho, dgrad;
        dP.alloc(nb2);
        dpos.alloc(3 * (size_t)npts);
        drho.alloc((size_t)npts);
        dgrad.alloc(3 * (size_t)npts);
        CK(cudaMemcpyAsync(dP.p, P, nb2 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        const size_t per_point = (size_t)4 * h->nbf * sizeof(double), limit = 200u << 10;
        if (per_point <= limit)
            launch_rect(h, 4 * per_point <= limit ? 4 : (2 * per_point <= limit ? 2 : 1), dP.p, size, dp, npts, dpos.p, drho.p, dgrad.p);
        else
            throw std::runtime_error("rectangular grid: basis too large for the shared-memory staging (nbf > 6400)");
        CK(cudaGetLastError());
        if (pos) CK(cudaMemcpyAsync(pos, dpos.p, 3 * (size_t)npts * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        if (rho) CK(cudaMemcpyAsync(rho, drho.p, (size_t)npts * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        if (grad) CK(cudaMemcpyAsync(grad, dgrad.p, 3 * (size_t)npts * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    });
}

int dftgrid_get_rho_lm(dftgrid_t* h, double* out) {
    if (h->group) return dftgrid_get_rho_lm(h->group->subs[0], out);
    return guarded([&] {
        use_device(h);
        const size_t nshell = (size_t)h->g.natoms * h->g.nrad;
        CK(cudaMemcpyAsync(out, h->d_shell2.p + nshell * 2, nshell * h->g.nlm * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    });
}
int dftgrid_get_U_lm(dftgrid_t* h, double* out) {
    if (h->group) return dftgrid_get_U_lm(h->group->subs[0], out);
    return guarded([&] { download(h, h->d_U_lm, out, (size_t)h->g.natoms * h->g.nrad * h->g.nlm); });
}

int dftgrid_last_timings(dftgrid_t* h, double* out, int n) {
    if (h->group) {
        // per phase, the slowest rank
        const int cnt = std::min(n, (int)DFTGRID_T_COUNT);
        std::vector<std::vector<double>> t(h->group->subs.size(), std::vector<double>(DFTGRID_T_COUNT, 0.0));
        int rc = h->group->run([&](dftgrid* s, int r) -> int { return dftgrid_last_timings(s, t[r].data(), DFTGRID_T_COUNT); });
        if (rc) return rc;
        for (int i = 0; i < cnt; i++) {
            out[i] = 0.0;
            for (auto& v : t) out[i] = std::max(out[i], v[i]);
        }
        return 0;
    }
    return guarded([&] {
        use_device(h);
        CK(cudaStreamSynchronize(h->stream));
        if (h->timed_iter) finish_timings(h);
        for (int i = 0; i < n && i < DFTGRID_T_COUNT; i++) out[i] = h->t_ms[i];
    });
}

int dftgrid_timer_start(dftgrid_t* h) {
    GROUP_FORWARD(h, dftgrid_timer_start(s));
    return guarded([&] {
        use_device(h);
        CK(cudaEventRecord(h->ev_sw[0], h->stream));
    });
}

int dftgrid_timer_stop(dftgrid_t* h, double* ms) {
    if (h->group) {
        std::vector<double> t(h->group->subs.size(), 0.0);
        int rc = h->group->run([&](dftgrid* s, int r) -> int { return dftgrid_timer_stop(s, &t[r]); });
        if (rc) return rc;
        *ms = *std::max_element(t.begin(), t.end());
        return 0;
    }
    return guarded([&] {
        use_device(h);
        CK(cudaEventRecord(h->ev_sw[1], h->stream));
        CK(cudaEventSynchronize(h->ev_sw[1]));
        float t = 0.f;
        CK(cudaEventElapsedTime(&t, h->ev_sw[0], h->ev_sw[1]));
        *ms = t;
    });
}

long dftgrid_launch_count(const dftgrid_t* h) {
    if (h->group) {
        long n = 0;
        for (const dftgrid* s : h->group->subs) n += s->launches;
        return n;
    }
    return h->launches;
}

int dftgrid_debug_contract_schedule_nz(int nbp, long nchunk, int nsm, int nz, int max_segs, int* segs_out, int* cta_off_out, int* nctas, int* nsegs,
                                       int* block_chunks) {
    return guarded([&] {
        if (nbp <= 0 || nbp % kNbAlign != 0 || nchunk < 0 || nsm <= 0 || nz < 1 || nz > 2 || !segs_out || !cta_off_out || !nctas || !nsegs || !block_chunks)
            throw std::runtime_error("dftgrid_debug_contract_schedule: bad arguments");
        ContractSchedule S;
        compute_contract_schedule(nbp, nchunk, nsm, nz, S);
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

int dftgrid_debug_contract_schedule(int nbp, long nchunk, int nsm, int max_segs, int* segs_out, int* cta_off_out, int* nctas, int* nsegs,
                                    int* block_chunks) {
    return dftgrid_debug_contract_schedule_nz(nbp, nchunk, nsm, 2, max_segs, segs_out, cta_off_out, nctas, nsegs, block_chunks);
}

int dftgrid_scf_init(dftgrid_t* h, const double* H, const double* X, int nocc, double alpha) {
    GROUP_FORWARD(h, dftgrid_scf_init(s, H, X, nocc, alpha));
    return guarded([&] {
        use_device(h);
        scf_init(h, H, X, nocc, alpha);
    });
}

int dftgrid_scf_step(dftgrid_t* h, int include_xc, double* out8) {
    if (h->group) {
        // every device repeats the (deterministic) algebra on its own copy: no traffic, identical bits
        return h->group->run([&](dftgrid* s, int r) -> int {
            double tmp[8];
            return dftgrid_scf_step(s, include_xc, r == 0 ? out8 : tmp);
        });
    }
    return guarded([&] {
        use_device(h);
        scf_step(h, include_xc, out8);
    });
}

int dftgrid_scf_get_matrix(dftgrid_t* h, int which, double* out) {
    if (h->group) return dftgrid_scf_get_matrix(h->group->subs[0], which, out);
    return guarded([&] {
        use_device(h);
        if (!h->scf.ready) throw std::runtime_error("dftgrid_scf_init has not been called");
        const int nb = h->nbf, np = h->scf.np;
        const size_t nb2 = (size_t)nb * nb;
        if (which == DFTGRID_SCF_P) {
            CK(cudaMemcpyAsync(out, h->d_Praw.p, nb2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        } else if (which == DFTGRID_SCF_FGRID) {
            CK(cudaMemcpyAsync(out, h->d_fres.p, nb2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        } else if (which == DFTGRID_SCF_FPRIME || which == DFTGRID_SCF_DPRIME) {
            const double* src = which == DFTGRID_SCF_FPRIME ? h->scf.Fp.p : h->scf.D.p;
            CK(cudaMemcpy2DAsync(out, (size_t)nb * sizeof(double), src, (size_t)np * sizeof(double), (size_t)nb * sizeof(double), nb,
                                 cudaMemcpyDeviceToHost, h->stream));
        } else {
            throw std::runtime_error("dftgrid_scf_get_matrix: unknown matrix id");
        }
        CK(cudaStreamSynchronize(h->stream));
    });
}

int dftgrid_debug_screen_fraction(dftgrid_t* h, double* fraction) {
    if (h->group) return dftgrid_debug_screen_fraction(h->group->subs[0], fraction);
    if (!fraction) return 1;
    *fraction = h->screened ? h->screen_work_fraction : 1.0;
    return 0;
}

int dftgrid_debug_set_stress(dftgrid_t* h, int mode) {
    GROUP_FORWARD(h, dftgrid_debug_set_stress(s, mode));
    return guarded([&] {
#ifndef DFG_STRESS
        if (mode != 0) throw std::runtime_error("this build has no stress hook (build libdftgrid_stress.so with -DDFG_STRESS)");
#endif
        use_device(h);
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaMemcpyToSymbol(c_stress_mode, &mode, sizeof mode));
    });
}

int dftgrid_host_register(void* p, size_t bytes) {
    return guarded([&] { CK(cudaHostRegister(p, bytes, cudaHostRegisterPortable)); });
}
int dftgrid_host_unregister(void* p) {
    return guarded([&] { CK(cudaHostUnregister(p)); });
}

}  // extern "C"
