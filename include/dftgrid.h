/*
 * dftgrid.h — C ABI of the B200-native numerical-grid engine for dftcxx.
 *
 * This is the drop-in boundary for the per-SCF-iteration grid hot path.  The
 * reference (ifilot/dftcxx) has no FFI; the boundary is the public surface of
 * its MolecularGrid class as used by DFT (reference src/moleculargrid.h:84-167,
 * call sites src/dft.cpp:57-61, 362-365, 483, 395-432, 111).  Every entry point
 * below names the reference interface it replaces.
 *
 * Conventions
 *  - plain C types only; all buffers are caller-owned HOST memory, FP64.
 *  - matrices are nb x nb, symmetric, column-major (Eigen's MatrixXd layout).
 *  - point order is the reference's: atom-major, then radial index p-1 (r descending),
 *    then Lebedev index (src/atomicgrid.cpp:50-82).  Basis-function order is whatever
 *    the caller passes (the reference's is by element, src/molecule.cpp:222-235).
 *  - every call returns 0 on success, non-zero on failure; dftgrid_last_error() then
 *    holds a message (the reference throws std::runtime_error; the C++ host wrapper
 *    in dftcxx_b200/host rethrows).  No C++ exception crosses this boundary.
 *  - a handle is not re-entrant; calls come from one host thread (src/dft.cpp:95-104).
 *  - there is NO CPU fallback: without a CUDA device every compute call fails.
 *
 * Multi-GPU, two ways.  (1) Single process (what `dftcxx -i ... --gpus N` uses, reference src/dft.cpp:57-61 has one
 * MolecularGrid per run): dftgrid_create_multi returns ONE handle that drives N devices of the box with one worker
 * thread per device; every call below accepts it and behaves as on a whole-grid handle.  (2) One process per GPU
 * ("rank"): dftgrid_create(rank, nranks) + dftgrid_comm_* / dftgrid_peer_* wire NCCL and the peer-memory exchange
 * between the handles of one job (ids are exchanged by the host, e.g. torch.distributed).  Grid points are sharded
 * by (atom, radial shell) units either way.
 */
#ifndef DFTGRID_H
#define DFTGRID_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dftgrid dftgrid_t;

/* Molecule + basis, flattened.  Replaces the Molecule/Atom/CGF/GTO accessors the reference grid code
 * reads (src/molecule.h:111-175, src/cgf.h:41-314). */
typedef struct dftgrid_system {
    int natoms;
    const int* Z;             /* [natoms] nuclear charges                                   */
    const double* xyz;        /* [natoms][3] positions in bohr                               */
    int nbf;                  /* number of contracted basis functions (CGFs)                 */
    const int* bf_nprim;      /* [nbf] primitives per CGF                                    */
    const double* bf_center;  /* [nbf][3] centre of each CGF (bohr)                          */
    int nprim;                /* total primitives = sum(bf_nprim)                            */
    const double* alpha;      /* [nprim] exponents, CGF after CGF                            */
    const double* coeff;      /* [nprim] contraction coefficients                            */
    const double* norm;       /* [nprim] GTO normalisation constants (src/cgf.cpp:102-114)    */
    const int* lmn;           /* [nprim][3] Cartesian powers, l+m+n <= 2 (src/cgf.cpp:185-232) */
} dftgrid_system;

/* MolecularGrid::set_grid_parameters (src/moleculargrid.cpp:175-179; presets src/settings.cpp:158-187) */
typedef struct dftgrid_params {
    int radial_points;  /* Gauss-Chebyshev nodes per atom                                    */
    int lebedev_order;  /* index 0..10 into {6,14,26,38,50,74,86,110,146,170,194}            */
    int lmax;           /* highest l of the multipole expansion of the Hartree potential     */
} dftgrid_params;

const char* dftgrid_last_error(void);
/* ABI version of this header; bump on any signature change. */
int dftgrid_abi_version(void);

/* MolecularGrid::MolecularGrid + set_grid_parameters (src/moleculargrid.cpp:32-35,175-179).
 * device: CUDA ordinal.  rank/nranks: this handle's shard of the grid (0/1 = whole grid). */
int dftgrid_create(dftgrid_t** h, const dftgrid_system* sys, const dftgrid_params* prm, int device, int rank, int nranks);
/* The same for ngpus devices of this box driven from one process (`int ngpus`): one handle, the
 * grid sharded over devices[0..ngpus) (NULL = ordinals 0..ngpus-1), NCCL communicators from ncclCommInitAll for the small
 * per-shell sums, the [J | XC] / F sum over peer memory (cudaDeviceEnablePeerAccess) when every device pair has a P2P
 * route.  ngpus = 1 is dftgrid_create(device, 0, 1).  Results come back identical to the bit on every device; matrix
 * downloads are split row-wise over the devices' PCIe links. */
int dftgrid_create_multi(dftgrid_t** h, const dftgrid_system* sys, const dftgrid_params* prm, int ngpus, const int* devices);
/* Number of devices behind a handle (1 unless it came from dftgrid_create_multi). */
int dftgrid_ngpus(const dftgrid_t* h);
/* RAII teardown of unique_ptr<MolecularGrid> (src/dft.h:44). */
void dftgrid_destroy(dftgrid_t* h);

/* The shard of rank `rank` out of `nranks`: a contiguous block of the natoms*nrad (atom, radial shell) units, balanced by
 * count.  Pure host arithmetic (no device needed).  dftgrid_build uses this rule when the screening of Phi is off; with
 * screening (the default) it cuts the same contiguous blocks at equal estimated WORK instead (a shell far from most atoms
 * is cheap; see shard_shells in csrc/dftgrid_api.cu) — dftgrid_point_offset / dftgrid_npoints_local report the result. */
int dftgrid_shard_range(long nshell_total, int rank, int nranks, long* first_shell, long* nshell);

/* NCCL wiring for nranks > 1.  id is a 128-byte opaque blob produced on rank 0 and handed to every rank. */
int dftgrid_comm_unique_id(void* id128);
int dftgrid_comm_init(dftgrid_t* h, const void* id128);

/* Optional peer-memory path for the last collective of an iteration (nranks 2..16, one process per GPU, GPUs with a P2P
 * path — NVLink / NVSwitch): the sum of the ranks' [J | XC] partial matrices is then done by the library's own kernels
 * reading the other ranks' buffers directly (k_contract_reduce_publish + k_peer_sum) instead of ncclAllReduce; same
 * rank-ordered sum on every rank, so results stay bit-identical across ranks.  dftgrid_peer_export writes this rank's
 * 64-byte CUDA IPC handle; the host gathers all ranks' handles (rank order, 64 bytes each) and hands them to
 * dftgrid_peer_connect on every rank.  Without these calls, or if the mapping fails, NCCL is used.  (No reference
 * counterpart: the reference is single-process.) */
int dftgrid_peer_export(dftgrid_t* h, void* handle64);
int dftgrid_peer_connect(dftgrid_t* h, const void* handles /* [nranks][64] */);
int dftgrid_peer_active(const dftgrid_t* h);
/* The choice between the peer-memory path and NCCL must be the same on every rank: if dftgrid_peer_connect failed on
 * ANY rank (the host checks, e.g. with an all-reduce(min) of the return codes), every rank calls dftgrid_peer_disable. */
int dftgrid_peer_disable(dftgrid_t* h);
/* Bound of the peer kernels' spin-waits in seconds of wall time (default 60): a rank that does not show up within it
 * makes the waiting ranks fail with an error instead of hanging their GPUs.  The error is fatal for the handle. */
int dftgrid_peer_set_timeout(dftgrid_t* h, double seconds);

/* MolecularGrid::create_grid (src/moleculargrid.cpp:193-261): points, quadrature weights, CGF amplitudes,
 * Becke weights; also factorises the radial Poisson operators used by dftgrid_hartree_J. */
int dftgrid_build(dftgrid_t* h);

/* sizes */
long dftgrid_npoints(const dftgrid_t* h);        /* whole molecule                                     */
long dftgrid_npoints_local(const dftgrid_t* h);  /* this rank's shard                                  */
long dftgrid_point_offset(const dftgrid_t* h);   /* global index of the first local point              */
int dftgrid_nbf(const dftgrid_t* h);
int dftgrid_nlm(const dftgrid_t* h);

/* MolecularGrid::set_density + correct_densities (src/moleculargrid.cpp:48-53,132-146; src/dft.cpp:362-365):
 * rho_p = 2 phi_p^T P phi_p, then rho *= sum(Z)/sum(w rho). */
int dftgrid_set_density(dftgrid_t* h, const double* P);
/* MolecularGrid::calculate_hartree_potential (src/moleculargrid.cpp:336-389): J from the current density. */
int dftgrid_hartree_J(dftgrid_t* h, double* J);
/* DFT::calculate_exchange_correlation_matrix (src/dft.cpp:394-433) incl. Functional::xalpha_x_functional /
 * vwm_c_functional (src/functionals.cpp:24-114): XC matrix and E_xc. */
int dftgrid_xc(dftgrid_t* h, double* XC, double* exc);
/* MolecularGrid::calculate_density (src/moleculargrid.cpp:157-166): sum(w rho). */
int dftgrid_electron_count(dftgrid_t* h, double* nelec);

/* One SCF iteration's whole grid path in one call (set_density, hartree_J, xc, electron_count with a single
 * upload of P and a single download of [J | XC | exc | nelec]); same results as the four calls above. */
int dftgrid_iteration(dftgrid_t* h, const double* P, double* J, double* XC, double* exc, double* nelec);

/* Fused Fock-matrix contribution of the grid (the only way src/dft.cpp:334 uses J and XC is F = H + 2J + XC):
 * F_grid = 2J + XC = Phi^T diag(w (V_Hartree + v_xc)) Phi as ONE symmetric contraction (half the tensor work of
 * dftgrid_iteration), E_J = 2 tr(P J) (src/dft.cpp:443) from the pointwise identity 1/2 sum_p w_p V_p rho_p[P], E_xc and
 * the electron count.  include_xc = 0 leaves XC out (the reference's first iteration builds F from J(P0) alone,
 * src/dft.cpp:219-226).  Same density / potential kernels as the calls above; F agrees with 2J + XC to rounding. */
int dftgrid_fock(dftgrid_t* h, const double* P, int include_xc, double* F, double* e_j, double* exc, double* nelec);
int dftgrid_fock_device(dftgrid_t* h, int include_xc);
int dftgrid_download_fock(dftgrid_t* h, double* F, double* e_j, double* exc, double* nelec);

/* Device-resident SCF algebra: DFT::calculate_density_matrix + DFT::calculate_energy
 * (src/dft.cpp:330-366, 441-447) without leaving the GPU.  dftgrid_scf_init uploads the core Hamiltonian H and the
 * orthogonalisation matrix X = U s^-1/2 (src/dft.cpp:300-316; both nb x nb ROW-major, X is not symmetric) once.  Every
 * dftgrid_scf_step then forms F = H + F_grid of the previous step (F = H before the first one: core guess), F' = X^T F X,
 * the projector D' onto the nocc lowest eigenvectors of F' (Palser-Manolopoulos purification: FP64 tensor-core products
 * only, see csrc/kernels_scf.cuh), Pnew = X D' X^T, the reference's 50 % mixing P = (1 - alpha) Pnew + alpha P (first
 * density unmixed), and runs the grid path for the new P as dftgrid_fock_device(include_xc).  P, F, H, X stay in HBM; the
 * host receives out8 = { E_one = 2 tr(P H), E_J, E_xc, electron count, purification steps, idempotency residual,
 * device ms of the algebra, device ms of the grid path }.  The reference's sequence is one step with include_xc = 0
 * (DFT::construct_matrices, src/dft.cpp:219-226) followed by steps with include_xc = 1 (src/dft.cpp:100-103).  Fails
 * (non-zero) when F' has no gap at the Fermi level; the caller then uses its own eigen-solver with dftgrid_fock. */
int dftgrid_scf_init(dftgrid_t* h, const double* H, const double* X, int nocc, double alpha);
int dftgrid_scf_step(dftgrid_t* h, int include_xc, double* out8);
enum { DFTGRID_SCF_P = 0, DFTGRID_SCF_FGRID = 1, DFTGRID_SCF_FPRIME = 2, DFTGRID_SCF_DPRIME = 3 };
int dftgrid_scf_get_matrix(dftgrid_t* h, int which, double* out /* nb x nb, row-major */);

/* Page-lock / release a caller buffer (cudaHostRegister, portable across the devices of a multi-GPU handle) so that the
 * matrix uploads and downloads above DMA straight from / into it; for hosts that do not link the CUDA runtime. */
int dftgrid_host_register(void* p, size_t bytes);
int dftgrid_host_unregister(void* p);

/* Device-resident variant for benchmarking: P already uploaded by dftgrid_upload_density; runs all kernels and the
 * collectives, leaves results on the device; dftgrid_download_results copies them out. */
int dftgrid_upload_density(dftgrid_t* h, const double* P);
int dftgrid_iteration_device(dftgrid_t* h);
int dftgrid_download_results(dftgrid_t* h, double* J, double* XC, double* exc, double* nelec);
int dftgrid_synchronize(dftgrid_t* h);

/* Getters mirroring MolecularGrid::get_weights / get_densities / get_amplitudes (src/moleculargrid.cpp:61-127)
 * plus parity/debug views of the intermediates.  All return this rank's LOCAL points, in order, except the
 * per-atom tables which are whole-molecule. */
int dftgrid_get_positions(dftgrid_t* h, double* xyz /* [nloc][3] */);
int dftgrid_get_weights(dftgrid_t* h, double* w /* [nloc] */);
int dftgrid_get_becke_weights(dftgrid_t* h, double* wb /* [nloc] */);
int dftgrid_get_densities(dftgrid_t* h, double* rho /* [nloc] */);
int dftgrid_get_amplitudes(dftgrid_t* h, double* phi /* [nloc][nbf], point-major */);
int dftgrid_get_potential(dftgrid_t* h, double* V /* [nloc] Hartree potential */);
int dftgrid_get_rho_lm(dftgrid_t* h, double* rho_lm /* [natoms][nrad][nlm] */);
int dftgrid_get_U_lm(dftgrid_t* h, double* U_lm /* [natoms][nrad][nlm] */);

/* The one-electron integrals of DFT::construct_matrices (src/dft.cpp:185-198; Integrator::overlap / kinetic / nuclear,
 * src/integrals.cpp:43-387) over the handle's basis and nuclei: overlap S, kinetic energy T and nuclear attraction V
 * (summed over all nuclei with their charges), nb x nb each, symmetric.  McMurchie-Davidson on the device with the
 * reference's numerical conventions (pi = 3.14159265359 in the nuclear prefactor, Boys argument clamped at 1e-8).
 * One-time work of a run.  Needs a built handle; any output may be NULL. */
int dftgrid_one_electron(dftgrid_t* h, double* S, double* T, double* V);

/* RectangularGrid::build_grid(size, dp) + set_density(P) (src/rectangulargrid.cpp:34-80) — the data DFT::finalize's density
 * dump writes (src/dft.cpp:489-504, commented out in the reference "needs to be connected to interface"): dp^3 points,
 * point (i, j, k) at index (i*dp + j)*dp + k and position ((k, j, i) * size/(dp-1) - size/2) in the molecule's frame,
 * rho = 2 phi^T P phi (src/gridpoint.cpp:82-84) and the density gradient of GridPoint::set_gradient (src/gridpoint.cpp:94-109)
 * with the basis-function gradients exactly as CGF::get_grad evaluates them (src/cgf.cpp:67-94, 164-172).  Needs a built
 * handle (device basis tables); any of pos [n][3], rho [n], grad [n][3] may be NULL. */
int dftgrid_rectangular_density(dftgrid_t* h, double size, int dp, const double* P, double* pos, double* rho, double* grad);

/* Per-phase device times (CUDA events on the handle's stream) of the most recent build / iteration, in ms.
 * Slots: see DFTGRID_T_* below.  n = number of doubles available in out. */
int dftgrid_last_timings(dftgrid_t* h, double* out, int n);
/* Device stopwatch on the handle's stream (CUDA events): start records an event now, stop records a second one, waits
 * for it and returns the elapsed device time in ms.  Used by bench.py to time K iterations where the work is queued. */
int dftgrid_timer_start(dftgrid_t* h);
int dftgrid_timer_stop(dftgrid_t* h, double* ms);
/* Kernel launches issued by this handle since creation. */
long dftgrid_launch_count(const dftgrid_t* h);

/* Test hook, pure host arithmetic (no device needed): the stream-K schedule the [J | XC] contraction would use for a padded
 * basis size nbp (multiple of 32), nchunk non-zero 32-point chunks and nsm SMs.  segs_out: [nsegs][4] = (matrix, tile pair,
 * begin, end) with begin/end 31-bit fixed-point fractions of the item's chunks; cta_off_out: [nctas+1] (caller provides
 * nsm+1 ints).  Chunk x belongs to the segment whose [begin, end) holds ((x * 2654435769) mod 2^32) >> 1. */
int dftgrid_debug_contract_schedule(int nbp, long nchunk, int nsm, int max_segs, int* segs_out, int* cta_off_out, int* nctas, int* nsegs,
                                    int* block_chunks);
/* The same for nz = 1 (fused Fock build) or 2 ([XC | J]) matrices. */
int dftgrid_debug_contract_schedule_nz(int nbp, long nchunk, int nsm, int nz, int max_segs, int* segs_out, int* cta_off_out, int* nctas,
                                       int* nsegs, int* block_chunks);

/* Screening statistics of the built grid: the mean fraction of a full contraction stage's tensor work that a chunk costs a
 * tile pair under the block map of Phi (1 = nothing skipped; see csrc/kernels_dense.cuh k_chunk_masks).  The threshold on
 * |phi| is 1e-20 unless the developer switch DFTGRID_SCREEN_TAU says otherwise (0 = exact zeros only, negative = off). */
int dftgrid_debug_screen_fraction(dftgrid_t* h, double* fraction);

/* Test hook: perturb the timing of the producer (bit 0) / consumer (bit 1) warps of the two tensor kernels' mbarrier
 * pipelines with pseudo-random delays on this handle's device(s); results must not change by a bit.  0 = off (default). */
int dftgrid_debug_set_stress(dftgrid_t* h, int mode);

enum {
    DFTGRID_T_POINTS = 0,   /* build: points + raw weights            */
    DFTGRID_T_BECKE = 1,    /* build: Becke fuzzy-cell weights        */
    DFTGRID_T_PHI = 2,      /* build: CGF amplitudes                  */
    DFTGRID_T_RHO = 3,      /* iteration: rho = 2 rowsum((Phi P) o Phi) */
    DFTGRID_T_XCPOINT = 4,  /* iteration: charge sums, rescale, LDA pointwise */
    DFTGRID_T_RHOLM = 5,    /* iteration: Ylm projection              */
    DFTGRID_T_POISSON = 6,  /* iteration: radial solves, own-cell V, splines */
    DFTGRID_T_INTERP = 7,   /* iteration: cross-atom interpolation    */
    DFTGRID_T_CONTRACT = 8, /* iteration: [J | XC] contraction + reduction */
    DFTGRID_T_COMM = 9,     /* iteration: collectives                 */
    DFTGRID_T_TOTAL = 10,   /* iteration: first kernel to last        */
    DFTGRID_T_COUNT = 11
};

#ifdef __cplusplus
}
#endif
#endif /* DFTGRID_H */
