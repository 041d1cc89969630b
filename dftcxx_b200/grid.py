"""ctypes view of the C ABI (include/dftgrid.h, built as dftcxx_b200/libdftgrid.so) with the method names
of the reference's MolecularGrid (reference src/moleculargrid.h:84-167).  Plumbing only: every number is
produced by the CUDA library; there is no Python or CPU compute path and a missing library is a hard error.
"""
import ctypes as C
import os

import numpy as np

from .molecule import LEBEDEV_COUNTS

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DFTGRID_LIB") or os.path.join(HERE, "libdftgrid.so")  # DFTGRID_LIB: developer A/B builds

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

T_NAMES = ["points", "becke", "phi", "rho", "xc_point", "rho_lm", "poisson", "interp", "contract", "comm", "total"]


class _System(C.Structure):
    _fields_ = [("natoms", C.c_int), ("Z", _ip), ("xyz", _dp), ("nbf", C.c_int), ("bf_nprim", _ip), ("bf_center", _dp),
                ("nprim", C.c_int), ("alpha", _dp), ("coeff", _dp), ("norm", _dp), ("lmn", _ip)]


class _Params(C.Structure):
    _fields_ = [("radial_points", C.c_int), ("lebedev_order", C.c_int), ("lmax", C.c_int)]


_lib = None


def lib():
    """Load libdftgrid.so (fails loudly if the CUDA extension has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("dftcxx_b200/libdftgrid.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.dftgrid_last_error.restype = C.c_char_p
    L.dftgrid_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(_System), C.POINTER(_Params), C.c_int, C.c_int, C.c_int]
    L.dftgrid_create_multi.argtypes = [C.POINTER(C.c_void_p), C.POINTER(_System), C.POINTER(_Params), C.c_int, _ip]
    L.dftgrid_ngpus.argtypes = [C.c_void_p]
    L.dftgrid_peer_set_timeout.argtypes = [C.c_void_p, C.c_double]
    L.dftgrid_fock.argtypes = [C.c_void_p, _dp, C.c_int, _dp, _dp, _dp, _dp]
    L.dftgrid_fock_device.argtypes = [C.c_void_p, C.c_int]
    L.dftgrid_download_fock.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
    L.dftgrid_scf_init.argtypes = [C.c_void_p, _dp, _dp, C.c_int, C.c_double]
    L.dftgrid_scf_step.argtypes = [C.c_void_p, C.c_int, _dp]
    L.dftgrid_scf_get_matrix.argtypes = [C.c_void_p, C.c_int, _dp]
    L.dftgrid_debug_set_stress.argtypes = [C.c_void_p, C.c_int]
    L.dftgrid_destroy.argtypes = [C.c_void_p]
    L.dftgrid_destroy.restype = None
    L.dftgrid_comm_unique_id.argtypes = [C.c_void_p]
    L.dftgrid_comm_init.argtypes = [C.c_void_p, C.c_void_p]
    L.dftgrid_peer_export.argtypes = [C.c_void_p, C.c_void_p]
    L.dftgrid_peer_connect.argtypes = [C.c_void_p, C.c_char_p]
    L.dftgrid_peer_active.argtypes = [C.c_void_p]
    L.dftgrid_peer_disable.argtypes = [C.c_void_p]
    L.dftgrid_timer_stop.argtypes = [C.c_void_p, _dp]
    for n in ("dftgrid_build", "dftgrid_iteration_device", "dftgrid_synchronize", "dftgrid_timer_start"):
        getattr(L, n).argtypes = [C.c_void_p]
    for n in ("dftgrid_npoints", "dftgrid_npoints_local", "dftgrid_point_offset", "dftgrid_launch_count"):
        getattr(L, n).argtypes = [C.c_void_p]
        getattr(L, n).restype = C.c_long
    for n in ("dftgrid_nbf", "dftgrid_nlm"):
        getattr(L, n).argtypes = [C.c_void_p]
    for n in ("dftgrid_set_density", "dftgrid_hartree_J", "dftgrid_electron_count", "dftgrid_upload_density",
              "dftgrid_get_positions", "dftgrid_get_weights", "dftgrid_get_becke_weights", "dftgrid_get_densities",
              "dftgrid_get_amplitudes", "dftgrid_get_potential", "dftgrid_get_rho_lm", "dftgrid_get_U_lm"):
        getattr(L, n).argtypes = [C.c_void_p, _dp]
    L.dftgrid_xc.argtypes = [C.c_void_p, _dp, _dp]
    L.dftgrid_iteration.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _dp]
    L.dftgrid_download_results.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
    L.dftgrid_last_timings.argtypes = [C.c_void_p, _dp, C.c_int]
    L.dftgrid_one_electron.argtypes = [C.c_void_p, _dp, _dp, _dp]
    L.dftgrid_rectangular_density.argtypes = [C.c_void_p, C.c_double, C.c_int, _dp, _dp, _dp, _dp]
    _lib = L
    return L


def _ptr(a, typ=_dp):
    return a.ctypes.data_as(typ)


class GridError(RuntimeError):
    pass


def shard_range(nshell_total, rank, nranks):
    """(first_shell, nshell) of a rank — the C library's own rule (dftgrid_shard_range)."""
    first, count = C.c_long(), C.c_long()
    L = lib()
    L.dftgrid_shard_range.argtypes = [C.c_long, C.c_int, C.c_int, C.POINTER(C.c_long), C.POINTER(C.c_long)]
    if L.dftgrid_shard_range(nshell_total, rank, nranks, C.byref(first), C.byref(count)) != 0:
        raise GridError(L.dftgrid_last_error().decode())
    return first.value, count.value


def comm_unique_id():
    buf = C.create_string_buffer(128)
    if lib().dftgrid_comm_unique_id(buf) != 0:
        raise GridError(lib().dftgrid_last_error().decode())
    return buf.raw


class MolecularGrid:
    """B200 grid engine handle.  mol: anything with Z, xyz, bf_nprim, bf_center, alpha, coeff, norm, lmn arrays
    (dftcxx_b200.molecule.Molecule or the dict oracle/refpy.Ref.system() returns)."""

    def __init__(self, mol, device=0, rank=0, nranks=1, ngpus=1, devices=None):
        """ngpus > 1: ONE handle driving that many devices of this box from this process (dftgrid_create_multi);
        rank/nranks: one process per GPU (dftgrid_create + comm_id / connect_peers)."""
        get = (lambda k: mol[k]) if isinstance(mol, dict) else (lambda k: getattr(mol, k))
        self._keep = dict(
            Z=np.ascontiguousarray(get("Z"), dtype=np.int32), xyz=np.ascontiguousarray(get("xyz"), dtype=np.float64),
            bf_nprim=np.ascontiguousarray(get("bf_nprim"), dtype=np.int32),
            bf_center=np.ascontiguousarray(get("bf_center"), dtype=np.float64),
            alpha=np.ascontiguousarray(get("alpha"), dtype=np.float64), coeff=np.ascontiguousarray(get("coeff"), dtype=np.float64),
            norm=np.ascontiguousarray(get("norm"), dtype=np.float64), lmn=np.ascontiguousarray(get("lmn"), dtype=np.int32))
        self.natoms = len(self._keep["Z"])
        self.nbf = len(self._keep["bf_nprim"])
        self.nprim = len(self._keep["alpha"])
        self.device, self.rank, self.nranks = device, rank, nranks
        self.ngpus, self.devices = int(ngpus), devices
        if self.ngpus > 1 and nranks > 1:
            raise GridError("ngpus > 1 (single process) and nranks > 1 (one process per GPU) are exclusive")
        self.h = None
        self.radial_points = self.lebedev_order = self.lmax = None

    # -- reference surface --------------------------------------------------------------------------------
    def set_grid_parameters(self, radial_points, lebedev_order, lmax):
        """MolecularGrid::set_grid_parameters (src/moleculargrid.cpp:175-179)."""
        self.radial_points, self.lebedev_order, self.lmax = int(radial_points), int(lebedev_order), int(lmax)

    def create_grid(self, comm_id=None):
        """MolecularGrid::create_grid (src/moleculargrid.cpp:193-261)."""
        if self.radial_points is None:
            raise GridError("set_grid_parameters has not been called")
        k = self._keep
        sysd = _System(self.natoms, _ptr(k["Z"], _ip), _ptr(k["xyz"]), self.nbf, _ptr(k["bf_nprim"], _ip), _ptr(k["bf_center"]),
                       self.nprim, _ptr(k["alpha"]), _ptr(k["coeff"]), _ptr(k["norm"]), _ptr(k["lmn"], _ip))
        prm = _Params(self.radial_points, self.lebedev_order, self.lmax)
        h = C.c_void_p()
        if self.ngpus > 1:
            devs = None if self.devices is None else _ptr(np.ascontiguousarray(self.devices, dtype=np.int32), _ip)
            self._ck(lib().dftgrid_create_multi(C.byref(h), C.byref(sysd), C.byref(prm), self.ngpus, devs))
        else:
            self._ck(lib().dftgrid_create(C.byref(h), C.byref(sysd), C.byref(prm), self.device, self.rank, self.nranks))
        self.close()
        self.h = h
        try:
            if self.nranks > 1 and comm_id is not False:  # comm_id=False: developer profiling of one shard without NCCL
                if comm_id is None:
                    raise GridError("nranks > 1 needs the NCCL unique id from rank 0")
                self._ck(lib().dftgrid_comm_init(self.h, comm_id))
            self._ck(lib().dftgrid_build(self.h))
        except GridError:
            self.close()
            raise
        self.npoints = lib().dftgrid_npoints(self.h)
        self.nloc = lib().dftgrid_npoints_local(self.h)
        self.point_offset = lib().dftgrid_point_offset(self.h)
        self.nlm = lib().dftgrid_nlm(self.h)
        self.nang = LEBEDEV_COUNTS[self.lebedev_order]

    # -- peer-memory reduction (multi-GPU, optional) --------------------------------------------------------
    def peer_export(self):
        """This rank's 64-byte CUDA IPC handle of its [J | XC] exchange buffer (dftgrid_peer_export)."""
        buf = C.create_string_buffer(64)
        self._ck(lib().dftgrid_peer_export(self.h, buf))
        return buf.raw

    def peer_connect(self, handles):
        """handles: every rank's peer_export() blob in rank order.  Returns True when the peer path is active; on a
        mapping failure (no P2P path) the library keeps using NCCL and False is returned."""
        blob = b"".join(handles)
        if len(blob) != 64 * self.nranks:
            raise GridError("peer_connect needs one 64-byte handle per rank")
        rc = lib().dftgrid_peer_connect(self.h, blob)
        return rc == 0 and lib().dftgrid_peer_active(self.h) == 1

    def peer_disable(self):
        """Back to ncclAllReduce for the [J | XC] sum.  Every rank must make the same choice: call this on all ranks when
        peer_connect returned False on any of them (see connect_peers)."""
        self._ck(lib().dftgrid_peer_disable(self.h))

    def connect_peers(self, dist):
        """Collective helper for torch.distributed hosts: exchange the IPC handles, map them, and keep the peer-memory
        path only if the mapping succeeded on EVERY rank.  Returns True when the peer path is active."""
        import torch

        hs = [None] * self.nranks
        dist.all_gather_object(hs, self.peer_export())
        ok = self.peer_connect(hs)
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda:%d" % self.device if dist.get_backend() == "nccl" else "cpu")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            if ok:
                self.peer_disable()
            return False
        return True

    def set_density(self, P):
        """set_density + correct_densities (src/moleculargrid.cpp:48-53,132-146)."""
        self._ck(lib().dftgrid_set_density(self.h, _ptr(self._mat(P))))

    def correct_densities(self):
        """Folded into set_density (the reference always calls them back to back, src/dft.cpp:362-365)."""

    def calculate_hartree_potential(self):
        """MolecularGrid::calculate_hartree_potential (src/moleculargrid.cpp:336-389) -> J."""
        J = np.zeros((self.nbf, self.nbf))
        self._ck(lib().dftgrid_hartree_J(self.h, _ptr(J)))
        return J

    def calculate_exchange_correlation(self):
        """DFT::calculate_exchange_correlation_matrix (src/dft.cpp:394-433) -> (XC, E_xc)."""
        XC = np.zeros((self.nbf, self.nbf))
        exc = C.c_double()
        self._ck(lib().dftgrid_xc(self.h, _ptr(XC), C.cast(C.byref(exc), _dp)))
        return XC, exc.value

    def calculate_density(self):
        """MolecularGrid::calculate_density (src/moleculargrid.cpp:157-166)."""
        n = C.c_double()
        self._ck(lib().dftgrid_electron_count(self.h, C.cast(C.byref(n), _dp)))
        return n.value

    def iteration(self, P, out=None):
        """One SCF iteration's grid work through the single-call entry point -> (J, XC, E_xc, N_el).
        out = (J, XC): caller-owned result arrays to fill (C-contiguous float64 nbf x nbf, e.g. views of pinned host
        memory, which the library then DMAs into directly); fresh arrays otherwise."""
        if out is None:
            J = np.empty((self.nbf, self.nbf))
            XC = np.empty((self.nbf, self.nbf))
        else:
            J, XC = out
            for a in (J, XC):
                if a.dtype != np.float64 or a.shape != (self.nbf, self.nbf) or not a.flags.c_contiguous:
                    raise ValueError("out arrays must be C-contiguous float64 of shape (nbf, nbf)")
        exc, nel = C.c_double(), C.c_double()
        self._ck(lib().dftgrid_iteration(self.h, _ptr(self._mat(P)), _ptr(J), _ptr(XC), C.cast(C.byref(exc), _dp),
                                         C.cast(C.byref(nel), _dp)))
        return J, XC, exc.value, nel.value

    def fock(self, P, include_xc=True, out=None):
        """Fused Fock contribution (dftgrid_fock) -> (F_grid = 2J + XC, E_J, E_xc, N_el); out = caller-owned F array."""
        F = np.empty((self.nbf, self.nbf)) if out is None else out
        if F.dtype != np.float64 or F.shape != (self.nbf, self.nbf) or not F.flags.c_contiguous:
            raise ValueError("out must be C-contiguous float64 of shape (nbf, nbf)")
        ej, exc, nel = C.c_double(), C.c_double(), C.c_double()
        self._ck(lib().dftgrid_fock(self.h, _ptr(self._mat(P)), 1 if include_xc else 0, _ptr(F), C.cast(C.byref(ej), _dp),
                                    C.cast(C.byref(exc), _dp), C.cast(C.byref(nel), _dp)))
        return F, ej.value, exc.value, nel.value

    def fock_device(self, include_xc=True):
        self._ck(lib().dftgrid_fock_device(self.h, 1 if include_xc else 0))

    def download_fock(self):
        F = np.zeros((self.nbf, self.nbf))
        ej, exc, nel = C.c_double(), C.c_double(), C.c_double()
        self._ck(lib().dftgrid_download_fock(self.h, _ptr(F), C.cast(C.byref(ej), _dp), C.cast(C.byref(exc), _dp), C.cast(C.byref(nel), _dp)))
        return F, ej.value, exc.value, nel.value

    # -- device-resident SCF algebra (dftgrid_scf_*) -------------------------------------------------------
    def scf_init(self, H, X, nocc, alpha=0.5):
        """H: core Hamiltonian, X: orthogonalisation matrix U s^-1/2 (both nbf x nbf; X[i, j] = row i, column j)."""
        self._ck(lib().dftgrid_scf_init(self.h, _ptr(self._mat(H)), _ptr(self._mat(X)), int(nocc), float(alpha)))

    def scf_step(self, include_xc=True):
        """One SCF loop body on the device -> dict(e_one, e_j, exc, nel, purification_steps, idempotency, ms_algebra, ms_grid)."""
        o = np.zeros(8)
        self._ck(lib().dftgrid_scf_step(self.h, 1 if include_xc else 0, _ptr(o)))
        return dict(e_one=o[0], e_j=o[1], exc=o[2], nel=o[3], purification_steps=int(o[4]), idempotency=o[5], ms_algebra=o[6], ms_grid=o[7])

    def scf_matrix(self, which):
        """which: 'P' (mixed density matrix), 'F_grid', 'F_prime' (X^T F X), 'D_prime' (purified projector)."""
        out = np.zeros((self.nbf, self.nbf))
        self._ck(lib().dftgrid_scf_get_matrix(self.h, {"P": 0, "F_grid": 1, "F_prime": 2, "D_prime": 3}[which], _ptr(out)))
        return out

    def screen_fraction(self):
        """Mean fraction of a full contraction stage's tensor work left after screening (1.0 = nothing skipped)."""
        f = C.c_double()
        lib().dftgrid_debug_screen_fraction.argtypes = [C.c_void_p, _dp]
        self._ck(lib().dftgrid_debug_screen_fraction(self.h, C.cast(C.byref(f), _dp)))
        return f.value

    def debug_set_stress(self, mode):
        """Test hook: random delays in the producer (1) / consumer (2) warps of the tensor kernels' pipelines."""
        self._ck(lib().dftgrid_debug_set_stress(self.h, int(mode)))

    def peer_active(self):
        return lib().dftgrid_peer_active(self.h) == 1

    # -- device-resident path (benchmarks) ---------------------------------------------------------------
    def upload_density(self, P):
        self._ck(lib().dftgrid_upload_density(self.h, _ptr(self._mat(P))))

    def iteration_device(self):
        self._ck(lib().dftgrid_iteration_device(self.h))

    def synchronize(self):
        self._ck(lib().dftgrid_synchronize(self.h))

    def timer_start(self):
        self._ck(lib().dftgrid_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double()
        self._ck(lib().dftgrid_timer_stop(self.h, C.cast(C.byref(ms), _dp)))
        return ms.value

    def download_results(self):
        J = np.zeros((self.nbf, self.nbf))
        XC = np.zeros((self.nbf, self.nbf))
        exc, nel = C.c_double(), C.c_double()
        self._ck(lib().dftgrid_download_results(self.h, _ptr(J), _ptr(XC), C.cast(C.byref(exc), _dp), C.cast(C.byref(nel), _dp)))
        return J, XC, exc.value, nel.value

    # -- getters -------------------------------------------------------------------------------------------
    def _vec(self, fn, shape):
        out = np.zeros(shape)
        self._ck(fn(self.h, _ptr(out)))
        return out

    def get_positions(self):
        return self._vec(lib().dftgrid_get_positions, (self.nloc, 3))

    def get_weights(self):
        return self._vec(lib().dftgrid_get_weights, self.nloc)

    def get_becke_weights(self):
        return self._vec(lib().dftgrid_get_becke_weights, self.nloc)

    def get_densities(self):
        return self._vec(lib().dftgrid_get_densities, self.nloc)

    def get_amplitudes(self):
        """[nloc][nbf] (the reference returns the transpose, basis functions x points, src/moleculargrid.cpp:109-127)."""
        return self._vec(lib().dftgrid_get_amplitudes, (self.nloc, self.nbf))

    def get_potential(self):
        return self._vec(lib().dftgrid_get_potential, self.nloc)

    def get_rho_lm(self):
        return self._vec(lib().dftgrid_get_rho_lm, (self.natoms, self.radial_points, self.nlm))

    def get_U_lm(self):
        return self._vec(lib().dftgrid_get_U_lm, (self.natoms, self.radial_points, self.nlm))

    def one_electron(self):
        """Overlap, kinetic-energy and nuclear-attraction matrices of the handle's basis (DFT::construct_matrices,
        src/dft.cpp:185-198) from the device."""
        S, T, V = (np.zeros((self.nbf, self.nbf)) for _ in range(3))
        self._ck(lib().dftgrid_one_electron(self.h, _ptr(S), _ptr(T), _ptr(V)))
        return S, T, V

    def rectangular_density(self, size, dp, P):
        """RectangularGrid::build_grid(size, dp) + set_density(P) (src/rectangulargrid.cpp:34-80): positions [dp^3][3], density
        [dp^3] and density gradient [dp^3][3] on a box of edge `size` centred at the origin of the molecule's frame."""
        n = int(dp) ** 3
        pos, rho, grad = np.zeros((n, 3)), np.zeros(n), np.zeros((n, 3))
        self._ck(lib().dftgrid_rectangular_density(self.h, float(size), int(dp), _ptr(self._mat(P)), _ptr(pos), _ptr(rho), _ptr(grad)))
        return pos, rho, grad

    def timings(self):
        t = np.zeros(len(T_NAMES))
        self._ck(lib().dftgrid_last_timings(self.h, _ptr(t), len(T_NAMES)))
        return dict(zip(T_NAMES, t.tolist()))

    def launch_count(self):
        return lib().dftgrid_launch_count(self.h)

    # -- plumbing ------------------------------------------------------------------------------------------
    def _mat(self, M):
        M = np.ascontiguousarray(M, dtype=np.float64)
        if M.shape != (self.nbf, self.nbf):
            raise GridError("matrix must be nbf x nbf")
        return M

    def _ck(self, rc):
        if rc != 0:
            raise GridError(lib().dftgrid_last_error().decode())

    def close(self):
        if self.h:
            lib().dftgrid_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
