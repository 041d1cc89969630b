"""TEST INFRASTRUCTURE ONLY: ctypes view of oracle/_ref/libdftref*.so (the unmodified
reference classes behind oracle/ref_harness.cpp).  Imported by tests/, bench.py's
cpu_baseline / --impl reference legs and __graft_entry__.smoke() only."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DATA = os.path.join(ROOT, "dftcxx_b200", "data")
RUNDIR = os.path.join(DATA, "molecules")  # ../basis relative to this holds the .dat files

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def available(fast=False):
    return os.path.exists(os.path.join(HERE, "_ref", "libdftref_fast.so" if fast else "libdftref.so"))


_libs = {}


def _lib(fast=False):
    if fast in _libs:
        return _libs[fast]
    path = os.path.join(HERE, "_ref", "libdftref_fast.so" if fast else "libdftref.so")
    L = C.CDLL(path)
    L.ref_open.restype = C.c_void_p
    L.ref_open.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
    L.ref_last_error.restype = C.c_char_p
    L.ref_close.argtypes = [C.c_void_p]
    for n in ("ref_natoms", "ref_nbf", "ref_nprims", "ref_nrad", "ref_nang", "ref_lebedev_order", "ref_lmax", "ref_nelec"):
        getattr(L, n).restype = C.c_int
        getattr(L, n).argtypes = [C.c_void_p]
    L.ref_npoints.restype = C.c_long
    L.ref_npoints.argtypes = [C.c_void_p]
    L.ref_get_atoms.argtypes = [C.c_void_p, _ip, _dp]
    L.ref_get_basis.argtypes = [C.c_void_p, _ip, _dp, _dp, _dp, _dp, _ip]
    L.ref_get_grid.argtypes = [C.c_void_p, _dp, _dp, _dp]
    L.ref_get_amplitudes.argtypes = [C.c_void_p, _dp]
    L.ref_set_density.argtypes = [C.c_void_p, _dp]
    L.ref_set_density_raw.argtypes = [C.c_void_p, _dp]
    L.ref_set_density_raw.restype = C.c_double
    L.ref_get_densities.argtypes = [C.c_void_p, _dp]
    L.ref_electron_count.argtypes = [C.c_void_p]
    L.ref_electron_count.restype = C.c_double
    L.ref_hartree.argtypes = [C.c_void_p, _dp]
    L.ref_get_hartree_intermediates.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _dp]
    L.ref_spline_value.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
    L.ref_spline_value.restype = C.c_double
    L.ref_xc.argtypes = [C.c_void_p, _dp, _dp]
    L.ref_functional.argtypes = [_dp, C.c_long, _dp, _dp, _dp, _dp]
    L.ref_get_matrix.argtypes = [C.c_void_p, C.c_int, _dp]
    L.ref_get_matrix.restype = C.c_int
    L.ref_get_energies.argtypes = [C.c_void_p, _dp]
    L.ref_get_energies.restype = C.c_int
    L.ref_scf_step.argtypes = [C.c_void_p]
    L.ref_scf_step.restype = C.c_double
    L.ref_rect_density.argtypes = [C.c_void_p, C.c_double, C.c_int, _dp, _dp, _dp, _dp]
    L.ref_rect_density.restype = C.c_int
    L.ref_rect_write.argtypes = [C.c_void_p, C.c_double, C.c_int, _dp, C.c_char_p]
    L.ref_rect_write.restype = C.c_int
    L.ref_time_iteration.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _dp]
    L.ref_time_iteration.restype = C.c_double
    _libs[fast] = L
    return L


MATS = {"S": 0, "T": 1, "V": 2, "H": 3, "X": 4, "P": 5, "J": 6, "XC": 7, "C": 8}


class Ref:
    """One reference run.  full=True builds the reference's whole DFT object (integrals, core guess)."""

    def __init__(self, infile, full=False, quiet=True, fast=False, rundir=RUNDIR):
        self.L = _lib(fast)
        infile = os.path.abspath(infile)
        self.h = self.L.ref_open(infile.encode(), rundir.encode(), 0 if full else 1, 1 if quiet else 0)
        if not self.h:
            raise RuntimeError("reference failed: " + self.L.ref_last_error().decode())
        L, h = self.L, self.h
        self.full = full
        self.natoms, self.nbf, self.nprims = L.ref_natoms(h), L.ref_nbf(h), L.ref_nprims(h)
        self.npts, self.nrad, self.nang = L.ref_npoints(h), L.ref_nrad(h), L.ref_nang(h)
        self.lebedev_order, self.lmax, self.nelec = L.ref_lebedev_order(h), L.ref_lmax(h), L.ref_nelec(h)
        self.nlm = (self.lmax + 1) ** 2

    def close(self):
        if self.h:
            self.L.ref_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def system(self):
        """Flat description of atoms and basis functions in the reference's own ordering."""
        Z = np.zeros(self.natoms, np.int32)
        xyz = np.zeros((self.natoms, 3))
        self.L.ref_get_atoms(self.h, Z.ctypes.data_as(_ip), _p(xyz))
        nprim = np.zeros(self.nbf, np.int32)
        center = np.zeros((self.nbf, 3))
        alpha, coeff, norm = (np.zeros(self.nprims) for _ in range(3))
        lmn = np.zeros((self.nprims, 3), np.int32)
        self.L.ref_get_basis(self.h, nprim.ctypes.data_as(_ip), _p(center), _p(alpha), _p(coeff), _p(norm),
                             lmn.ctypes.data_as(_ip))
        return dict(Z=Z, xyz=xyz, bf_nprim=nprim, bf_center=center, alpha=alpha, coeff=coeff, norm=norm, lmn=lmn,
                    radial_points=self.nrad, lebedev_order=self.lebedev_order, lmax=self.lmax)

    def grid(self):
        xyz = np.zeros((self.npts, 3))
        w = np.zeros(self.npts)
        wb = np.zeros(self.npts)
        self.L.ref_get_grid(self.h, _p(xyz), _p(w), _p(wb))
        return xyz, w, wb

    def amplitudes(self):
        phi = np.zeros((self.npts, self.nbf))
        self.L.ref_get_amplitudes(self.h, _p(phi))
        return phi

    def set_density(self, P):
        P = np.asfortranarray(P, dtype=np.float64)
        self.L.ref_set_density(self.h, P.ctypes.data_as(_dp))

    def set_density_raw(self, P):
        P = np.asfortranarray(P, dtype=np.float64)
        return self.L.ref_set_density_raw(self.h, P.ctypes.data_as(_dp))

    def densities(self):
        rho = np.zeros(self.npts)
        self.L.ref_get_densities(self.h, _p(rho))
        return rho

    def rect_density(self, size, dp, P):
        """RectangularGrid::build_grid(size, dp) + set_density(P): positions [n][3], rho [n], gradient [n][3]."""
        P = np.asfortranarray(P, dtype=np.float64)
        n = dp ** 3
        pos, rho, grad = np.zeros((n, 3)), np.zeros(n), np.zeros((n, 3))
        if self.L.ref_rect_density(self.h, float(size), int(dp), P.ctypes.data_as(_dp), _p(pos), _p(rho), _p(grad)) != n:
            raise RuntimeError(self.L.ref_last_error().decode())
        return pos, rho, grad

    def rect_write(self, size, dp, P, filename):
        """RectangularGrid::write_gradient: the reference's own dump file for the same grid."""
        P = np.asfortranarray(P, dtype=np.float64)
        if self.L.ref_rect_write(self.h, float(size), int(dp), P.ctypes.data_as(_dp), filename.encode()) != 0:
            raise RuntimeError(self.L.ref_last_error().decode())

    def electron_count(self):
        return self.L.ref_electron_count(self.h)

    def hartree(self):
        J = np.zeros((self.nbf, self.nbf), order="F")
        self.L.ref_hartree(self.h, J.ctypes.data_as(_dp))
        return np.array(J)

    def hartree_intermediates(self):
        rho_lm = np.zeros((self.natoms, self.nrad, self.nlm))
        U_lm = np.zeros((self.natoms, self.nrad, self.nlm))
        V = np.zeros(self.npts)
        Vf = np.zeros(self.npts)
        q = np.zeros(self.natoms)
        self.L.ref_get_hartree_intermediates(self.h, _p(rho_lm), _p(U_lm), _p(V), _p(Vf), _p(q))
        return dict(rho_lm=rho_lm, U_lm=U_lm, V=V, V_fuzzy=Vf, q=q)

    def spline_value(self, atom, lm, r):
        return self.L.ref_spline_value(self.h, atom, lm, r)

    def xc(self):
        XC = np.zeros((self.nbf, self.nbf), order="F")
        exc = C.c_double(0.0)
        self.L.ref_xc(self.h, XC.ctypes.data_as(_dp), C.cast(C.byref(exc), _dp))
        return np.array(XC), exc.value

    def matrix(self, name):
        M = np.zeros((self.nbf, self.nbf), order="F")
        if self.L.ref_get_matrix(self.h, MATS[name], M.ctypes.data_as(_dp)) != 0:
            raise RuntimeError("matrix %s needs full=True" % name)
        return np.array(M)

    def energies(self):
        e = np.zeros(6)
        if self.L.ref_get_energies(self.h, _p(e)) != 0:
            raise RuntimeError("energies need full=True")
        return dict(et=e[0], exc=e[1], enuc=e[2], e_one=e[3], e_j=e[4], nel=e[5])

    def scf_step(self):
        return self.L.ref_scf_step(self.h)

    def time_iteration(self, P):
        P = np.asfortranarray(P, dtype=np.float64)
        ph = np.zeros(4)
        J = np.zeros((self.nbf, self.nbf), order="F")
        XC = np.zeros((self.nbf, self.nbf), order="F")
        exc = C.c_double(0.0)
        t = self.L.ref_time_iteration(self.h, P.ctypes.data_as(_dp), _p(ph), J.ctypes.data_as(_dp),
                                      XC.ctypes.data_as(_dp), C.cast(C.byref(exc), _dp))
        return t, ph, np.array(J), np.array(XC), exc.value


def functional(rho):
    rho = np.ascontiguousarray(rho, dtype=np.float64)
    out = [np.zeros_like(rho) for _ in range(4)]
    _lib().ref_functional(_p(rho), rho.size, *[_p(o) for o in out])
    return out  # ex, vx(alpha), ec, vc(alpha)
