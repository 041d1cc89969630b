"""TEST INFRASTRUCTURE ONLY: ctypes view of oracle/liboracle.so (the plain-C restatement in oracle_port.c)."""
import ctypes as C
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
LEBEDEV_COUNTS = [6, 14, 26, 38, 50, 74, 86, 110, 146, 170, 194]


def available():
    return os.path.exists(os.path.join(HERE, "liboracle.so"))


_L = None


def _lib():
    global _L
    if _L is None:
        L = C.CDLL(os.path.join(HERE, "liboracle.so"))
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.c_int, _ip, _dp, C.c_int, _ip, _dp, C.c_int, _dp, _dp, _dp, _ip, C.c_int, C.c_int, _dp, C.c_int]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_npoints.argtypes = [C.c_void_p]
        L.oracle_npoints.restype = C.c_long
        L.oracle_get_grid.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.oracle_get_phi.argtypes = [C.c_void_p, _dp]
        L.oracle_get_rho.argtypes = [C.c_void_p, _dp]
        L.oracle_get_hartree.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
        L.oracle_set_density.argtypes = [C.c_void_p, _dp, C.c_int]
        L.oracle_set_density.restype = C.c_double
        L.oracle_electron_count.argtypes = [C.c_void_p]
        L.oracle_electron_count.restype = C.c_double
        L.oracle_xc.argtypes = [C.c_void_p, _dp]
        L.oracle_xc.restype = C.c_double
        L.oracle_hartree.argtypes = [C.c_void_p, _dp]
        L.oracle_rect_density.argtypes = [C.c_void_p, C.c_double, C.c_int, _dp, _dp, _dp, _dp]
        _L = L
    return _L


def lebedev(order):
    tab = json.load(open(os.path.join(ROOT, "dftcxx_b200", "data", "lebedev.json")))
    off = sum(LEBEDEV_COUNTS[:order])
    return np.ascontiguousarray(tab["xyzw"][off:off + LEBEDEV_COUNTS[order]], dtype=np.float64)


class Port:
    def __init__(self, s, radial_points, lebedev_order, lmax):
        a = lambda k, t: np.ascontiguousarray(s[k], dtype=t)  # noqa: E731
        self.k = dict(Z=a("Z", np.int32), xyz=a("xyz", np.float64), bf_nprim=a("bf_nprim", np.int32), bf_center=a("bf_center", np.float64),
                      alpha=a("alpha", np.float64), coeff=a("coeff", np.float64), norm=a("norm", np.float64), lmn=a("lmn", np.int32))
        k = self.k
        self.leb = lebedev(lebedev_order)
        self.natoms, self.nbf, self.nrad, self.nang, self.nlm = len(k["Z"]), len(k["bf_nprim"]), radial_points, len(self.leb), (lmax + 1) ** 2
        p = lambda x, t=_dp: x.ctypes.data_as(t)  # noqa: E731
        self.h = _lib().oracle_create(self.natoms, p(k["Z"], _ip), p(k["xyz"]), self.nbf, p(k["bf_nprim"], _ip), p(k["bf_center"]),
                                      len(k["alpha"]), p(k["alpha"]), p(k["coeff"]), p(k["norm"]), p(k["lmn"], _ip), radial_points,
                                      self.nang, p(self.leb), lmax)
        self.npts = _lib().oracle_npoints(self.h)

    def grid(self):
        xyz, w, wb = np.zeros((self.npts, 3)), np.zeros(self.npts), np.zeros(self.npts)
        _lib().oracle_get_grid(self.h, xyz.ctypes.data_as(_dp), w.ctypes.data_as(_dp), wb.ctypes.data_as(_dp))
        return xyz, w, wb

    def amplitudes(self):
        phi = np.zeros((self.npts, self.nbf))
        _lib().oracle_get_phi(self.h, phi.ctypes.data_as(_dp))
        return phi

    def set_density(self, P, correct=True):
        P = np.asfortranarray(P, dtype=np.float64)
        return _lib().oracle_set_density(self.h, P.ctypes.data_as(_dp), 1 if correct else 0)

    def densities(self):
        rho = np.zeros(self.npts)
        _lib().oracle_get_rho(self.h, rho.ctypes.data_as(_dp))
        return rho

    def electron_count(self):
        return _lib().oracle_electron_count(self.h)

    def xc(self):
        XC = np.zeros((self.nbf, self.nbf), order="F")
        exc = _lib().oracle_xc(self.h, XC.ctypes.data_as(_dp))
        return np.array(XC), exc

    def hartree(self):
        J = np.zeros((self.nbf, self.nbf))
        _lib().oracle_hartree(self.h, J.ctypes.data_as(_dp))
        n = (self.natoms, self.nrad, self.nlm)
        rho_lm, U_lm, V, Vf = np.zeros(n), np.zeros(n), np.zeros(self.npts), np.zeros(self.npts)
        _lib().oracle_get_hartree(self.h, rho_lm.ctypes.data_as(_dp), U_lm.ctypes.data_as(_dp), V.ctypes.data_as(_dp), Vf.ctypes.data_as(_dp))
        return J, dict(rho_lm=rho_lm, U_lm=U_lm, V=V, V_fuzzy=Vf)

    def rect_density(self, size, dp, P):
        """RectangularGrid::build_grid(size, dp) + set_density(P) (src/rectangulargrid.cpp:34-80)."""
        P = np.asfortranarray(P, dtype=np.float64)
        n = int(dp) ** 3
        pos, rho, grad = np.zeros((n, 3)), np.zeros(n), np.zeros((n, 3))
        _lib().oracle_rect_density(self.h, float(size), int(dp), P.ctypes.data_as(_dp), pos.ctypes.data_as(_dp), rho.ctypes.data_as(_dp),
                                   grad.ctypes.data_as(_dp))
        return pos, rho, grad

    def close(self):
        if self.h:
            _lib().oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
