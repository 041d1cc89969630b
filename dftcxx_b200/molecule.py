"""Host-side input layer for tests and benchmarks: dftcxx `.in` files + basis-set files -> the flat
`dftgrid_system` the C ABI takes.  Mirrors the reference's Settings / Molecule classes
(reference src/settings.cpp:39-187, src/molecule.cpp:61-288, src/cgf.cpp:102-114,185-232); the
product's own host is the C++ one under dftcxx_b200/host — this module is the Python view of the
same rules so the parity tests can drive the C ABI directly.
"""
import math
import os

import numpy as np

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
ELEMENTS = ["H", "He", "Li", "Be", "B", "C", "N", "O", "F", "Ne", "Na", "Mg", "Al", "Si", "P", "S", "Cl", "Ar"]
ANGSTROM_TO_BOHR = 1.889725989  # src/molecule.cpp:65
LEBEDEV_COUNTS = [6, 14, 26, 38, 50, 74, 86, 110, 146, 170, 194]
GRID_PRESETS = {  # src/settings.cpp:158-187: radial points, lebedev order index, lmax
    "coarse": (10, 4, 5),
    "medium": (15, 7, 8),
    "fine": (20, 8, 10),
    "ultrafine": (30, 10, 11),
}
# Cartesian powers per shell type, in the reference's order (src/cgf.cpp:190-225, src/molecule.cpp:203-219)
SHELL_LMN = {
    "S": [(0, 0, 0)],
    "P": [(1, 0, 0), (0, 1, 0), (0, 0, 1)],
    "D": [(2, 0, 0), (1, 1, 0), (1, 0, 1), (0, 2, 0), (0, 1, 1), (0, 0, 2)],
}


class Settings:
    """key = value lines before `system:` (src/settings.cpp:39-70) and the grid defaults (:113-150)."""

    def __init__(self, text):
        self.key_values = {}
        for line in text.splitlines():
            pieces = [p for p in _split_compress(line, "=")]
            if len(pieces) == 2:
                self.key_values.setdefault(pieces[0].strip(), pieces[1].strip())
            if line.startswith("system:") and line.strip() == "system:":
                break
        grid = self.key_values.get("grid", "medium")
        self.radial_points, self.lebedev_order, self.lmax = GRID_PRESETS.get(grid, GRID_PRESETS["medium"])
        for key in ("radial_points", "lebedev_order", "lmax"):
            try:
                setattr(self, key, int(_strict_uint(self.key_values[key])))
            except (KeyError, ValueError):
                pass

    def get_value(self, key):
        if key not in self.key_values:
            raise KeyError("Could not find " + key)
        return self.key_values[key]


def _split_compress(line, seps):
    """boost::split(..., is_any_of(seps), token_compress_on): adjacent separators merge, a leading one yields ''."""
    out, cur, i = [], "", 0
    while i < len(line):
        if line[i] in seps:
            out.append(cur)
            cur = ""
            i += 1
            while i < len(line) and line[i] in seps:
                i += 1
        else:
            cur += line[i]
            i += 1
    out.append(cur)
    return out


def _strict_uint(s):
    if not s.isdigit():
        raise ValueError(s)
    return int(s)


def double_factorial(n):
    r = 1.0
    while n > 1:
        r *= n
        n -= 2
    return r


def gto_norm(alpha, l, m, n):
    """GTO::calculate_normalization_constant (src/cgf.cpp:102-114), truncated pi included."""
    pi = 3.14159265359
    nom = math.pow(2.0, 2.0 * (l + m + n) + 3.0 / 2.0) * math.pow(alpha, (l + m + n) + 3.0 / 2.0)
    denom = ((1 if l < 1 else double_factorial(2 * l - 1)) * (1 if m < 1 else double_factorial(2 * m - 1)) *
             (1 if n < 1 else double_factorial(2 * n - 1)) * math.pow(pi, 3.0 / 2.0))
    return math.sqrt(nom / denom)


def read_basis(path, highest_z):
    """Per element: list of shells (type, [(exponent, coefficient)...]); src/molecule.cpp:154-237."""
    shells = {}
    with open(path) as f:
        lines = f.read().splitlines()
    i = 0
    while i < len(lines):
        line = lines[i]
        i += 1
        if not line or line[0] == "#":
            continue
        pieces = _split_compress(line, " \t")
        z, nshell = _strict_uint(pieces[0]), _strict_uint(pieces[1])
        cur = []
        for _ in range(nshell):
            pieces = _split_compress(lines[i], " \t")
            i += 1
            typ, nprim = pieces[0][0], _strict_uint(pieces[1])
            prims = []
            for _ in range(nprim):
                pieces = _split_compress(lines[i], " \t")
                i += 1
                prims.append((float(pieces[1]), float(pieces[2])))
            cur.append((typ, prims))
        shells[z] = cur
        if z == highest_z:
            break
    return shells


class Molecule:
    """Atoms + contracted Gaussian basis in the reference's ordering: basis functions are appended element by
    element in basis-file order, and for each element atom by atom (src/molecule.cpp:222-235)."""

    def __init__(self, Z, xyz_bohr, basis="p631", basis_dir=None, settings=None, name="molecule"):
        self.Z = np.asarray(Z, dtype=np.int32)
        self.xyz = np.asarray(xyz_bohr, dtype=np.float64).reshape(-1, 3)
        self.name = name
        self.basis = basis
        self.settings = settings
        basis_dir = basis_dir or os.path.join(DATA, "basis")
        path = os.path.join(basis_dir, basis + ".dat")
        if not os.path.exists(path):
            raise RuntimeError("Cannot open " + path + "!")
        table = read_basis(path, int(self.Z.max()))
        bf_nprim, bf_center, alpha, coeff, norm, lmn, bf_atom = [], [], [], [], [], [], []
        for z in sorted(table):  # basis files list elements in ascending Z
            for a in np.nonzero(self.Z == z)[0]:
                for typ, prims in table[z]:
                    if typ not in SHELL_LMN:
                        continue  # the reference silently allocates nothing for unknown shell letters
                    for (l, m, n) in SHELL_LMN[typ]:
                        bf_nprim.append(len(prims))
                        bf_center.append(self.xyz[a])
                        bf_atom.append(a)
                        for e, c in prims:
                            alpha.append(e)
                            coeff.append(c)
                            norm.append(gto_norm(e, l, m, n))
                            lmn.append((l, m, n))
        missing = set(int(z) for z in self.Z) - set(table)
        self.bf_nprim = np.asarray(bf_nprim, dtype=np.int32)
        self.bf_center = np.asarray(bf_center, dtype=np.float64).reshape(-1, 3)
        self.bf_atom = np.asarray(bf_atom, dtype=np.int32)
        self.alpha = np.asarray(alpha, dtype=np.float64)
        self.coeff = np.asarray(coeff, dtype=np.float64)
        self.norm = np.asarray(norm, dtype=np.float64)
        self.lmn = np.asarray(lmn, dtype=np.int32).reshape(-1, 3)
        self.missing_elements = missing

    @property
    def natoms(self):
        return len(self.Z)

    @property
    def nbf(self):
        return len(self.bf_nprim)

    @property
    def nelec(self):
        return int(self.Z.sum())

    @classmethod
    def from_file(cls, path, basis_dir=None):
        """Molecule::read_molecule_from_file (src/molecule.cpp:61-128)."""
        if not os.path.exists(path):
            raise RuntimeError("Cannot open " + path + "!")
        text = open(path).read()
        st = Settings(text)
        lines = text.splitlines()
        k = 0
        while k < len(lines) and lines[k][:7] != "system:":
            k += 1
        natoms = _strict_uint(lines[k + 1])
        ang = st.key_values.get("units", "bohr") == "angstrom"
        Z, xyz = [], []
        for line in lines[k + 2:k + 2 + natoms]:
            pieces = _split_compress(line, " \t")
            if pieces[0] not in ELEMENTS:
                raise RuntimeError("Unknown element: " + pieces[0])
            Z.append(ELEMENTS.index(pieces[0]) + 1)
            c = [float(pieces[1]), float(pieces[2]), float(pieces[3])]
            if ang:
                c = [v * ANGSTROM_TO_BOHR for v in c]
            xyz.append(c)
        return cls(Z, xyz, basis=st.get_value("basis"), basis_dir=basis_dir, settings=st, name=st.key_values.get("name", ""))

    def to_input(self, grid="fine", extra=()):
        """Write this molecule back as a dftcxx `.in` text (coordinates in bohr, 17 significant digits)."""
        out = ["name = " + self.name, "basis = " + self.basis, "units = bohr"]
        if grid:
            out.append("grid = " + grid)
        out += list(extra)
        out += ["", "system:", str(self.natoms)]
        for z, c in zip(self.Z, self.xyz):
            out.append("%s %.17g %.17g %.17g" % (ELEMENTS[z - 1], c[0], c[1], c[2]))
        return "\n".join(out) + "\n"
