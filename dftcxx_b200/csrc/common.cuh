// Shared device-side declarations for the grid engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace dfg {

constexpr int kNbAlign = 32;  // Phi row length (doubles) is padded to a multiple of this; pad columns are zero

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Flattened basis for the amplitude kernel: per atom, the distinct exponents; per CGF, its primitive terms
// pointing at one of its atom's exponents.
struct BasisDev {
    int natoms;
    int nbf;
    int nbp;                  // padded nbf
    const double* atom_xyz;   // [natoms][3]
    const int* atom_exp_off;  // [natoms+1] into exp_alpha
    const double* exp_alpha;  // distinct exponents per atom
    const int* atom_bf_off;   // [natoms+1] into bf_* (CGFs grouped by atom, original index kept in bf_index)
    const int* bf_index;      // [nbf] output column of the grouped CGF
    const int* bf_prim_off;   // [nbf+1] into prim_*
    const int* prim_exp;      // [nprim] index into the atom's exponent list (relative to atom_exp_off[atom])
    const double* prim_coeff; // [nprim]
    const double* prim_norm;  // [nprim]
    const int* prim_lmn;      // [nprim] packed l | m<<4 | n<<8
    int max_exp_per_atom;
};

struct GridShape {
    int natoms, nrad, nang, lmax, nlm;
    long npts;       // whole molecule
    long nloc;       // this rank
    long shell0;     // first global shell (atom*nrad + i) of this rank
    long nshell_loc; // number of local shells
};

}  // namespace dfg
