/* TEST INFRASTRUCTURE ONLY (oracle/).  Never linked into, imported by, or called from the product.
 *
 * Plain-C restatement of the reference's numerical-grid hot path (ifilot/dftcxx), function by function, for use as the
 * portable parity oracle in tests/ (and, optionally, as a CPU timing port).  It deliberately keeps the reference's
 * scalar algorithm — per-point loops, pow()-based Becke cell function, trigonometric Y_lm through acos/atan2, one
 * dense pivoted solve per l, linear spline search — and cites the reference lines each routine follows.  Pinned
 * against the golden vectors produced by the unmodified reference (tests/golden/, tests/test_cpu.py).
 *
 * Build: make -C oracle port   (gcc -O2 -ffp-contract=off -fopenmp)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int natoms, nbf, nprim, nrad, nang, lmax, nlm;
    long npts;
    int *Z, *bf_nprim, *bf_off, *lmn;
    double *axyz, *bf_center, *alpha, *coeff, *norm, *leb; /* leb: [nang][4] */
    double *xyz, *rat, *w, *wb, *rho, *phi, *V, *Vfuzzy;   /* per point; rat = position relative to own atom */
    double *r_n;                                            /* [nrad] */
    double *rho_lm, *U_lm, *q;                              /* [natoms][nrad][nlm], [natoms] */
    double *spl;                                            /* [natoms][nlm][nrad-1][4] */
} oracle_t;

static const double PI = 3.14159265358979323846; /* src/atomicgrid.h:33 */

static double *dalloc(size_t n) { return (double *)calloc(n ? n : 1, sizeof(double)); }

/* ---- src/spherical_harmonics.cpp ------------------------------------------------------------------------ */
static double factorial(int n) {
    double f = 1.0;
    for (int i = 2; i <= n; i++) f *= (double)i;
    return f;
}
static double sh_prefactor(int l, int m) { /* :28-33 */
    const double pre = 1.0 / sqrt(4 * M_PI);
    const int am = abs(m);
    return pre * (m == 0 ? 1 : sqrt(2.0)) * sqrt((double)(2 * l + 1) * factorial(l - am) / factorial(l + am));
}
static double legendre_p(int n, int m, double x) { /* :81-117 */
    double v[64];
    for (int i = 0; i <= n; i++) v[i] = 0.0;
    if (m <= n) {
        v[m] = 1.0;
        double fact = 1.0;
        for (int k = 0; k < m; k++) {
            v[m] *= -fact * sqrt(1.0 - x * x);
            fact += 2.0;
        }
    }
    if (m + 1 <= n) v[m + 1] = x * (double)(2 * m + 1) * v[m];
    for (int j = m + 2; j <= n; j++) v[j] = ((double)(2 * j - 1) * x * v[j - 1] + (double)(-j - m + 1) * v[j - 2]) / (double)(j - m);
    return v[n];
}
static double spherical_harmonic(int l, int m, double pole, double azimuth) { /* :24-47 */
    const double polar = legendre_p(l, abs(m), cos(pole));
    if (m == 0) return polar;
    return polar * (m > 0 ? cos((double)m * azimuth) : sin(-(double)m * azimuth));
}

/* ---- src/cgf.cpp:49-57,146-154 ------------------------------------------------------------------------------ */
static double cgf_amp(const oracle_t *o, int b, const double *r) {
    double sum = 0.0;
    const double *c = o->bf_center + 3 * b;
    const double dx = r[0] - c[0], dy = r[1] - c[1], dz = r[2] - c[2];
    const double r2 = dx * dx + dy * dy + dz * dz;
    for (int k = o->bf_off[b]; k < o->bf_off[b + 1]; k++) {
        const double amp = o->norm[k] * pow(dx, o->lmn[3 * k]) * pow(dy, o->lmn[3 * k + 1]) * pow(dz, o->lmn[3 * k + 2]) * exp(-o->alpha[k] * r2);
        sum += o->coeff[k] * amp;
    }
    return sum;
}

/* ---- src/moleculargrid.cpp:275-329 -------------------------------------------------------------------------- */
static double norm3(const double *a, const double *b) {
    const double x = a[0] - b[0], y = a[1] - b[1], z = a[2] - b[2];
    return sqrt(x * x + y * y + z * z);
}
static double becke_cutoff(double mu) {
    for (int i = 0; i < 3; i++) mu = 3.0 / 2.0 * mu - 0.5 * pow(mu, 3.0);
    return 0.5 * (1.0 - mu);
}
static double becke_pn(const oracle_t *o, int i, const double *p0) {
    double wprod = 1.0;
    const double *p1 = o->axyz + 3 * i;
    for (int j = 0; j < o->natoms; j++) {
        if (i == j) continue;
        const double *p2 = o->axyz + 3 * j;
        const double mu = (norm3(p0, p1) - norm3(p0, p2)) / norm3(p2, p1);
        wprod *= becke_cutoff(mu);
    }
    return wprod;
}

/* ---- grid construction: src/atomicgrid.cpp:32-88, src/moleculargrid.cpp:193-261 ------------------------------- */
oracle_t *oracle_create(int natoms, const int *Z, const double *axyz, int nbf, const int *bf_nprim, const double *bf_center, int nprim,
                        const double *alpha, const double *coeff, const double *norm, const int *lmn, int nrad, int nang,
                        const double *leb, int lmax) {
    oracle_t *o = (oracle_t *)calloc(1, sizeof(oracle_t));
    o->natoms = natoms, o->nbf = nbf, o->nprim = nprim, o->nrad = nrad, o->nang = nang, o->lmax = lmax;
    o->nlm = (lmax + 1) * (lmax + 1);
    o->npts = (long)natoms * nrad * nang;
#define DUP(dst, src, n, T)                      \
    dst = (T *)malloc(sizeof(T) * (size_t)(n)); \
    memcpy(dst, src, sizeof(T) * (size_t)(n))
    DUP(o->Z, Z, natoms, int);
    DUP(o->axyz, axyz, 3 * natoms, double);
    DUP(o->bf_nprim, bf_nprim, nbf, int);
    DUP(o->bf_center, bf_center, 3 * nbf, double);
    DUP(o->alpha, alpha, nprim, double);
    DUP(o->coeff, coeff, nprim, double);
    DUP(o->norm, norm, nprim, double);
    DUP(o->lmn, lmn, 3 * nprim, int);
    DUP(o->leb, leb, 4 * nang, double);
    o->bf_off = (int *)calloc(nbf + 1, sizeof(int));
    for (int b = 0; b < nbf; b++) o->bf_off[b + 1] = o->bf_off[b] + bf_nprim[b];
    const long np = o->npts;
    o->xyz = dalloc(3 * np), o->rat = dalloc(3 * np), o->w = dalloc(np), o->wb = dalloc(np), o->rho = dalloc(np);
    o->V = dalloc(np), o->Vfuzzy = dalloc(np), o->phi = dalloc((size_t)np * nbf);
    o->r_n = dalloc(nrad);
    o->rho_lm = dalloc((size_t)natoms * nrad * o->nlm), o->U_lm = dalloc((size_t)natoms * nrad * o->nlm), o->q = dalloc(natoms);
    o->spl = dalloc((size_t)natoms * o->nlm * (nrad - 1) * 4);

    const double f = PI / (double)(nrad + 1);
    for (int at = 0; at < natoms; at++) {
        const double *p1 = axyz + 3 * at;
        for (int p = 1; p <= nrad; p++) {
            double w = f * pow(sin(f * (double)p), 2.0);
            const double x = cos(f * (double)p);
            const double r = (1.0 + x) / (1.0 - x);
            o->r_n[p - 1] = r;
            w = w / sqrt(1.0 - pow(x, 2.0)) * 2.0 / pow(1.0 - x, 2.0);
            for (int a = 0; a < nang; a++) {
                const long idx = ((long)at * nrad + (p - 1)) * nang + a;
                for (int d = 0; d < 3; d++) {
                    o->rat[3 * idx + d] = leb[4 * a + d] * r;
                    o->xyz[3 * idx + d] = p1[d] + o->rat[3 * idx + d];
                }
                o->w[idx] = w * leb[4 * a + 3];
            }
        }
    }
#pragma omp parallel for schedule(static)
    for (long i = 0; i < np; i++) {
        for (int b = 0; b < nbf; b++) o->phi[(size_t)i * nbf + b] = cgf_amp(o, b, o->xyz + 3 * i);
        const double *q = o->rat + 3 * i;
        o->w[i] *= (q[0] * q[0] + q[1] * q[1] + q[2] * q[2]) * 4.0 * PI; /* :85-87 */
    }
#pragma omp parallel for schedule(dynamic, 64)
    for (long i = 0; i < np; i++) { /* Becke weights, src/moleculargrid.cpp:229-254 */
        const int own = (int)(i / ((long)nrad * nang));
        double denom = 0.0, nom = 1.0;
        for (int k = 0; k < natoms; k++) {
            const double term = becke_pn(o, k, o->xyz + 3 * i);
            denom += term;
            if (own == k) nom = term;
        }
        o->wb[i] = nom / denom;
        o->w[i] *= o->wb[i]; /* src/atomicgrid.cpp:196-206 */
    }
    return o;
}

void oracle_destroy(oracle_t *o) {
    if (!o) return;
    void *ptrs[] = {o->Z, o->bf_nprim, o->bf_off, o->lmn, o->axyz, o->bf_center, o->alpha, o->coeff, o->norm, o->leb, o->xyz, o->rat,
                    o->w, o->wb, o->rho, o->phi, o->V, o->Vfuzzy, o->r_n, o->rho_lm, o->U_lm, o->q, o->spl};
    for (size_t i = 0; i < sizeof ptrs / sizeof *ptrs; i++) free(ptrs[i]);
    free(o);
}

long oracle_npoints(const oracle_t *o) { return o->npts; }
void oracle_get_grid(const oracle_t *o, double *xyz, double *w, double *wb) {
    if (xyz) memcpy(xyz, o->xyz, sizeof(double) * 3 * o->npts);
    if (w) memcpy(w, o->w, sizeof(double) * o->npts);
    if (wb) memcpy(wb, o->wb, sizeof(double) * o->npts);
}
void oracle_get_phi(const oracle_t *o, double *phi) { memcpy(phi, o->phi, sizeof(double) * (size_t)o->npts * o->nbf); }
void oracle_get_rho(const oracle_t *o, double *rho) { memcpy(rho, o->rho, sizeof(double) * o->npts); }
void oracle_get_hartree(const oracle_t *o, double *rho_lm, double *U_lm, double *V, double *Vf) {
    const size_t n = (size_t)o->natoms * o->nrad * o->nlm;
    if (rho_lm) memcpy(rho_lm, o->rho_lm, sizeof(double) * n);
    if (U_lm) memcpy(U_lm, o->U_lm, sizeof(double) * n);
    if (V) memcpy(V, o->V, sizeof(double) * o->npts);
    if (Vf) memcpy(Vf, o->Vfuzzy, sizeof(double) * o->npts);
}

/* ---- density: src/gridpoint.cpp:82-84, src/moleculargrid.cpp:132-166 ---------------------------------------------- */
static double atom_charge(const oracle_t *o, int at) { /* src/atomicgrid.cpp:520-530 */
    const long n = (long)o->nrad * o->nang;
    double d = 0.0;
    for (long i = at * n; i < (at + 1) * n; i++) d += o->w[i] * o->rho[i];
    return d;
}
double oracle_set_density(oracle_t *o, const double *P, int correct) {
    const int nb = o->nbf;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < o->npts; i++) {
        const double *ph = o->phi + (size_t)i * nb;
        double acc = 0.0;
        for (int a = 0; a < nb; a++) { /* phi . (P phi) */
            double t = 0.0;
            for (int b = 0; b < nb; b++) t += P[(size_t)b * nb + a] * ph[b];
            acc += ph[a] * t;
        }
        o->rho[i] = 2.0 * acc;
    }
    double sum = 0.0, charge = 0.0;
    for (int at = 0; at < o->natoms; at++) {
        sum += atom_charge(o, at);
        charge += (double)o->Z[at];
    }
    if (correct) {
        const double s = charge / sum;
        for (long i = 0; i < o->npts; i++) o->rho[i] *= s;
    }
    return sum; /* electron count before the rescale */
}
double oracle_electron_count(const oracle_t *o) {
    double s = 0.0;
    for (int at = 0; at < o->natoms; at++) s += atom_charge(o, at);
    return s;
}

/* ---- functionals: src/functionals.cpp:24-158 (closed shell) ---------------------------------------------------------- */
static double vwn_xx(double x, double b, double c) { return x * x + b * x + c; }
static double vwn_eps(double x, double a, double x0, double b, double c) {
    const double q = sqrt(4.0 * c - b * b);
    return a * (log(x * x / vwn_xx(x, b, c)) - b * (x0 / vwn_xx(x0, b, c)) * log(pow(x - x0, 2.0) / vwn_xx(x, b, c)) +
                (2.0 * b / q) * (1.0 - (x0 * (2.0 * x0 + b) / vwn_xx(x0, b, c))) * atan(q / (2.0 * x + b)));
}
static double vwn_deps(double x, double a, double x0, double b, double c) {
    const double q = sqrt(4.0 * c - b * b);
    return a * (2.0 / x - (2.0 * x + b) / vwn_xx(x, b, c) - 4.0 * b / (pow(2.0 * x + b, 2.0) + q * q) -
                (b * x0 / vwn_xx(x0, b, c)) * (2.0 / (x - x0) - (2.0 * x + b) / vwn_xx(x, b, c) - 4.0 * (2.0 * x0 + b) / (pow(2.0 * x + b, 2.0) + q * q)));
}
void oracle_functional(const double *rho, long n, double *exc_dens, double *vxc) {
    const double tol = 1e-10, xalpha = 2.0 / 3.0;
    const double fac = -2.25 * xalpha * pow(3.0 / 4.0 / PI, 1.0 / 3.0);
    for (long i = 0; i < n; i++) {
        const double da = rho[i] * 0.5, db = rho[i] * 0.5;
        double ex = 0.0, vxa = 0.0, vxb = 0.0, ec = 0.0, vca = 0.0, vcb = 0.0;
        if (!(da < tol)) {
            const double rho3 = pow(da, 1.0 / 3.0);
            ex += fac * da * rho3;
            vxa += 4.0 / 3.0 * fac * rho3;
        }
        if (!(db < tol)) {
            const double rho3 = pow(db, 1.0 / 3.0);
            ex += fac * db * rho3;
            vxb += 4.0 / 3.0 * fac * rho3;
        }
        const double dens = da + db;
        if (!(dens < tol)) { /* zeta = 0 => g = 0 < tol: paramagnetic branch (:88-99) */
            const double x = pow(3.0 / 4.0 / PI / dens, 1.0 / 6.0);
            const double epsp = vwn_eps(x, 0.0310907, -0.10498, 3.72744, 12.9352);
            const double depsp = vwn_deps(x, 0.0310907, -0.10498, 3.72744, 12.9352);
            ec = epsp * dens;
            vca = vcb = epsp - (x / 6.0) * depsp;
        }
        exc_dens[i] = ex + ec;
        vxc[i] = (vxa + vxb + vca + vcb) * 0.5; /* src/dft.cpp:424 */
    }
}

/* XC matrix and energy: src/dft.cpp:394-433 */
double oracle_xc(const oracle_t *o, double *XC) {
    const int nb = o->nbf;
    const long np = o->npts;
    double *e = dalloc(np), *v = dalloc(np);
    oracle_functional(o->rho, np, e, v);
    double exc = 0.0;
    for (long i = 0; i < np; i++) exc += o->w[i] * e[i];
#pragma omp parallel for schedule(dynamic)
    for (int i = 0; i < nb; i++)
        for (int j = 0; j < nb; j++) {
            double s = 0.0;
            for (long p = 0; p < np; p++) s += (o->w[p] * v[p] * o->phi[(size_t)p * nb + i]) * o->phi[(size_t)p * nb + j];
            XC[(size_t)j * nb + i] = s;
        }
    free(e), free(v);
    return exc;
}

/* ---- Hartree: src/atomicgrid.cpp:241-556, src/cspline.cpp:66-172, src/moleculargrid.cpp:336-389 ------------------------ */
static double d2zdr2(double r, double m) {
    const double nom = m * m * (m + 3.0 * r);
    const double denom = 2.0 * M_PI * pow((m * r) / ((m + r) * (m + r)), 1.5) * pow(m + r, 5.0);
    return nom / denom;
}
static double dzdrsq(double r, double m) { return m / (M_PI * M_PI * r * (m + r) * (m + r)); }

static void lu_solve_dense(int n, double *M, double *g) { /* partial-pivot LU + solve, what Eigen::PartialPivLU::solve does */
    for (int k = 0; k < n; k++) {
        int p = k;
        double best = fabs(M[(size_t)k * n + k]);
        for (int i = k + 1; i < n; i++)
            if (fabs(M[(size_t)i * n + k]) > best) best = fabs(M[(size_t)i * n + k]), p = i;
        if (p != k) {
            for (int j = 0; j < n; j++) {
                const double t = M[(size_t)k * n + j];
                M[(size_t)k * n + j] = M[(size_t)p * n + j];
                M[(size_t)p * n + j] = t;
            }
            const double t = g[k];
            g[k] = g[p];
            g[p] = t;
        }
        for (int i = k + 1; i < n; i++) {
            const double l = M[(size_t)i * n + k] / M[(size_t)k * n + k];
            if (l == 0.0) continue;
            M[(size_t)i * n + k] = l;
            for (int j = k + 1; j < n; j++) M[(size_t)i * n + j] -= l * M[(size_t)k * n + j];
            g[i] -= l * g[k];
        }
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = g[i];
        for (int j = i + 1; j < n; j++) s -= M[(size_t)i * n + j] * g[j];
        g[i] = s / M[(size_t)i * n + i];
    }
}

static void fd_matrix(const oracle_t *o, double *A) { /* src/atomicgrid.cpp:316-389 */
    const int N = o->nrad, n = N + 2;
    const double h = 1.0 / (double)(N + 1);
    memset(A, 0, sizeof(double) * (size_t)n * n);
#define AA(i, j) A[(size_t)(i)*n + (j)]
    for (int i = 0; i < n; i++) {
        double c1 = 0.0, c2 = 0.0;
        if (i > 0 && i < N + 1) c1 = dzdrsq(o->r_n[i - 1], 1.0), c2 = d2zdr2(o->r_n[i - 1], 1.0);
        if (i == 0) {
            AA(0, 0) = 1.0;
        } else if (i == 1) {
            c1 /= 12.0 * h * h, c2 /= 12.0 * h;
            AA(i, 0) = 11.0 * c1 - 3.0 * c2, AA(i, 1) = -20.0 * c1 - 10.0 * c2, AA(i, 2) = 6.0 * c1 + 18.0 * c2;
            AA(i, 3) = 4.0 * c1 - 6.0 * c2, AA(i, 4) = -1.0 * c1 + 1.0 * c2;
        } else if (i == 2) {
            c1 /= 12.0 * h * h, c2 /= 60.0 * h;
            AA(i, 0) = -1.0 * c1 + 3.0 * c2, AA(i, 1) = 16.0 * c1 - 30.0 * c2, AA(i, 2) = -30.0 * c1 - 20.0 * c2;
            AA(i, 3) = 16.0 * c1 + 60.0 * c2, AA(i, 4) = -1.0 * c1 - 15.0 * c2, AA(i, 5) = 0.0 * c1 + 2.0 * c2;
        } else if (i == N - 1) {
            c1 /= 12.0 * h * h, c2 /= 60.0 * h;
            AA(i, N - 4) = 0.0 * c1 - 2.0 * c2, AA(i, N - 3) = -1.0 * c1 + 15.0 * c2, AA(i, N - 2) = 16.0 * c1 - 60.0 * c2;
            AA(i, N - 1) = -30.0 * c1 + 20.0 * c2, AA(i, N) = 16.0 * c1 + 30.0 * c2, AA(i, N + 1) = -1.0 * c1 - 3.0 * c2;
        } else if (i == N) {
            c1 /= 12.0 * h * h, c2 /= 12.0 * h;
            AA(i, N - 3) = -1.0 * c1 - 1.0 * c2, AA(i, N - 2) = 4.0 * c1 + 6.0 * c2, AA(i, N - 1) = 6.0 * c1 - 18.0 * c2;
            AA(i, N) = -20.0 * c1 + 10.0 * c2, AA(i, N + 1) = 11.0 * c1 + 3.0 * c2;
        } else if (i == N + 1) {
            AA(i, i) = 1.0;
        } else {
            c1 /= 180.0 * h * h, c2 /= 60.0 * h;
            AA(i, i - 3) = 2.0 * c1 - 1.0 * c2, AA(i, i - 2) = -27.0 * c1 + 9.0 * c2, AA(i, i - 1) = 270.0 * c1 - 45.0 * c2;
            AA(i, i) = -490.0 * c1, AA(i, i + 1) = 270.0 * c1 + 45.0 * c2, AA(i, i + 2) = -27.0 * c1 - 9.0 * c2, AA(i, i + 3) = 2.0 * c1 + 1.0 * c2;
        }
    }
#undef AA
}

static void spline_generate(int n, const double *x, const double *y, double *sp /*[n-1][4]*/) { /* src/cspline.cpp:66-142 */
    double *A = dalloc(n), *B = dalloc(n), *C = dalloc(n), *D = dalloc(n), *Y = dalloc(n);
    double h0 = x[1] - x[0], h1 = x[2] - x[1], r0 = (y[1] - y[0]) / h0, r1 = (y[2] - y[1]) / h1;
    B[0] = h1 * (h0 + h1);
    C[0] = (h0 + h1) * (h0 + h1);
    Y[0] = r0 * (3 * h0 * h1 + 2 * h1 * h1) + r1 * h0 * h0;
    for (int i = 1; i < n - 1; i++) {
        h0 = x[i] - x[i - 1], h1 = x[i + 1] - x[i];
        r0 = (y[i] - y[i - 1]) / h0, r1 = (y[i + 1] - y[i]) / h1;
        A[i] = h1, B[i] = 2 * (h0 + h1), C[i] = h0, Y[i] = 3 * (r0 * h1 + r1 * h0);
    }
    A[n - 1] = (h0 + h1) * (h0 + h1);
    B[n - 1] = h0 * (h0 + h1);
    Y[n - 1] = r0 * h1 * h1 + r1 * (3 * h0 * h1 + 2 * h0 * h0);
    C[0] = C[0] / B[0];
    for (int i = 1; i < n - 1; i++) C[i] = C[i] / (B[i] - A[i] * C[i - 1]);
    Y[0] = Y[0] / B[0];
    for (int i = 1; i < n; i++) Y[i] = (Y[i] - A[i] * Y[i - 1]) / (B[i] - A[i] * C[i - 1]);
    D[n - 1] = Y[n - 1];
    for (int i = n - 1; i > 0; i--) D[i - 1] = Y[i - 1] - C[i - 1] * D[i];
    for (int i = 0; i < n - 1; i++) {
        const double dx = 1.0 / (x[i + 1] - x[i]), dy = (y[i + 1] - y[i]) * dx;
        sp[4 * i] = y[i], sp[4 * i + 1] = D[i], sp[4 * i + 2] = dx * (3 * dy - 2 * D[i] - D[i + 1]);
        sp[4 * i + 3] = dx * dx * (-2 * dy + D[i] + D[i + 1]);
    }
    free(A), free(B), free(C), free(D), free(Y);
}
static double spline_eval(int n, const double *x, const double *y, const double *sp, double xx) { /* src/cspline.cpp:151-172 */
    if (xx < x[0]) return y[0];
    if (xx >= x[n - 1]) return y[n - 1];
    for (int i = 1; i < n; i++)
        if (xx <= x[i]) {
            const double t = xx - x[i - 1];
            return sp[4 * (i - 1)] + sp[4 * (i - 1) + 1] * t + sp[4 * (i - 1) + 2] * t * t + sp[4 * (i - 1) + 3] * t * t * t;
        }
    return 0.0;
}

void oracle_hartree(oracle_t *o, double *J) {
    const int N = o->nrad, n = N + 2, nlm = o->nlm, nang = o->nang, nb = o->nbf, L = o->lmax;
    const long npa = (long)N * nang;
    const double sqrt4pi = sqrt(4.0 * M_PI);
    double *A = dalloc((size_t)n * n);
    fd_matrix(o, A);
    double *xs = dalloc(N), *ys = dalloc((size_t)o->natoms * nlm * N);
    for (int i = 0; i < N; i++) xs[i] = o->r_n[N - 1 - i];
    for (int at = 0; at < o->natoms; at++) {
        double *rho_lm = o->rho_lm + (size_t)at * N * nlm, *U = o->U_lm + (size_t)at * N * nlm;
        const double q_n = atom_charge(o, at);
        o->q[at] = q_n;
#pragma omp parallel for schedule(static)
        for (int i = 0; i < N; i++) { /* calculate_rho_lm, src/atomicgrid.cpp:258-295 */
            int cnt = 0;
            for (int l = 0; l <= L; l++)
                for (int m = -l; m <= l; m++, cnt++) {
                    const double pre = sh_prefactor(l, m);
                    double acc = 0.0;
                    for (int j = 0; j < nang; j++) {
                        const long idx = at * npa + (long)i * nang + j;
                        const double *pos = o->rat + 3 * idx;
                        const double r = sqrt(pos[0] * pos[0] + pos[1] * pos[1] + pos[2] * pos[2]);
                        const double y_lm = pre * spherical_harmonic(l, m, acos(pos[2] / r), atan2(pos[1], pos[0]));
                        acc += o->rho[idx] * y_lm * o->leb[4 * j + 3] * o->wb[idx];
                    }
                    rho_lm[(size_t)i * nlm + cnt] = acc * (4.0 * M_PI);
                }
        }
#pragma omp parallel for schedule(dynamic)
        for (int lm = 0; lm < nlm; lm++) { /* calculate_U_lm, :395-432 */
            int l = 0;
            while ((l + 1) * (l + 1) <= lm) l++;
            double *M = (double *)malloc(sizeof(double) * (size_t)n * n), *g = dalloc(n);
            memcpy(M, A, sizeof(double) * (size_t)n * n);
            g[0] = lm == 0 ? sqrt4pi * q_n : 0.0;
            for (int i = 1; i < N + 1; i++) {
                M[(size_t)i * n + i] -= (double)l * (double)(l + 1) / (o->r_n[i - 1] * o->r_n[i - 1]);
                g[i] = -4.0 * M_PI * o->r_n[i - 1] * rho_lm[(size_t)(i - 1) * nlm + lm];
            }
            lu_solve_dense(n, M, g);
            for (int i = 1; i < N + 1; i++) U[(size_t)(i - 1) * nlm + lm] = g[i];
            free(M), free(g);
            /* interpolate_sh_coeff, :535-556 */
            double *y = ys + ((size_t)at * nlm + lm) * N;
            for (int i = 0; i < N; i++) y[i] = U[(size_t)(N - 1 - i) * nlm + lm];
            spline_generate(N, xs, y, o->spl + ((size_t)at * nlm + lm) * (N - 1) * 4);
        }
#pragma omp parallel for schedule(static)
        for (long t = 0; t < npa; t++) { /* own-cell potential, :437-462 */
            const long idx = at * npa + t;
            const int i = (int)(t / nang);
            const double *pos = o->rat + 3 * idx;
            const double r = sqrt(pos[0] * pos[0] + pos[1] * pos[1] + pos[2] * pos[2]);
            const double az = atan2(pos[1], pos[0]), pole = acos(pos[2] / r);
            double v = 0.0;
            int cnt = 0;
            for (int l = 0; l <= L; l++)
                for (int m = -l; m <= l; m++, cnt++) v += 1.0 / r * (sh_prefactor(l, m) * spherical_harmonic(l, m, pole, az)) * U[(size_t)i * nlm + cnt];
            o->Vfuzzy[idx] = v;
        }
    }
#pragma omp parallel for schedule(dynamic, 16)
    for (long idx = 0; idx < o->npts; idx++) { /* src/moleculargrid.cpp:342-380 */
        const int own = (int)(idx / npa);
        double V = 0.0;
        for (int k = 0; k < o->natoms; k++) {
            if (k == own) {
                V += o->Vfuzzy[idx];
                continue;
            }
            const double pos[3] = {o->xyz[3 * idx] - o->axyz[3 * k], o->xyz[3 * idx + 1] - o->axyz[3 * k + 1], o->xyz[3 * idx + 2] - o->axyz[3 * k + 2]};
            const double r = sqrt(pos[0] * pos[0] + pos[1] * pos[1] + pos[2] * pos[2]);
            const double az = atan2(pos[1], pos[0]), pole = acos(pos[2] / r);
            int lm = 0;
            for (int l = 0; l <= L; l++)
                for (int m = -l; m <= l; m++, lm++) {
                    const double y_lm = sh_prefactor(l, m) * spherical_harmonic(l, m, pole, az);
                    V += 1.0 / r * y_lm * spline_eval(N, xs, ys + ((size_t)k * nlm + lm) * N, o->spl + ((size_t)k * nlm + lm) * (N - 1) * 4, r);
                }
        }
        o->V[idx] = V;
    }
    /* J = sum over atoms of 0.5 * sum_p w V phi_i phi_j (src/atomicgrid.cpp:471-488, src/moleculargrid.cpp:382-386) */
#pragma omp parallel for schedule(dynamic)
    for (int i = 0; i < nb; i++)
        for (int j = i; j < nb; j++) {
            double tot = 0.0;
            for (int at = 0; at < o->natoms; at++) {
                double s = 0.0;
                for (long p = at * npa; p < (at + 1) * npa; p++) s += (o->w[p] * o->V[p] * o->phi[(size_t)p * nb + i]) * o->phi[(size_t)p * nb + j];
                tot += s * 0.5;
            }
            J[(size_t)i * nb + j] = J[(size_t)j * nb + i] = tot;
        }
    free(A), free(xs), free(ys);
}

/* ---- density dump: src/rectangulargrid.cpp:34-80, src/cgf.cpp:67-94,164-172, src/gridpoint.cpp:82-109 ------------------ */
/* GTO::get_grad as the reference evaluates it (separable exponentials, derivative of the monomial WITHOUT its factor l, no
 * normalisation constant, contraction coefficient applied twice by CGF::get_grad); restated as written, not corrected. */
static void cgf_grad(const oracle_t *o, int b, const double *r, double *g) {
    g[0] = g[1] = g[2] = 0.0;
    for (int k = o->bf_off[b]; k < o->bf_off[b + 1]; k++) {
        const double al = o->alpha[k], c = o->coeff[k];
        const int *lmn = o->lmn + 3 * k;
        const double d[3] = {r[0] - o->bf_center[3 * b], r[1] - o->bf_center[3 * b + 1], r[2] - o->bf_center[3 * b + 2]};
        double e[3], f[3], q[3];
        for (int a = 0; a < 3; a++) {
            e[a] = exp(-al * pow(d[a], 2));
            f[a] = pow(d[a], lmn[a]) * e[a];
            q[a] = -2.0 * al * d[a] * f[a];
            if (lmn[a] > 0) q[a] += pow(d[a], lmn[a] - 1) * e[a];
        }
        g[0] += c * (c * q[0] * f[1] * f[2]);
        g[1] += c * (c * f[0] * q[1] * f[2]);
        g[2] += c * (c * f[0] * f[1] * q[2]);
    }
}

/* RectangularGrid::build_grid(size, dp) + set_density(P): pos [dp^3][3], rho [dp^3], grad [dp^3][3]; P column-major nbf x nbf */
void oracle_rect_density(const oracle_t *o, double size, int dp, const double *P, double *pos, double *rho, double *grad) {
    const int nb = o->nbf;
    const double gd = size / (double)(dp - 1);
#pragma omp parallel
    {
        double *amp = dalloc(nb), *gr = dalloc(3 * (size_t)nb), *Pa = dalloc(nb), *Pg = dalloc(nb);
#pragma omp for schedule(static)
        for (long idx = 0; idx < (long)dp * dp * dp; idx++) {
            const long k = idx % dp, j = (idx / dp) % dp, i = idx / ((long)dp * dp);
            const double r[3] = {(double)k * gd - size / 2.0, (double)j * gd - size / 2.0, (double)i * gd - size / 2.0};
            for (int c = 0; c < 3; c++) pos[3 * idx + c] = r[c];
            for (int b = 0; b < nb; b++) {
                amp[b] = cgf_amp(o, b, r);
                double g3[3];
                cgf_grad(o, b, r, g3);
                for (int c = 0; c < 3; c++) gr[(size_t)c * nb + b] = g3[c];
            }
            /* rho = 2 amp . (P amp)   (src/gridpoint.cpp:82-84) */
            for (int a = 0; a < nb; a++) {
                double s = 0.0;
                for (int b = 0; b < nb; b++) s += P[(size_t)b * nb + a] * amp[b];
                Pa[a] = s;
            }
            double s = 0.0;
            for (int a = 0; a < nb; a++) s += amp[a] * Pa[a];
            rho[idx] = 2.0 * s;
            /* g_c = 2 amp . (P d_c) + 2 d_c . (P amp)   (src/gridpoint.cpp:94-109) */
            for (int c = 0; c < 3; c++) {
                const double *dc = gr + (size_t)c * nb;
                for (int a = 0; a < nb; a++) {
                    double t = 0.0;
                    for (int b = 0; b < nb; b++) t += P[(size_t)b * nb + a] * dc[b];
                    Pg[a] = t;
                }
                double s1 = 0.0, s2 = 0.0;
                for (int a = 0; a < nb; a++) {
                    s1 += amp[a] * Pg[a];
                    s2 += dc[a] * Pa[a];
                }
                grad[3 * idx + c] = 2.0 * s1 + 2.0 * s2;
            }
        }
        free(amp);
        free(gr);
        free(Pa);
        free(Pg);
    }
}

