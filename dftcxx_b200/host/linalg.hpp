// Small dense linear algebra for the host SCF driver (the reference leans on Eigen, which this build does not
// need): a row-major matrix, products, and a symmetric eigen-solver (Householder tridiagonalisation followed by
// implicit-shift QL), eigenvalues ascending like Eigen::SelfAdjointEigenSolver (reference src/dft.cpp:303,339).
// The O(n^3) loops are OpenMP-parallel over independent rows / columns; every element is still computed by the same
// sequence of operations as in the serial algorithm, so results do not depend on the thread count.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <stdexcept>
#include <vector>

namespace dftcxx {

class Mat {
public:
    Mat() : r_(0), c_(0) {}
    Mat(size_t r, size_t c, double v = 0.0) : r_(r), c_(c), d_(r * c, v) {}
    size_t rows() const { return r_; }
    size_t cols() const { return c_; }
    double& operator()(size_t i, size_t j) { return d_[i * c_ + j]; }
    double operator()(size_t i, size_t j) const { return d_[i * c_ + j]; }
    double* data() { return d_.data(); }
    const double* data() const { return d_.data(); }
    void fill(double v) { std::fill(d_.begin(), d_.end(), v); }

private:
    size_t r_, c_;
    std::vector<double> d_;
};

inline Mat matmul(const Mat& a, const Mat& b) {
    if (a.cols() != b.rows()) throw std::runtime_error("matmul: shape mismatch");
    Mat c(a.rows(), b.cols(), 0.0);
    const size_t n = a.rows(), m = b.cols(), k = a.cols();
    // c(i,j) = sum_l a(i,l) b(l,j) with l ascending for every element; blocked 4 rows x 256 columns so that a strip of b
    // is reused by four rows while their c strip stays in L1
    constexpr size_t RB = 4, CB = 256;
#pragma omp parallel for schedule(static) if (n * m * k > 200000)
    for (long ib = 0; ib < (long)((n + RB - 1) / RB); ib++) {
        const size_t i0 = (size_t)ib * RB, i1 = std::min(n, i0 + RB);
        for (size_t j0 = 0; j0 < m; j0 += CB) {
            const size_t j1 = std::min(m, j0 + CB);
            for (size_t l = 0; l < k; l++) {
                const double* br = b.data() + l * m;
                for (size_t i = i0; i < i1; i++) {
                    const double ail = a(i, l);
                    double* cr = c.data() + i * m;
                    for (size_t j = j0; j < j1; j++) cr[j] += ail * br[j];
                }
            }
        }
    }
    return c;
}

inline Mat transpose(const Mat& a) {
    Mat t(a.cols(), a.rows());
    for (size_t i = 0; i < a.rows(); i++)
        for (size_t j = 0; j < a.cols(); j++) t(j, i) = a(i, j);
    return t;
}

inline double trace_of_product(const Mat& a, const Mat& b) {  // tr(A B)
    double s = 0.0;
    for (size_t i = 0; i < a.rows(); i++)
        for (size_t j = 0; j < a.cols(); j++) s += a(i, j) * b(j, i);
    return s;
}

// Eigen-decomposition of a real symmetric matrix: A = V diag(w) V^T, w ascending, eigenvectors in the COLUMNS of V.
inline void sym_eigen(const Mat& A, std::vector<double>& w, Mat& V) {
    const int n = (int)A.rows();
    if ((int)A.cols() != n) throw std::runtime_error("sym_eigen: matrix not square");
    V = A;
    w.assign(n, 0.0);
    std::vector<double> e(n, 0.0), rot_s, rot_c, gvec;
    if (n == 0) return;
    // --- Householder reduction to tridiagonal form, accumulating the orthogonal transformation in V
    for (int i = n - 1; i > 0; i--) {
        const int l = i - 1;
        double h = 0.0, scale = 0.0;
        if (l > 0) {
            for (int k = 0; k <= l; k++) scale += std::fabs(V(i, k));
            if (scale == 0.0) {
                e[i] = V(i, l);
            } else {
                for (int k = 0; k <= l; k++) {
                    V(i, k) /= scale;
                    h += V(i, k) * V(i, k);
                }
                double f = V(i, l);
                double g = f >= 0.0 ? -std::sqrt(h) : std::sqrt(h);
                e[i] = scale * g;
                h -= f * g;
                V(i, l) = f - g;
                // e = A v / h over the stored lower triangle: e_j = sum_{k<=j} A(j,k) v_k + sum_{k>j} A(k,j) v_k, every sum
                // in ascending k.  The second part is gathered row by row (contiguous) instead of down column j; threads
                // own ranges of j.
#pragma omp parallel if (l > 96)
                {
#pragma omp for schedule(dynamic, 16)
                    for (int j = 0; j <= l; j++) {
                        V(j, i) = V(i, j) / h;
                        double gj = 0.0;
                        for (int k = 0; k <= j; k++) gj += V(j, k) * V(i, k);
                        e[j] = gj;
                    }
#pragma omp for schedule(static)
                    for (int jb = 0; jb <= l; jb += 64) {
                        const int je = std::min(jb + 64, l + 1);
                        for (int k = jb + 1; k <= l; k++) {
                            const double vk = V(i, k);
                            const double* rk = &V(k, 0);
                            const int jmax = std::min(je, k);
                            for (int j = jb; j < jmax; j++) e[j] += rk[j] * vk;
                        }
                    }
#pragma omp for schedule(static)
                    for (int j = 0; j <= l; j++) e[j] /= h;
                }
                f = 0.0;
                for (int j = 0; j <= l; j++) f += e[j] * V(i, j);
                const double hh = f / (h + h);
                for (int j = 0; j <= l; j++) e[j] -= hh * V(i, j);
                // rank-2 update of the lower triangle, row by row
#pragma omp parallel for schedule(dynamic, 16) if (l > 96)
                for (int j = 0; j <= l; j++) {
                    const double fj = V(i, j), gj = e[j];
                    for (int k = 0; k <= j; k++) V(j, k) -= fj * e[k] + gj * V(i, k);
                }
            }
        } else {
            e[i] = V(i, l);
        }
        w[i] = h;
    }
    w[0] = 0.0;
    e[0] = 0.0;
    for (int i = 0; i < n; i++) {
        const int l = i - 1;
        if (w[i] != 0.0) {
            // g_j = sum_k V(i,k) V(k,j) (ascending k), then V(k,j) -= g_j V(k,i): both walked along rows
            gvec.assign(l + 1, 0.0);
#pragma omp parallel if (l > 96)
            {
#pragma omp for schedule(static)
                for (int jb = 0; jb <= l; jb += 64) {
                    const int je = std::min(jb + 64, l + 1);
                    for (int k = 0; k <= l; k++) {
                        const double vik = V(i, k);
                        const double* rk = &V(k, 0);
                        for (int j = jb; j < je; j++) gvec[j] += vik * rk[j];
                    }
                }
#pragma omp for schedule(static)
                for (int k = 0; k <= l; k++) {
                    const double vki = V(k, i);
                    double* rk = &V(k, 0);
                    for (int j = 0; j <= l; j++) rk[j] -= gvec[j] * vki;
                }
            }
        }
        w[i] = V(i, i);
        V(i, i) = 1.0;
        for (int j = 0; j <= l; j++) V(j, i) = V(i, j) = 0.0;
    }
    // --- implicit QL on the tridiagonal matrix (diagonal w, sub-diagonal e)
    for (int i = 1; i < n; i++) e[i - 1] = e[i];
    e[n - 1] = 0.0;
    for (int l = 0; l < n; l++) {
        int iter = 0, m;
        do {
            for (m = l; m < n - 1; m++) {
                const double dd = std::fabs(w[m]) + std::fabs(w[m + 1]);
                if (std::fabs(e[m]) <= 2.3e-16 * dd) break;
            }
            if (m != l) {
                if (iter++ == 200) throw std::runtime_error("sym_eigen: QL iteration did not converge");
                double g = (w[l + 1] - w[l]) / (2.0 * e[l]);
                double r = std::hypot(g, 1.0);
                g = w[m] - w[l] + e[l] / (g + (g >= 0.0 ? std::fabs(r) : -std::fabs(r)));
                double s = 1.0, c = 1.0, p = 0.0;
                int i;
                // the sweep's plane rotations are recorded and then applied to the eigenvector rows in parallel: row k
                // sees them in the same order as in the serial algorithm
                rot_s.clear();
                rot_c.clear();
                const int i_first = m - 1;
                for (i = m - 1; i >= l; i--) {
                    double f = s * e[i];
                    const double b = c * e[i];
                    e[i + 1] = (r = std::hypot(f, g));
                    if (r == 0.0) {
                        w[i + 1] -= p;
                        e[m] = 0.0;
                        break;
                    }
                    s = f / r;
                    c = g / r;
                    g = w[i + 1] - p;
                    r = (w[i] - g) * s + 2.0 * c * b;
                    w[i + 1] = g + (p = s * r);
                    g = c * r - b;
                    rot_s.push_back(s);
                    rot_c.push_back(c);
                }
                {
                    const int nrot = (int)rot_s.size();
#pragma omp parallel for schedule(static) if ((long)nrot * n > 20000)
                    for (int k = 0; k < n; k++) {
                        double* row = &V(k, 0);
                        for (int t = 0; t < nrot; t++) {
                            const int ii = i_first - t;
                            const double fk = row[ii + 1];
                            row[ii + 1] = rot_s[t] * row[ii] + rot_c[t] * fk;
                            row[ii] = rot_c[t] * row[ii] - rot_s[t] * fk;
                        }
                    }
                }
                if (r == 0.0 && i >= l) continue;
                w[l] -= p;
                e[l] = g;
                e[m] = 0.0;
            }
        } while (m != l);
    }
    // --- ascending order
    std::vector<int> idx(n);
    for (int i = 0; i < n; i++) idx[i] = i;
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return w[a] < w[b]; });
    std::vector<double> ws(n);
    Mat Vs(n, n);
    for (int j = 0; j < n; j++) {
        ws[j] = w[idx[j]];
        for (int i = 0; i < n; i++) Vs(i, j) = V(i, idx[j]);
    }
    w.swap(ws);
    V = Vs;
}

}  // namespace dftcxx
