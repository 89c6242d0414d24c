// ORACLE / TEST INFRASTRUCTURE. A ~300-line fixed-size subset of Eigen 3.3 written for this repo (no Eigen code):
// just enough API for /root/reference/Library/Math/{Distance/*.h, BARRIER.h, UTILS.h} to compile unchanged.
// Semantics transcribed from Eigen that affect results:
//   * 3-term reductions (dot / squaredNorm) associate as x0 + (x1 + x2); n-term ones split in halves recursively;
//   * LDLT is the pivoted in-place algorithm (largest |diagonal| first), solve zeroes pivots <= numeric_limits::min;
//   * SelfAdjointEigenSolver: eigenvalues ascending, lower triangle read (implemented with cyclic Jacobi).
// Expressions are evaluated eagerly, coefficient by coefficient, which matches Eigen's lazy coefficient-wise order.
#pragma once
#include <cmath>
#include <cstring>
#include <vector>
#include <algorithm>
#include <limits>
#include <iostream>

namespace Eigen {

const int Dynamic = -1;

template <class T, int R, int C> struct Matrix;
template <class T, int R, int C> struct Array;

template <class T> inline T redux_sum(const T* x, int start, int len)
{
    if (len == 1) return x[start];
    const int half = len / 2;
    return redux_sum(x, start, half) + redux_sum(x, start + half, len - half);
}

template <class T, int R, int C>
struct Matrix {
    T d[R * C]; // column major
    Matrix() {}
    explicit Matrix(const T* p) { for (int i = 0; i < R * C; ++i) d[i] = p[i]; }
    Matrix(T a, T b) { static_assert(R * C == 2, "size"); d[0] = a; d[1] = b; }
    Matrix(T a, T b, T c) { static_assert(R * C == 3, "size"); d[0] = a; d[1] = b; d[2] = c; }
    // vectors convert between row and column orientation on assignment / construction (as in Eigen)
    template <int R2, int C2, class = typename std::enable_if<(R2 == C && C2 == R && (R == 1 || C == 1) && R != C)>::type>
    Matrix(const Matrix<T, R2, C2>& o) { for (int i = 0; i < R * C; ++i) d[i] = o.d[i]; }
    Matrix(const Array<T, R, C>& a);
    static Matrix Zero() { Matrix m; m.setZero(); return m; }
    T& operator()(int i, int j) { return d[i + j * R]; }
    const T& operator()(int i, int j) const { return d[i + j * R]; }
    T& operator()(int i) { return d[i]; }
    const T& operator()(int i) const { return d[i]; }
    T& operator[](int i) { return d[i]; }
    const T& operator[](int i) const { return d[i]; }
    T* data() { return d; }
    const T* data() const { return d; }
    int rows() const { return R; }
    int cols() const { return C; }
    void setZero() { for (int i = 0; i < R * C; ++i) d[i] = T(0); }
    Matrix operator-() const { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = -d[i]; return m; }
    Matrix operator+(const Matrix& o) const { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = d[i] + o.d[i]; return m; }
    Matrix operator-(const Matrix& o) const { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = d[i] - o.d[i]; return m; }
    Matrix operator*(T s) const { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = d[i] * s; return m; }
    Matrix operator/(T s) const { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = d[i] / s; return m; }
    Matrix& operator+=(const Matrix& o) { for (int i = 0; i < R * C; ++i) d[i] += o.d[i]; return *this; }
    Matrix& operator-=(const Matrix& o) { for (int i = 0; i < R * C; ++i) d[i] -= o.d[i]; return *this; }
    Matrix& operator*=(T s) { for (int i = 0; i < R * C; ++i) d[i] *= s; return *this; }
    Matrix<T, C, R> transpose() const
    {
        Matrix<T, C, R> m;
        for (int i = 0; i < R; ++i) for (int j = 0; j < C; ++j) m(j, i) = (*this)(i, j);
        return m;
    }
    template <int C2> Matrix<T, R, C2> operator*(const Matrix<T, C, C2>& o) const
    {
        Matrix<T, R, C2> m;
        for (int i = 0; i < R; ++i)
            for (int j = 0; j < C2; ++j) {
                T t[C];
                for (int k = 0; k < C; ++k) t[k] = (*this)(i, k) * o(k, j);
                m(i, j) = redux_sum(t, 0, C);
            }
        return m;
    }
    // vector API (any orientation)
    template <int R2, int C2> T dot(const Matrix<T, R2, C2>& o) const
    {
        static_assert(R2 * C2 == R * C, "size");
        T t[R * C];
        for (int i = 0; i < R * C; ++i) t[i] = d[i] * o.d[i];
        return redux_sum(t, 0, R * C);
    }
    T squaredNorm() const
    {
        T t[R * C];
        for (int i = 0; i < R * C; ++i) t[i] = d[i] * d[i];
        return redux_sum(t, 0, R * C);
    }
    T norm() const { return std::sqrt(squaredNorm()); }
    template <int R2, int C2> Matrix cross(const Matrix<T, R2, C2>& o) const
    {
        static_assert(R * C == 3 && R2 * C2 == 3, "cross needs 3-vectors");
        Matrix m;
        m.d[0] = d[1] * o.d[2] - d[2] * o.d[1];
        m.d[1] = d[2] * o.d[0] - d[0] * o.d[2];
        m.d[2] = d[0] * o.d[1] - d[1] * o.d[0];
        return m;
    }
    Array<T, R, C> array() const;

    struct RowRef {
        Matrix& m; int i;
        operator Matrix<T, 1, C>() const { Matrix<T, 1, C> r; for (int j = 0; j < C; ++j) r.d[j] = m(i, j); return r; }
        RowRef& operator=(const Matrix<T, 1, C>& r) { for (int j = 0; j < C; ++j) m(i, j) = r.d[j]; return *this; }
        RowRef& operator=(const RowRef& r) { return *this = (Matrix<T, 1, C>)r; }
        template <int R2, int C2> Matrix<T, 1, C> cross(const Matrix<T, R2, C2>& o) const { return ((Matrix<T, 1, C>)*this).cross(o); }
        Matrix<T, 1, C> cross(const RowRef& o) const { return ((Matrix<T, 1, C>)*this).cross((Matrix<T, 1, C>)o); }
        Matrix<T, C, 1> transpose() const { return ((Matrix<T, 1, C>)*this).transpose(); }
    };
    struct ColRef {
        Matrix& m; int j;
        operator Matrix<T, R, 1>() const { Matrix<T, R, 1> r; for (int i = 0; i < R; ++i) r.d[i] = m(i, j); return r; }
        ColRef& operator=(const Matrix<T, R, 1>& r) { for (int i = 0; i < R; ++i) m(i, j) = r.d[i]; return *this; }
    };
    RowRef row(int i) { return RowRef{*this, i}; }
    ColRef col(int j) { return ColRef{*this, j}; }
    template <int N> struct SegRef {
        Matrix& m; int s;
        operator Matrix<T, N, 1>() const { Matrix<T, N, 1> r; for (int i = 0; i < N; ++i) r.d[i] = m.d[s + i]; return r; }
        SegRef& operator=(const Matrix<T, N, 1>& r) { for (int i = 0; i < N; ++i) m.d[s + i] = r.d[i]; return *this; }
        Matrix<T, N, 1> operator-() const { return -((Matrix<T, N, 1>)*this); }
    };
    template <int N> SegRef<N> segment(int s) { return SegRef<N>{*this, s}; }
    struct DiagRef {
        Matrix& m;
        void setConstant(T v) { for (int i = 0; i < (R < C ? R : C); ++i) m(i, i) = v; }
        T& operator[](int i) { return m(i, i); }
    };
    DiagRef diagonal() { return DiagRef{*this}; }

    // pivoted LDLT of a symmetric matrix (lower triangle), Eigen 3.3 ldlt_inplace<Lower>::unblocked + _solve_impl
    struct LDLT {
        Matrix L; int tr[R];
        Matrix<T, R, 1> solve(const Matrix<T, R, 1>& b) const
        {
            Matrix<T, R, 1> x = b;
            for (int k = 0; k < R; ++k) std::swap(x.d[k], x.d[tr[k]]);
            for (int i = 0; i < R; ++i) for (int k = 0; k < i; ++k) x.d[i] -= L(i, k) * x.d[k];
            const T tol = (std::numeric_limits<T>::min)();
            for (int i = 0; i < R; ++i) x.d[i] = (std::fabs(L(i, i)) > tol) ? x.d[i] / L(i, i) : T(0);
            for (int i = R - 1; i >= 0; --i) for (int k = i + 1; k < R; ++k) x.d[i] -= L(k, i) * x.d[k];
            for (int k = R - 1; k >= 0; --k) std::swap(x.d[k], x.d[tr[k]]);
            return x;
        }
    };
    LDLT ldlt() const
    {
        static_assert(R == C, "square");
        LDLT f; f.L = *this;
        Matrix& m = f.L;
        for (int k = 0; k < R; ++k) {
            int big = k;
            for (int i = k + 1; i < R; ++i) if (std::fabs(m(i, i)) > std::fabs(m(big, big))) big = i;
            f.tr[k] = big;
            if (big != k) { // symmetric transposition on the lower triangle
                for (int j = 0; j < k; ++j) std::swap(m(k, j), m(big, j));
                for (int i = big + 1; i < R; ++i) std::swap(m(i, k), m(i, big));
                std::swap(m(k, k), m(big, big));
                for (int i = k + 1; i < big; ++i) std::swap(m(i, k), m(big, i));
            }
            for (int j = 0; j < k; ++j) m(k, k) -= m(k, j) * (m(j, j) * m(k, j));
            for (int i = k + 1; i < R; ++i) for (int j = 0; j < k; ++j) m(i, k) -= m(i, j) * (m(j, j) * m(k, j));
            const bool valid = std::fabs(m(k, k)) > T(0);
            if (k == 0 && !valid) { for (int j = 1; j < R; ++j) f.tr[j] = j; break; }
            if (valid) for (int i = k + 1; i < R; ++i) m(i, k) /= m(k, k);
        }
        return f;
    }
};
template <class T, int R, int C> inline Matrix<T, R, C> operator*(T s, const Matrix<T, R, C>& m) { Matrix<T, R, C> r; for (int i = 0; i < R * C; ++i) r.d[i] = s * m.d[i]; return r; }
template <class T, int R, int C> inline Matrix<T, R, C> operator*(int s, const Matrix<T, R, C>& m) { return T(s) * m; }
template <class T, int R, int C> inline std::ostream& operator<<(std::ostream& o, const Matrix<T, R, C>& m) { for (int i = 0; i < R * C; ++i) o << m.d[i] << ' '; return o; }

template <class T, int R, int C>
struct Array {
    T d[R * C];
    Array max(const Array& o) const { Array a; for (int i = 0; i < R * C; ++i) a.d[i] = std::max(d[i], o.d[i]); return a; }
    Array min(const Array& o) const { Array a; for (int i = 0; i < R * C; ++i) a.d[i] = std::min(d[i], o.d[i]); return a; }
    Array operator-(const Array& o) const { Array a; for (int i = 0; i < R * C; ++i) a.d[i] = d[i] - o.d[i]; return a; }
    Array operator-(T s) const { Array a; for (int i = 0; i < R * C; ++i) a.d[i] = d[i] - s; return a; }
    Array operator+(T s) const { Array a; for (int i = 0; i < R * C; ++i) a.d[i] = d[i] + s; return a; }
    Array<bool, R, C> operator>(T s) const { Array<bool, R, C> a; for (int i = 0; i < R * C; ++i) a.d[i] = d[i] > s; return a; }
    bool any() const { for (int i = 0; i < R * C; ++i) if (d[i]) return true; return false; }
};
template <class T, int R, int C> Array<T, R, C> Matrix<T, R, C>::array() const { Array<T, R, C> a; for (int i = 0; i < R * C; ++i) a.d[i] = d[i]; return a; }
template <class T, int R, int C> Matrix<T, R, C>::Matrix(const Array<T, R, C>& a) { for (int i = 0; i < R * C; ++i) d[i] = a.d[i]; }

typedef Matrix<double, 1, 3> RowVector3d;

template <class T, int N>
struct DiagonalMatrix {
    Matrix<T, N, 1> v;
    DiagonalMatrix(const Matrix<T, N, 1>& x) : v(x) {}
    Matrix<T, N, 1>& diagonal() { return v; }
};
template <class T, int N> inline Matrix<T, N, N> operator*(const Matrix<T, N, N>& m, const DiagonalMatrix<T, N>& D)
{
    Matrix<T, N, N> r;
    for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) r(i, j) = m(i, j) * D.v.d[j];
    return r;
}

template <class MatT> struct SelfAdjointEigenSolver;
template <class T, int N>
struct SelfAdjointEigenSolver<Matrix<T, N, N>> {
    Matrix<T, N, 1> lam;
    Matrix<T, N, N> vec;
    explicit SelfAdjointEigenSolver(const Matrix<T, N, N>& Ain)
    {
        T A[N][N], V[N][N];
        for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) { A[i][j] = (i >= j) ? Ain(i, j) : Ain(j, i); V[i][j] = (i == j); }
        for (int sweep = 0; sweep < 100; ++sweep) {
            T off = 0, tot = 0;
            for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) { tot += A[i][j] * A[i][j]; if (i != j) off += A[i][j] * A[i][j]; }
            if (off <= 1e-34 * tot || off == 0) break;
            for (int p = 0; p < N - 1; ++p)
                for (int q = p + 1; q < N; ++q) {
                    if (A[p][q] == 0) continue;
                    const T th = (A[q][q] - A[p][p]) / (2 * A[p][q]);
                    const T t = (th >= 0 ? 1 : -1) / (std::fabs(th) + std::sqrt(th * th + 1));
                    const T c = 1 / std::sqrt(t * t + 1), s = t * c;
                    for (int k = 0; k < N; ++k) { const T a = A[k][p], b = A[k][q]; A[k][p] = c * a - s * b; A[k][q] = s * a + c * b; }
                    for (int k = 0; k < N; ++k) { const T a = A[p][k], b = A[q][k]; A[p][k] = c * a - s * b; A[q][k] = s * a + c * b; }
                    for (int k = 0; k < N; ++k) { const T a = V[k][p], b = V[k][q]; V[k][p] = c * a - s * b; V[k][q] = s * a + c * b; }
                }
        }
        int idx[N];
        for (int i = 0; i < N; ++i) idx[i] = i;
        std::sort(idx, idx + N, [&](int a, int b) { return A[a][a] < A[b][b]; });
        for (int j = 0; j < N; ++j) { lam.d[j] = A[idx[j]][idx[j]]; for (int k = 0; k < N; ++k) vec(k, j) = V[k][idx[j]]; }
    }
    const Matrix<T, N, 1>& eigenvalues() const { return lam; }
    const Matrix<T, N, N>& eigenvectors() const { return vec; }
};

} // namespace Eigen
