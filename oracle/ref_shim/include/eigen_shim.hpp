// ORACLE / TEST INFRASTRUCTURE. A ~300-line fixed-size subset of Eigen 3.3 written for this repo (no Eigen code):
// just enough API for /root/reference/Library/Math/{Distance/*.h, BARRIER.h, UTILS.h} to compile unchanged.
// Semantics transcribed from Eigen that affect results:
//   * 3-term reductions (dot / squaredNorm) associate as x0 + (x1 + x2); n-term ones split in halves recursively;
//   * LDLT is the pivoted in-place algorithm (largest |diagonal| first), solve zeroes pivots <= numeric_limits::min;
//   * SelfAdjointEigenSolver: eigenvalues ascending, lower triangle read (Householder tridiagonalisation + implicit QL).
// Expressions are evaluated eagerly, coefficient by coefficient, which matches Eigen's lazy coefficient-wise order.
// For FEM/IPC.h and Grid/SPATIAL_HASH.h (ref_ipc_capi.cpp) it also has: run-time sized Matrix<T, Dynamic, C> (row(),
// colwise().minCoeff()/maxCoeff(), mean() = zero-padded power-of-two tree sum / n, the order the oracle and the CUDA path
// use -- Eigen's own vectorised order is not recoverable offline), coefficient-wise Array<T,R,C> with Matrix <-> Array and
// row <-> column vector conversions, Triplet, and a 3x3 fullPivLu().solve (only used by dead code of IPC.h).
#pragma once
#include <cmath>
#include <cstring>
#include <vector>
#include <algorithm>
#include <limits>
#include <iostream>
#include <cassert>
#include <numeric>
#include <type_traits>

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
    template <int R2, int C2, class = typename std::enable_if<(R2 == C && C2 == R && R != C)>::type>
    Matrix(const Array<T, R2, C2>& a) { for (int i = 0; i < R * C; ++i) d[i] = a.d[i]; }
    static Matrix Zero() { Matrix m; m.setZero(); return m; }
    static Matrix Identity() { Matrix m; m.setZero(); for (int i = 0; i < (R < C ? R : C); ++i) m(i, i) = T(1); return m; }
    static Matrix UnitX() { Matrix m; m.setZero(); m.d[0] = T(1); return m; }
    static Matrix UnitY() { Matrix m; m.setZero(); m.d[1] = T(1); return m; }
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
    Matrix normalized() const { Matrix r = *this; const T n = norm(); for (int i = 0; i < R * C; ++i) r.d[i] = d[i] / n; return r; }
    T prod() const { T p = d[0]; for (int i = 1; i < R * C; ++i) p *= d[i]; return p; }
    // Gaussian elimination with full pivoting (3x3 use in dead code of FEM/IPC.h; not result relevant)
    struct FullPivLU {
        Matrix A;
        Matrix<T, R, 1> solve(const Matrix<T, R, 1>& b) const
        {
            Matrix M = A; Matrix<T, R, 1> x = b; int perm[R];
            for (int i = 0; i < R; ++i) perm[i] = i;
            for (int k = 0; k < R; ++k) {
                int pi = k, pj = k; T best = -1;
                for (int i = k; i < R; ++i) for (int j = k; j < R; ++j) if (std::fabs(M(i, j)) > best) { best = std::fabs(M(i, j)); pi = i; pj = j; }
                for (int j = 0; j < R; ++j) std::swap(M(k, j), M(pi, j));
                std::swap(x.d[k], x.d[pi]);
                for (int i = 0; i < R; ++i) std::swap(M(i, k), M(i, pj));
                std::swap(perm[k], perm[pj]);
                if (M(k, k) == T(0)) continue;
                for (int i = k + 1; i < R; ++i) { const T f = M(i, k) / M(k, k); for (int j = k; j < R; ++j) M(i, j) -= f * M(k, j); x.d[i] -= f * x.d[k]; }
            }
            Matrix<T, R, 1> y;
            for (int i = R - 1; i >= 0; --i) { T t = x.d[i]; for (int j = i + 1; j < R; ++j) t -= M(i, j) * y.d[j]; y.d[i] = (M(i, i) != T(0)) ? t / M(i, i) : T(0); }
            Matrix<T, R, 1> out;
            for (int i = 0; i < R; ++i) out.d[perm[i]] = y.d[i];
            return out;
        }
    };
    FullPivLU fullPivLu() const { return FullPivLU{*this}; }
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
        Matrix<T, 1, C> operator-() const { return -((Matrix<T, 1, C>)*this); }
    };
    // comma initialiser, rows at a time (m << row0, row1, ...)
    struct RowComma {
        Matrix& m; int i;
        RowComma& operator,(const Matrix<T, 1, C>& r) { for (int j = 0; j < C; ++j) m(i, j) = r.d[j]; ++i; return *this; }
    };
    RowComma operator<<(const Matrix<T, 1, C>& r) { RowComma c{*this, 0}; c, r; return c; }
    struct ColRef {
        Matrix& m; int j;
        operator Matrix<T, R, 1>() const { Matrix<T, R, 1> r; for (int i = 0; i < R; ++i) r.d[i] = m(i, j); return r; }
        ColRef& operator=(const Matrix<T, R, 1>& r) { for (int i = 0; i < R; ++i) m(i, j) = r.d[i]; return *this; }
        template <int R2, int C2> ColRef& operator=(const Matrix<T, R2, C2>& r) { static_assert(R2 * C2 == R, "size"); for (int i = 0; i < R; ++i) m(i, j) = r.d[i]; return *this; }
        Matrix<T, R, 1> cross(const ColRef& o) const { return ((Matrix<T, R, 1>)*this).cross((Matrix<T, R, 1>)o); }
        template <int R2, int C2> Matrix<T, R, 1> cross(const Matrix<T, R2, C2>& o) const { return ((Matrix<T, R, 1>)*this).cross(o); }
    };
    T determinant() const
    {
        static_assert(R == 3 && C == 3, "3x3 only");
        const Matrix& a = *this;
        return a(0, 0) * (a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1)) - a(0, 1) * (a(1, 0) * a(2, 2) - a(1, 2) * a(2, 0)) + a(0, 2) * (a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0));
    }
    RowRef row(int i) { return RowRef{*this, i}; }
    ColRef col(int j) { return ColRef{*this, j}; }
    template <int N> struct SegRef {
        Matrix& m; int s;
        operator Matrix<T, N, 1>() const { Matrix<T, N, 1> r; for (int i = 0; i < N; ++i) r.d[i] = m.d[s + i]; return r; }
        SegRef& operator=(const Matrix<T, N, 1>& r) { for (int i = 0; i < N; ++i) m.d[s + i] = r.d[i]; return *this; }
        Matrix<T, N, 1> operator-() const { return -((Matrix<T, N, 1>)*this); }
        friend Matrix<T, N, 1> operator*(T a, const SegRef& r) { Matrix<T, N, 1> o; for (int i = 0; i < N; ++i) o.d[i] = a * r.m.d[r.s + i]; return o; }
    };
    template <int N> SegRef<N> segment(int s) { return SegRef<N>{*this, s}; }
    template <int BR, int BC> struct BlockRef {
        Matrix& m; int i0, j0;
        operator Matrix<T, BR, BC>() const { Matrix<T, BR, BC> r; for (int i = 0; i < BR; ++i) for (int j = 0; j < BC; ++j) r(i, j) = m(i0 + i, j0 + j); return r; }
        template <class O> BlockRef& operator+=(const O& o) { const Matrix<T, BR, BC> v = o; for (int i = 0; i < BR; ++i) for (int j = 0; j < BC; ++j) m(i0 + i, j0 + j) += v(i, j); return *this; }
        template <class O> BlockRef& operator-=(const O& o) { const Matrix<T, BR, BC> v = o; for (int i = 0; i < BR; ++i) for (int j = 0; j < BC; ++j) m(i0 + i, j0 + j) -= v(i, j); return *this; }
        template <class O> BlockRef& operator=(const O& o) { const Matrix<T, BR, BC> v = o; for (int i = 0; i < BR; ++i) for (int j = 0; j < BC; ++j) m(i0 + i, j0 + j) = v(i, j); return *this; }
        Matrix<T, BC, BR> transpose() const { return ((Matrix<T, BR, BC>)*this).transpose(); }
        void setZero() { for (int i = 0; i < BR; ++i) for (int j = 0; j < BC; ++j) m(i0 + i, j0 + j) = T(0); }
        struct BlockDiag { Matrix& m; int i0, j0; void setConstant(T v) { for (int i = 0; i < (BR < BC ? BR : BC); ++i) m(i0 + i, j0 + i) = v; } };
        BlockDiag diagonal() { return BlockDiag{m, i0, j0}; }
    };
    template <int BR, int BC> BlockRef<BR, BC> block(int i, int j) { return BlockRef<BR, BC>{*this, i, j}; }
    struct DiagRef {
        Matrix& m;
        void setConstant(T v) { for (int i = 0; i < (R < C ? R : C); ++i) m(i, i) = v; }
        T& operator[](int i) { return m(i, i); }
        struct ArrRef { Matrix& m; ArrRef& operator+=(T v) { for (int i = 0; i < (R < C ? R : C); ++i) m(i, i) += v; return *this; } };
        ArrRef array() { return ArrRef{m}; }
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
    Array() {}
    // vectors convert between orientations; matrices convert to arrays implicitly (Eigen: MatrixBase -> Array assignment)
    template <int R2, int C2, class = typename std::enable_if<(R2 * C2 == R * C)>::type>
    Array(const Matrix<T, R2, C2>& m) { for (int i = 0; i < R * C; ++i) d[i] = m.d[i]; }
    template <int R2, int C2, class = typename std::enable_if<(R2 == C && C2 == R && R != C)>::type>
    Array(const Array<T, R2, C2>& o) { for (int i = 0; i < R * C; ++i) d[i] = o.d[i]; }
    static Array Zero() { Array a; for (int i = 0; i < R * C; ++i) a.d[i] = T(0); return a; }
    static Array Ones() { Array a; for (int i = 0; i < R * C; ++i) a.d[i] = T(1); return a; }
    void setOnes() { for (int i = 0; i < R * C; ++i) d[i] = T(1); }
    void setZero() { for (int i = 0; i < R * C; ++i) d[i] = T(0); }
    T& operator[](int i) { return d[i]; }
    const T& operator[](int i) const { return d[i]; }
    T& operator()(int i) { return d[i]; }
    const T& operator()(int i) const { return d[i]; }
    Array max(const Array& o) const { Array a; for (int i = 0; i < R * C; ++i) a.d[i] = std::max(d[i], o.d[i]); return a; }
    Array min(const Array& o) const { Array a; for (int i = 0; i < R * C; ++i) a.d[i] = std::min(d[i], o.d[i]); return a; }
    Array operator-(const Array& o) const { Array a; for (int i = 0; i < R * C; ++i) a.d[i] = d[i] - o.d[i]; return a; }
    Array operator+(const Array& o) const { Array a; for (int i = 0; i < R * C; ++i) a.d[i] = d[i] + o.d[i]; return a; }
    Array operator-(T s) const { Array a; for (int i = 0; i < R * C; ++i) a.d[i] = d[i] - s; return a; }
    Array operator+(T s) const { Array a; for (int i = 0; i < R * C; ++i) a.d[i] = d[i] + s; return a; }
    Array operator*(T s) const { Array a; for (int i = 0; i < R * C; ++i) a.d[i] = d[i] * s; return a; }
    Array<bool, R, C> operator>(T s) const { Array<bool, R, C> a; for (int i = 0; i < R * C; ++i) a.d[i] = d[i] > s; return a; }
    bool any() const { for (int i = 0; i < R * C; ++i) if (d[i]) return true; return false; }
    Array ceil() const { Array a; for (int i = 0; i < R * C; ++i) a.d[i] = std::ceil(d[i]); return a; }
    Array floor() const { Array a; for (int i = 0; i < R * C; ++i) a.d[i] = std::floor(d[i]); return a; }
    template <class U> Array<U, R, C> cast() const { Array<U, R, C> a; for (int i = 0; i < R * C; ++i) a.d[i] = static_cast<U>(d[i]); return a; }
    T prod() const { T p = d[0]; for (int i = 1; i < R * C; ++i) p *= d[i]; return p; }
    T minCoeff() const { T m = d[0]; for (int i = 1; i < R * C; ++i) m = std::min(m, d[i]); return m; }
    T maxCoeff() const { T m = d[0]; for (int i = 1; i < R * C; ++i) m = std::max(m, d[i]); return m; }
    Matrix<T, R, C> matrix() const { Matrix<T, R, C> m; for (int i = 0; i < R * C; ++i) m.d[i] = d[i]; return m; }
    Array<T, C, R> transpose() const { Array<T, C, R> a; for (int i = 0; i < R * C; ++i) a.d[i] = d[i]; return a; } // vectors only
};
template <class T, int R, int C> Array<T, R, C> Matrix<T, R, C>::array() const { Array<T, R, C> a; for (int i = 0; i < R * C; ++i) a.d[i] = d[i]; return a; }
template <class T, int R, int C> Matrix<T, R, C>::Matrix(const Array<T, R, C>& a) { for (int i = 0; i < R * C; ++i) d[i] = a.d[i]; }

// run-time sized matrix with a fixed number of columns (column major)
template <class T, int C>
struct Matrix<T, Dynamic, C> {
    std::vector<T> d;
    int nr = 0;
    Matrix() {}
    Matrix(long rows, long cols) : d((size_t)rows * cols), nr((int)rows) { (void)cols; }
    explicit Matrix(long rows) : d((size_t)rows * C), nr((int)rows) {} // VectorXd v(n)
    static Matrix Zero(long rows) { Matrix m(rows); std::fill(m.d.begin(), m.d.end(), T(0)); return m; }
    long size() const { return (long)d.size(); }
    T* data() { return d.data(); }
    const T* data() const { return d.data(); }
    template <int N> struct SegRef {
        Matrix& m; int s;
        template <class O> SegRef& operator+=(const O& o) { const Matrix<T, N, 1> v = o; for (int i = 0; i < N; ++i) m.d[s + i] += v.d[i]; return *this; }
        operator Matrix<T, N, 1>() const { Matrix<T, N, 1> r; for (int i = 0; i < N; ++i) r.d[i] = m.d[s + i]; return r; }
    };
    template <int N> SegRef<N> segment(int s) { return SegRef<N>{*this, s}; }
    T dot(const Matrix& o) const { T s = T(0); for (size_t i = 0; i < d.size(); ++i) s += d[i] * o.d[i]; return s; }
    T squaredNorm() const { return dot(*this); }
    Matrix operator-(const Matrix& o) const { Matrix r = *this; for (size_t i = 0; i < d.size(); ++i) r.d[i] -= o.d[i]; return r; }
    T& operator()(int i, int j) { return d[i + (size_t)j * nr]; }
    const T& operator()(int i, int j) const { return d[i + (size_t)j * nr]; }
    T& operator[](int i) { return d[i]; }
    const T& operator[](int i) const { return d[i]; }
    T& operator()(int i) { return d[i]; }
    const T& operator()(int i) const { return d[i]; }
    int rows() const { return nr; }
    int cols() const { return C; }
    Matrix<T, 1, C> row(int i) const { Matrix<T, 1, C> r; for (int j = 0; j < C; ++j) r.d[j] = (*this)(i, j); return r; }
    struct Colwise {
        const Matrix& m;
        Matrix<T, 1, C> minCoeff() const
        {
            Matrix<T, 1, C> r;
            for (int j = 0; j < C; ++j) { T v = m(0, j); for (int i = 1; i < m.nr; ++i) v = std::min(v, m(i, j)); r.d[j] = v; }
            return r;
        }
        Matrix<T, 1, C> maxCoeff() const
        {
            Matrix<T, 1, C> r;
            for (int j = 0; j < C; ++j) { T v = m(0, j); for (int i = 1; i < m.nr; ++i) v = std::max(v, m(i, j)); r.d[j] = v; }
            return r;
        }
    };
    Colwise colwise() const { return Colwise{*this}; }
    static T tree(const T* a, long n, long lo, long len)
    {
        if (lo >= n) return T(0);
        if (len == 1) return a[lo];
        return tree(a, n, lo, len / 2) + tree(a, n, lo + len / 2, len / 2);
    }
    T sum() const { const long n = (long)d.size(); if (n <= 0) return T(0); long P = 1; while (P < n) P <<= 1; return tree(d.data(), n, 0, P); }
    T mean() const { return sum() / T(d.size()); }
};

template <class T>
struct Triplet {
    int r, c; T v;
    Triplet() : r(0), c(0), v(0) {}
    Triplet(int i, int j, const T& x) : r(i), c(j), v(x) {}
    int row() const { return r; }
    int col() const { return c; }
    const T& value() const { return v; }
};

typedef Matrix<double, 1, 3> RowVector3d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, Dynamic, 1> VectorXd;
inline void setNbThreads(int) {}
typedef Matrix<double, 3, 3> Matrix3d;

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
// Symmetric eigen-decomposition the way Eigen does it: Householder reduction to tridiagonal form followed by implicit
// shifted QL/QR iterations on the tridiagonal matrix with the rotations accumulated (classic tred2 / tql2 structure),
// eigenvalues ascending, lower triangle read. (A first version used cyclic Jacobi; it was 5x slower than this, which made
// the reference loops compiled against this subset an unfairly slow CPU baseline.)
template <class T, int N>
struct SelfAdjointEigenSolver<Matrix<T, N, N>> {
    Matrix<T, N, 1> lam;
    Matrix<T, N, N> vec;
    explicit SelfAdjointEigenSolver(const Matrix<T, N, N>& Ain)
    {
        T V[N][N], d[N], e[N];
        for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) V[i][j] = (i >= j) ? Ain(i, j) : Ain(j, i);
        // ---- Householder tridiagonalisation (rows from the bottom up), V accumulates the transformation
        for (int j = 0; j < N; ++j) d[j] = V[N - 1][j];
        for (int i = N - 1; i > 0; --i) {
            T scale = 0, h = 0;
            for (int k = 0; k < i; ++k) scale += std::fabs(d[k]);
            if (scale == T(0)) {
                e[i] = d[i - 1];
                for (int j = 0; j < i; ++j) { d[j] = V[i - 1][j]; V[i][j] = 0; V[j][i] = 0; }
            }
            else {
                for (int k = 0; k < i; ++k) { d[k] /= scale; h += d[k] * d[k]; }
                T f = d[i - 1], g = std::sqrt(h);
                if (f > 0) g = -g;
                e[i] = scale * g;
                h -= f * g;
                d[i - 1] = f - g;
                for (int j = 0; j < i; ++j) e[j] = 0;
                for (int j = 0; j < i; ++j) {
                    f = d[j];
                    V[j][i] = f;
                    g = e[j] + V[j][j] * f;
                    for (int k = j + 1; k <= i - 1; ++k) { g += V[k][j] * d[k]; e[k] += V[k][j] * f; }
                    e[j] = g;
                }
                f = 0;
                for (int j = 0; j < i; ++j) { e[j] /= h; f += e[j] * d[j]; }
                const T hh = f / (h + h);
                for (int j = 0; j < i; ++j) e[j] -= hh * d[j];
                for (int j = 0; j < i; ++j) {
                    f = d[j]; g = e[j];
                    for (int k = j; k <= i - 1; ++k) V[k][j] -= (f * e[k] + g * d[k]);
                    d[j] = V[i - 1][j];
                    V[i][j] = 0;
                }
            }
            d[i] = h;
        }
        for (int i = 0; i < N - 1; ++i) {
            V[N - 1][i] = V[i][i];
            V[i][i] = 1;
            const T h = d[i + 1];
            if (h != T(0)) {
                for (int k = 0; k <= i; ++k) d[k] = V[k][i + 1] / h;
                for (int j = 0; j <= i; ++j) {
                    T g = 0;
                    for (int k = 0; k <= i; ++k) g += V[k][i + 1] * V[k][j];
                    for (int k = 0; k <= i; ++k) V[k][j] -= g * d[k];
                }
            }
            for (int k = 0; k <= i; ++k) V[k][i + 1] = 0;
        }
        for (int j = 0; j < N; ++j) { d[j] = V[N - 1][j]; V[N - 1][j] = 0; }
        V[N - 1][N - 1] = 1;
        e[0] = 0;
        // ---- implicit QL on the tridiagonal matrix
        for (int i = 1; i < N; ++i) e[i - 1] = e[i];
        e[N - 1] = 0;
        T f = 0, tst1 = 0;
        const T eps = std::numeric_limits<T>::epsilon();
        for (int l = 0; l < N; ++l) {
            tst1 = std::max(tst1, std::fabs(d[l]) + std::fabs(e[l]));
            int m = l;
            while (m < N) { if (std::fabs(e[m]) <= eps * tst1) break; ++m; }
            if (m > l) {
                int iter = 0;
                do {
                    ++iter;
                    T g = d[l];
                    T p = (d[l + 1] - g) / (2 * e[l]);
                    T r = std::hypot(p, T(1));
                    if (p < 0) r = -r;
                    d[l] = e[l] / (p + r);
                    d[l + 1] = e[l] * (p + r);
                    const T dl1 = d[l + 1];
                    T h = g - d[l];
                    for (int i = l + 2; i < N; ++i) d[i] -= h;
                    f += h;
                    p = d[m];
                    T c = 1, c2 = c, c3 = c, s = 0, s2 = 0;
                    const T el1 = e[l + 1];
                    for (int i = m - 1; i >= l; --i) {
                        c3 = c2; c2 = c; s2 = s;
                        g = c * e[i];
                        h = c * p;
                        r = std::hypot(p, e[i]);
                        e[i + 1] = s * r;
                        s = e[i] / r;
                        c = p / r;
                        p = c * d[i] - s * g;
                        d[i + 1] = h + s * (c * g + s * d[i]);
                        for (int k = 0; k < N; ++k) { h = V[k][i + 1]; V[k][i + 1] = s * V[k][i] + c * h; V[k][i] = c * V[k][i] - s * h; }
                    }
                    p = -s * s2 * c3 * el1 * e[l] / dl1;
                    e[l] = s * p;
                    d[l] = c * p;
                } while (std::fabs(e[l]) > eps * tst1 && iter < 60);
            }
            d[l] += f;
            e[l] = 0;
        }
        int idx[N];
        for (int i = 0; i < N; ++i) idx[i] = i;
        std::sort(idx, idx + N, [&](int a, int b) { return d[a] < d[b]; });
        for (int j = 0; j < N; ++j) { lam.d[j] = d[idx[j]]; for (int k = 0; k < N; ++k) vec(k, j) = V[k][idx[j]]; }
    }
    const Matrix<T, N, 1>& eigenvalues() const { return lam; }
    const Matrix<T, N, N>& eigenvectors() const { return vec; }
};

// ---- Eigen::SparseMatrix<T, RowMajor>: the subset Math/CSR_MATRIX.h and `sysMtr += M` (INC_POTENTIAL.h:383-386) use --------
// setFromTriplets follows Eigen 3.3's set_from_triplets: duplicates are summed in triplet order, inner indices end up
// ascending, explicit zeros stay; operator+= is the sparse sum (union of the patterns).
enum { ColMajor = 0, RowMajor = 1 };
template <class T, int Options = ColMajor>
class SparseMatrix {
    int nr = 0, nc = 0;
    std::vector<int> outer, inner;
    std::vector<T> vals;
public:
    void resize(int r, int c) { nr = r; nc = c; outer.assign((size_t)r + 1, 0); inner.clear(); vals.clear(); }
    void setZero() { std::fill(outer.begin(), outer.end(), 0); inner.clear(); vals.clear(); }
    void reserve(size_t n) { inner.resize(n); vals.resize(n); }
    void finalize() {}
    int rows() const { return nr; }
    int cols() const { return nc; }
    long nonZeros() const { return (long)outer[nr]; }
    int outerSize() const { return nr; }
    T* valuePtr() { return vals.data(); }
    int* innerIndexPtr() { return inner.data(); }
    int* outerIndexPtr() { return outer.data(); }
    template <class It>
    void setFromTriplets(It begin, It end)
    {
        std::vector<std::vector<std::pair<int, T>>> perRow((size_t)nr);
        for (It it = begin; it != end; ++it) perRow[it->row()].push_back({it->col(), it->value()});
        inner.clear(); vals.clear();
        outer.assign((size_t)nr + 1, 0);
        for (int r = 0; r < nr; ++r) {
            auto& e = perRow[r];
            std::stable_sort(e.begin(), e.end(), [](const std::pair<int, T>& a, const std::pair<int, T>& b) { return a.first < b.first; });
            for (size_t k = 0; k < e.size(); ++k) {
                if (k && e[k].first == e[k - 1].first) vals.back() += e[k].second;
                else { inner.push_back(e[k].first); vals.push_back(e[k].second); }
            }
            outer[r + 1] = (int)inner.size();
        }
    }
    T& coeffRef(int i, int j)
    {
        for (int p = outer[i]; p < outer[i + 1]; ++p) if (inner[p] == j) return vals[p];
        static T zero; zero = T(0); return zero; // (insertion is not needed by the compiled path)
    }
    SparseMatrix& operator+=(const SparseMatrix& o)
    {
        std::vector<int> no((size_t)nr + 1, 0), ni; std::vector<T> nv;
        for (int r = 0; r < nr; ++r) {
            int p = outer[r], q = o.outer[r];
            while (p < outer[r + 1] || q < o.outer[r + 1]) {
                const int cp = p < outer[r + 1] ? inner[p] : 2147483647, cq = q < o.outer[r + 1] ? o.inner[q] : 2147483647;
                if (cp == cq) { ni.push_back(cp); nv.push_back(vals[p] + o.vals[q]); ++p; ++q; }
                else if (cp < cq) { ni.push_back(cp); nv.push_back(vals[p]); ++p; }
                else { ni.push_back(cq); nv.push_back(o.vals[q]); ++q; }
            }
            no[r + 1] = (int)ni.size();
        }
        outer.swap(no); inner.swap(ni); vals.swap(nv);
        return *this;
    }
    Matrix<T, Dynamic, 1> operator*(const Matrix<T, Dynamic, 1>& x) const // row-major product (CSR_MATRIX keeps RowMajor matrices)
    {
        Matrix<T, Dynamic, 1> y = Matrix<T, Dynamic, 1>::Zero(nr);
        for (int r = 0; r < nr; ++r) { T s = T(0); for (int p = outer[r]; p < outer[r + 1]; ++p) s += vals[p] * x[inner[p]]; y[r] = s; }
        return y;
    }
    T coeff(int i, int j) const { for (int p = outer[i]; p < outer[i + 1]; ++p) if (inner[p] == j) return vals[p]; return T(0); }
    class InnerIterator {
        SparseMatrix& m; int k, p;
    public:
        InnerIterator(SparseMatrix& mat, int outerI) : m(mat), k(outerI), p(mat.outer[outerI]) {}
        operator bool() const { return p < m.outer[k + 1]; }
        InnerIterator& operator++() { ++p; return *this; }
        int row() const { return k; }
        int col() const { return m.inner[p]; }
        const T& value() const { return m.vals[p]; }
        T& valueRef() { return m.vals[p]; }
    };
};

} // namespace Eigen
