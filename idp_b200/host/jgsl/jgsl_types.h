// Host-side value and container types behind the Python-visible `JGSL` module of this repository.
// The reference keeps these in Cabana AoSoA storages (Library/Storage/storage.hpp) and Eigen sparse matrices
// (Library/Math/CSR_MATRIX.h); neither dependency exists here and the hot path lives on the device, so the host side only
// needs plain row containers that the drivers (Python/Drivers/FEMDiscreteShellBase.py) create and pass back in.
#pragma once
#include <array>
#include <cmath>
#include <cstddef>
#include <map>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

namespace jgsl {

template <class T, int d>
struct Vec {
    T data[d];
    Vec() { for (int i = 0; i < d; ++i) data[i] = T(0); }
    explicit Vec(T a) { for (int i = 0; i < d; ++i) data[i] = a; }
    Vec(T a, T b) { static_assert(d == 2, "2 components"); data[0] = a; data[1] = b; }
    Vec(T a, T b, T c) { static_assert(d == 3, "3 components"); data[0] = a; data[1] = b; data[2] = c; }
    Vec(T a, T b, T c, T e) { static_assert(d == 4, "4 components"); data[0] = a; data[1] = b; data[2] = c; data[3] = e; }
    T& operator[](int i) { return data[i]; }
    const T& operator[](int i) const { return data[i]; }
    Vec operator+(const Vec& o) const { Vec r; for (int i = 0; i < d; ++i) r.data[i] = data[i] + o.data[i]; return r; }
    Vec operator-(const Vec& o) const { Vec r; for (int i = 0; i < d; ++i) r.data[i] = data[i] - o.data[i]; return r; }
    Vec operator*(T s) const { Vec r; for (int i = 0; i < d; ++i) r.data[i] = data[i] * s; return r; }
    Vec operator/(T s) const { Vec r; for (int i = 0; i < d; ++i) r.data[i] = data[i] / s; return r; }
    Vec& operator+=(const Vec& o) { for (int i = 0; i < d; ++i) data[i] += o.data[i]; return *this; }
    Vec& operator-=(const Vec& o) { for (int i = 0; i < d; ++i) data[i] -= o.data[i]; return *this; }
    Vec& operator*=(T s) { for (int i = 0; i < d; ++i) data[i] *= s; return *this; }
    Vec& operator/=(T s) { for (int i = 0; i < d; ++i) data[i] /= s; return *this; }
    bool operator==(const Vec& o) const { for (int i = 0; i < d; ++i) if (data[i] != o.data[i]) return false; return true; }
    T dot(const Vec& o) const { T s = T(0); for (int i = 0; i < d; ++i) s += data[i] * o.data[i]; return s; }
    T length2() const { return dot(*this); }
    T length() const { return (T)std::sqrt((double)length2()); }
    Vec normalized() const { return *this / length(); }
    T sum() const { T s = T(0); for (int i = 0; i < d; ++i) s += data[i]; return s; }
    T average() const { return sum() / T(d); }
};

template <class T>
inline Vec<T, 3> cross(const Vec<T, 3>& a, const Vec<T, 3>& b)
{
    return Vec<T, 3>(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}

template <class T, int d>
struct Mat { // column major like the reference's MATRIX<T,dim>; m(i, j) = row i, column j
    T a[d * d];
    Mat() { for (int i = 0; i < d * d; ++i) a[i] = T(0); }
    explicit Mat(T diag) { for (int i = 0; i < d * d; ++i) a[i] = T(0); for (int i = 0; i < d; ++i) a[i * d + i] = diag; }
    T& operator()(int i, int j) { return a[j * d + i]; }
    const T& operator()(int i, int j) const { return a[j * d + i]; }
};

template <class T>
struct Scalar { T value = T(0); };

// rows of tuples: what BASE_STORAGE<Fields...> is to the drivers (construct, pass around, read size)
template <class... F>
struct Storage {
    typedef std::tuple<F...> Row;
    std::vector<Row> rows;
    int size() const { return (int)rows.size(); }
    void append(const F&... f) { rows.emplace_back(f...); }
    void clear() { rows.clear(); }
};

typedef Storage<Vec<double, 3>> NodeStorage;                                            // MESH_NODE<double,3>
typedef Storage<Vec<int, 3>> TriStorage;                                                // MESH_ELEM<2>
typedef Storage<Vec<double, 3>, Vec<double, 3>, Vec<double, 3>, double> NodeAttrStorage; // x0, v, g, m
typedef Storage<Mat<double, 2>, Mat<double, 2>> ElemAttrStorage;                        // IB, D (P(0,0) of element 0 = hinge k)
typedef Storage<Mat<double, 2>, double, double, double> Fcr2Storage;                    // F, vol, lambda, mu
typedef Storage<Mat<double, 3>, double, double, double> Fcr3Storage;
typedef Storage<Vec<double, 4>> DbcStorage;                                             // (vertex, target xyz)
typedef Storage<Vec<int, 2>, Vec<double, 3>, Vec<double, 3>, Vec<double, 3>, double> DbcMotionStorage;

// The mass matrix the drivers carry around (CSR_MATRIX<double>); the shell path only ever builds a diagonal one
// (Library/FEM/Shell/DISCRETE_SHELL.h:279-318), kept as CSR arrays so that Construct_From_CSR-style consumers can read it.
struct CsrMatrix {
    int n = 0;
    std::vector<int> ptr, col;
    std::vector<double> val;
    void set_diagonal(const std::vector<double>& d)
    {
        n = (int)d.size();
        ptr.resize(n + 1); col.resize(n); val = d;
        for (int i = 0; i < n; ++i) { ptr[i] = i; col[i] = i; }
        ptr[n] = n;
    }
    double coeff(int r, int c) const
    {
        for (int k = ptr[r]; k < ptr[r + 1]; ++k) if (col[k] == c) return val[k];
        return 0.0;
    }
    double diagonal_mean() const
    {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += coeff(i, i);
        return n ? s / n : 0.0;
    }
};

typedef std::map<std::pair<int, int>, int> EdgeToTri;

} // namespace jgsl
