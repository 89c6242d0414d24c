// Test-only stand-ins for the reference's container types (Library/Math/VECTOR.h:33-44, Library/Storage/storage.hpp,
// Library/FEM/DATA_TYPE.h:6-18): same member names and access pattern (size, Get_Unchecked(i) -> tuple of references,
// 4-lane padded VECTOR), so that idp_b200/host/IPC_B200.h is compiled and exercised here without Cabana/Kokkos/Eigen.
#pragma once
#include <tuple>
#include <vector>
namespace JGSL {
template <class T, int dim>
struct alignas(sizeof(T) * 4) VECTOR {
    T data[4];
    VECTOR() : data{0, 0, 0, 0} {}
    VECTOR(T x, T y) : data{x, y, 0, 0} {}
    VECTOR(T x, T y, T z) : data{x, y, z, 0} {}
    VECTOR(T x, T y, T z, T w) : data{x, y, z, w} {}
    T& operator[](int i) { return data[i]; }
    const T& operator[](int i) const { return data[i]; }
};
template <class... Ts>
struct BASE_STORAGE {
    std::vector<std::tuple<Ts...>> rows;
    std::size_t size = 0;
    void Append(const Ts&... v) { rows.emplace_back(v...); size = rows.size(); }
    std::tuple<Ts&...> Get_Unchecked(std::size_t i) { return std::apply([](Ts&... a) { return std::tuple<Ts&...>(a...); }, rows[i]); }
};
template <class T, int dim> using MESH_NODE = BASE_STORAGE<VECTOR<T, dim>>;
template <class T, int dim> using MESH_NODE_ATTR = BASE_STORAGE<VECTOR<T, dim>, VECTOR<T, dim>, VECTOR<T, dim>, T>; // x0, v, g, m
} // namespace JGSL
namespace Eigen {
template <class T> struct Triplet {
    int r, c; T v;
    Triplet() : r(0), c(0), v(0) {}
    Triplet(int r_, int c_, T v_) : r(r_), c(c_), v(v_) {}
    int row() const { return r; } int col() const { return c; } T value() const { return v; }
};
} // namespace Eigen
