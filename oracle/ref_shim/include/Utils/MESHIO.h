// ORACLE / TEST INFRASTRUCTURE: shadows the reference's <Utils/MESHIO.h> (mesh file I/O on top of the Cabana / Kokkos
// storages, neither of which exists in this image) with the minimum the contact path needs from it: the reference's own
// Math/VECTOR.h, and a plain-std::vector stand-in for the BASE_STORAGE API that FEM/IPC.h and Grid/SPATIAL_HASH.h use
// (Get_Unchecked -> tuple of references, size, Append, Reserve, Par_Each(id, tuple), FIELDS<...>::x0/v/g/m). Par_Each is an
// OpenMP parallel for when compiled with -fopenmp -- the reference runs it as a Kokkos::OpenMP parallel_for, so its bodies
// are thread safe by construction -- and a serial loop otherwise.
#pragma once
#include <Math/VECTOR.h>
#include <cstdio>
#include <complex>
#include <deque>
#include <list>
#include <queue>
#include <cassert>
#include <numeric>
#include <algorithm>
#include <iostream>
#include <string>
#include <cstdlib>
#include <map>
#include <set>
#include <tuple>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#define TIMER_FLAG(name) do { } while (0)

namespace JGSL {

template <class... Ts>
struct BASE_STORAGE {
    std::vector<std::tuple<Ts...>> rows;
    int size = 0;
    BASE_STORAGE() {}
    explicit BASE_STORAGE(std::size_t reserve) { rows.reserve(reserve); }
    void Reserve(std::size_t n) { rows.reserve(n); }
    int Append(const Ts&... v) { rows.emplace_back(v...); return size++; }
    std::tuple<Ts&...> Get_Unchecked(int i) { return ref_tuple(rows[i], std::index_sequence_for<Ts...>()); }
    template <class F> void Par_Each(F f)
    {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 64)
#endif
        for (int i = 0; i < size; ++i) f(i, Get_Unchecked(i));
    }
    template <class F> void Each(F f) { for (int i = 0; i < size; ++i) f(i, Get_Unchecked(i)); }
    void deep_copy_to(BASE_STORAGE& o) const { o.rows = rows; o.size = size; }
    template <std::size_t I, class V> void Fill(const V& v) { for (auto& r : rows) std::get<I>(r) = v; }
    int Insert(int i, const Ts&... v) { if (i >= size) { rows.resize((std::size_t)i + 1); size = i + 1; } rows[i] = std::tuple<Ts...>(v...); return i; }
    // Join(other).Par_Each(f): f(id, tuple of references to this row's fields followed by the other storage's)
    template <class... Us> struct JOINED {
        BASE_STORAGE& a; BASE_STORAGE<Us...>& b;
        template <class F> void Par_Each(F f)
        {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 64)
#endif
            for (int i = 0; i < a.size; ++i) f(i, std::tuple_cat(a.Get_Unchecked(i), b.Get_Unchecked(i)));
        }
        template <class F> void Each(F f) { for (int i = 0; i < a.size; ++i) f(i, std::tuple_cat(a.Get_Unchecked(i), b.Get_Unchecked(i))); }
    };
    template <class... Us> JOINED<Us...> Join(BASE_STORAGE<Us...>& o) { return JOINED<Us...>{*this, o}; }
private:
    template <std::size_t... I> static std::tuple<Ts&...> ref_tuple(std::tuple<Ts...>& t, std::index_sequence<I...>) { return std::tuple<Ts&...>(std::get<I>(t)...); }
};

template <std::size_t OFFSET, class S> struct FIELDS_WITH_OFFSET;
template <class S> using FIELDS = FIELDS_WITH_OFFSET<0, S>;

template <class T, int dim> using MESH_NODE = BASE_STORAGE<VECTOR<T, dim>>;
template <int dim> using MESH_ELEM = BASE_STORAGE<VECTOR<int, dim + 1>>;
template <class T, int dim> using MESH_NODE_ATTR = BASE_STORAGE<VECTOR<T, dim>, VECTOR<T, dim>, VECTOR<T, dim>, T>;
template <std::size_t OFFSET, class T, int dim>
struct FIELDS_WITH_OFFSET<OFFSET, MESH_NODE_ATTR<T, dim>> {
    enum INDICES { x0 = OFFSET, v, g, m };
};

} // namespace JGSL
