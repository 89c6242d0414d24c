// ORACLE / TEST INFRASTRUCTURE: the three metafunctions of range-v3's meta library that the reference's
// Physics/FIXED_COROTATED.h uses (meta::if_, meta::equal_to, meta::int_); the library itself is not in this image.
#pragma once
#include <type_traits>
namespace meta {
template <int N> using int_ = std::integral_constant<int, N>;
template <class A, class B> using equal_to = std::integral_constant<bool, (A::value == B::value)>;
template <class C, class T, class F> using if_ = typename std::conditional<C::value, T, F>::type;
} // namespace meta
