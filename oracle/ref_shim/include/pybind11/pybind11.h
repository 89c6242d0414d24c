// ORACLE / TEST INFRASTRUCTURE: stub that shadows <pybind11/pybind11.h> when the reference's headers
// (/root/reference/Library/...) are compiled into oracle/_ref. Nothing is bound: the registration helpers the reference
// defines inline (e.g. Math/VECTOR.h Export_Vector) only have to parse and compile, they are never called.
#pragma once
#include <utility>
namespace pybind11 {
struct module {
    template <class... A> module& def(A&&...) { return *this; }
    template <class... A> module def_submodule(A&&...) { return *this; }
};
using module_ = module;
template <class... Ts> struct init {};
enum class return_value_policy { automatic, reference, reference_internal, copy, move, take_ownership };
struct is_operator {};
struct self_t {};
static const self_t self = self_t();
struct op_ {};
template <class B> inline op_ operator+(const self_t&, const B&) { return op_(); }
template <class B> inline op_ operator-(const self_t&, const B&) { return op_(); }
template <class B> inline op_ operator*(const self_t&, const B&) { return op_(); }
template <class B> inline op_ operator/(const self_t&, const B&) { return op_(); }
template <class B> inline op_ operator+=(const self_t&, const B&) { return op_(); }
template <class B> inline op_ operator-=(const self_t&, const B&) { return op_(); }
template <class B> inline op_ operator*=(const self_t&, const B&) { return op_(); }
template <class B> inline op_ operator/=(const self_t&, const B&) { return op_(); }
template <class T, class... Opts> struct class_ {
    template <class... A> class_(A&&...) {}
    template <class... A> class_& def(A&&...) { return *this; }
    template <class... A> class_& def_readwrite(A&&...) { return *this; }
    template <class... A> class_& def_property(A&&...) { return *this; }
    template <class... A> class_& def_static(A&&...) { return *this; }
};
} // namespace pybind11
