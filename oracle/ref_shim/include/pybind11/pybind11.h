// ORACLE / TEST INFRASTRUCTURE: stub that shadows <pybind11/pybind11.h> when the reference's header-only math
// (/root/reference/Library/Math/...) is compiled into oracle/_ref. The math headers only need the namespace to exist.
#pragma once
namespace pybind11 {}
