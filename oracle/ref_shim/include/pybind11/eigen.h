// ORACLE / TEST INFRASTRUCTURE: shadows <pybind11/eigen.h>; provides the small Eigen subset the reference's
// Math/Distance, BARRIER.h and UTILS.h headers use (Eigen itself is not installed in this image).
#pragma once
#include "../eigen_shim.hpp"
