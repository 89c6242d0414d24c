// ORACLE / TEST INFRASTRUCTURE: shadows <pybind11/operators.h>; see pybind11.h in this directory.
#pragma once
#include "pybind11.h"
