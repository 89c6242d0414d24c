// ORACLE / TEST INFRASTRUCTURE: shadows the reference's <Utils/PROFILER.h> (timers, logging, pybind11 iostream) when
// Math/CSR_MATRIX.h is compiled into oracle/_ref: TIMER_FLAG becomes a no-op, `py` is the inert pybind11 stand-in.
#pragma once
#include <pybind11/pybind11.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <tuple>
#include <vector>
namespace py = pybind11;
#ifndef TIMER_FLAG
#define TIMER_FLAG(name) do { } while (0)
#endif
