// ORACLE / TEST INFRASTRUCTURE: inert stand-in for <pybind11/stl.h> (nothing is bound in oracle/_ref).
#pragma once
