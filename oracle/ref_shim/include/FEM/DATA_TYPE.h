// ORACLE / TEST INFRASTRUCTURE: shadows the reference's <FEM/DATA_TYPE.h> (aliases on top of the Cabana storage) with the same
// aliases on the std::vector stand-in of include/Utils/MESHIO.h, so that the shell energy headers (FEM/Shell/MEMBRANE.h ...)
// compile from where they lie.
#pragma once
#include <Utils/MESHIO.h>

namespace JGSL {

template <class T, int dim> using MESH_ELEM_ATTR = BASE_STORAGE<MATRIX<T, dim>, MATRIX<T, dim>>; // IB, P
template <std::size_t OFFSET, class T, int dim>
struct FIELDS_WITH_OFFSET<OFFSET, MESH_ELEM_ATTR<T, dim>> {
    enum INDICES { IB = OFFSET, P };
};
template <class T> using SCALAR_STORAGE = BASE_STORAGE<T>;
template <class T, int dim> using VECTOR_STORAGE = BASE_STORAGE<VECTOR<T, dim>>;

} // namespace JGSL
