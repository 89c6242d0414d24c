// Python module `JGSL` (pybind11) hosting the device-resident contact path: the subset of the reference's module surface
// (Library/EXPORTER.cpp:17-36) that Projects/FEMShell/12-14_normal_flow.py and 16_fix_char_seq.py ->
// Python/Drivers/{SimulationBase,FEMDiscreteShellBase}.py touch, with the same names, argument order and side effects, so
// that those scripts run unchanged with this module on their import path:
//   module level   Kokkos_Initialize, Set_Parameter / Get_Parameter, TIMER_FLUSH, Scalar*/Vector*/Matrix*, StdVector*,
//                  StdMapPairiToi, CSR_MATRIX_D, FIXED_COROTATED_{2,3}.Create
//   Storage.*      the storages the drivers construct
//   MeshIO.*       Append_Attribute, Read_TriMesh_Obj, Write_TriMesh_Obj, Load_Velocity_X0, Zero_Velocity
//   FEM.*          Boundary_Dirichlet, Init/Step/Turn/Reset/Load_Dirichlet;  FEM.DiscreteShell.*  Add_Shell,
//                  Initialize_Shell_Hinge_EIPC, Initialize_OIPC, Update_Normal_Flow_Neumann, Advance_One_Step_IE_Flow,
//                  Advance_One_Step_IE_Hinge  (DISCRETE_SHELL.h:1087-1126)
// Everything else of the reference's module raises NotImplementedError by name (B200_NOT_BUILT) instead of being absent
// silently. The time step itself is shell_flow.h on the backend selected at compile time: backend_b200.h (the product) or,
// for the test-only trace checker built under tests/host_shim/, the reference's own CPU loops.
#include <pybind11/operators.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>
#include <pybind11/stl_bind.h>

#include <chrono>
#include <memory>
#include <variant>

#ifndef JGSL_BACKEND_HEADER
#define JGSL_BACKEND_HEADER "backend_b200.h"
#define JGSL_BACKEND_CLASS jgsl::B200Backend
#endif
#include JGSL_BACKEND_HEADER
#include "dirichlet.h"
#ifdef JGSL_STEP_HOOK_HEADER
#include JGSL_STEP_HOOK_HEADER // test builds may take over a whole time step (tests/host_shim/ref_driver_hook.h); never set for the product
#endif

namespace py = pybind11;
using namespace jgsl;

typedef std::vector<double> StdVectorXd;
typedef std::vector<int> StdVectorXi;
typedef std::vector<Vec<int, 2>> StdVectorVector2i;
typedef std::vector<Vec<int, 3>> StdVectorVector3i;
typedef std::vector<Vec<int, 4>> StdVectorVector4i;
typedef std::vector<Vec<double, 3>> StdVectorVector3d;
PYBIND11_MAKE_OPAQUE(StdVectorXd);
PYBIND11_MAKE_OPAQUE(StdVectorXi);
PYBIND11_MAKE_OPAQUE(StdVectorVector2i);
PYBIND11_MAKE_OPAQUE(StdVectorVector3i);
PYBIND11_MAKE_OPAQUE(StdVectorVector4i);
PYBIND11_MAKE_OPAQUE(StdVectorVector3d);
PYBIND11_MAKE_OPAQUE(EdgeToTri);

namespace {

typedef std::variant<bool, int, double, std::string> ParamValue;
std::map<std::string, ParamValue>& params()
{
    static std::map<std::string, ParamValue> p;
    return p;
}
template <class T>
void set_param(const std::string& k, const T& v) { params()[k] = v; }
template <class T>
T get_param(const std::string& k, const T& dflt)
{
    const auto it = params().find(k);
    if (it == params().end()) return dflt;
    if (const T* p = std::get_if<T>(&it->second)) return *p;
    return dflt;
}

std::unique_ptr<JGSL_BACKEND_CLASS>& backend_slot()
{
    static std::unique_ptr<JGSL_BACKEND_CLASS> b;
    return b;
}
JGSL_BACKEND_CLASS& backend()
{
    auto& b = backend_slot();
    if (!b) {
        const char* dev = std::getenv("IDP_DEVICE");
        b.reset(new JGSL_BACKEND_CLASS(dev ? std::atoi(dev) : 0));
    }
    return *b;
}

template <class T, int d, class Cls>
void register_vector_ops(Cls& cls)
{
    cls.def(py::init<>())
        .def(py::init<T>())
        .def(py::self + py::self)
        .def(py::self - py::self)
        .def(py::self * T())
        .def(py::self / T())
        .def(py::self += py::self)
        .def(py::self -= py::self)
        .def(py::self *= T())
        .def(py::self /= T())
        .def("__eq__", [](const Vec<T, d>& a, const Vec<T, d>& b) { return a == b; }, py::is_operator())
        .def("__getitem__", [](const Vec<T, d>& v, int i) {
            if (i < 0 || i >= d) throw py::index_error();
            return v.data[i];
        })
        .def("__len__", [](const Vec<T, d>&) { return d; })
        .def("dot", &Vec<T, d>::dot)
        .def("length", &Vec<T, d>::length)
        .def("length2", &Vec<T, d>::length2)
        .def("normalized", &Vec<T, d>::normalized)
        .def("sum", &Vec<T, d>::sum)
        .def("average", &Vec<T, d>::average);
}

template <class T>
void register_vectors(py::module_& m, const char* suffix)
{
    const std::string s(suffix);
    auto v2 = py::class_<Vec<T, 2>>(m, ("Vector2" + s).c_str()).def(py::init<T, T>());
    register_vector_ops<T, 2>(v2);
    auto v3 = py::class_<Vec<T, 3>>(m, ("Vector3" + s).c_str()).def(py::init<T, T, T>());
    register_vector_ops<T, 3>(v3);
    auto v4 = py::class_<Vec<T, 4>>(m, ("Vector4" + s).c_str()).def(py::init<T, T, T, T>());
    register_vector_ops<T, 4>(v4);
}

template <class T>
void register_matrices(py::module_& m, const char* suffix)
{
    const std::string s(suffix);
    py::class_<Mat<T, 2>>(m, ("Matrix2" + s).c_str()).def(py::init<>()).def(py::init<T>())
        .def("__getitem__", [](const Mat<T, 2>& a, std::tuple<int, int> ij) { return a(std::get<0>(ij), std::get<1>(ij)); });
    py::class_<Mat<T, 3>>(m, ("Matrix3" + s).c_str()).def(py::init<>()).def(py::init<T>())
        .def("__getitem__", [](const Mat<T, 3>& a, std::tuple<int, int> ij) { return a(std::get<0>(ij), std::get<1>(ij)); });
    py::class_<Scalar<T>>(m, ("Scalar" + s).c_str()).def(py::init<>());
}

template <class S>
py::class_<S> register_storage(py::module_& m, const char* name)
{
    return py::class_<S>(m, name).def(py::init<>()).def("__len__", [](const S& s) { return s.size(); }).def_property_readonly("size", &S::size);
}

void not_built(py::module_& m, const char* name)
{
    const std::string n(name);
    m.def(name, [n](py::args, py::kwargs) {
        PyErr_SetString(PyExc_NotImplementedError,
            ("JGSL." + n + " is outside the B200 contact-path build (only the normal-flow shell path is hosted; see DESIGN.md section 7)").c_str());
        throw py::error_already_set();
    });
}

} // namespace

PYBIND11_MODULE(JGSL, m)
{
    m.doc() = "JGSL-compatible module hosting the B200-resident IPC contact path (flow shell time step)";

    // ---- module level (EXPORTER.cpp:17-36, Utils/PARAMETER.h:21-32, Utils/PROFILER.h:196-201) --------------------------------
    m.def("Kokkos_Initialize", []() {}); // no host execution space to start: the parallel work is on the device
    m.def("Set_Parameter", &set_param<bool>);
    m.def("Set_Parameter", &set_param<int>);
    m.def("Set_Parameter", &set_param<double>);
    m.def("Set_Parameter", &set_param<std::string>);
    m.def("Get_Parameter", &get_param<bool>);
    m.def("Get_Parameter", &get_param<int>);
    m.def("Get_Parameter", &get_param<double>);
    m.def("Get_Parameter", &get_param<std::string>);
    m.def("TIMER_FLUSH", [](int frame, int frameNum, double t, double frameDt) {
        static auto last = std::chrono::steady_clock::now();
        const auto now = std::chrono::steady_clock::now();
        printf("[frame %d/%d] %.6g of %.6g, %.3f s since the previous flush", frame, frameNum, t, frameDt, std::chrono::duration<double>(now - last).count());
        last = now;
        if (backend_slot()) printf("  (%s backend)", backend_slot()->name());
        printf("\n");
        fflush(stdout);
    });

    register_vectors<float>(m, "f");
    register_vectors<double>(m, "d");
    register_matrices<float>(m, "f");
    register_matrices<double>(m, "d");
    py::class_<Vec<int, 2>>(m, "Vector2i").def(py::init<>()).def(py::init<int>()).def(py::init<int, int>())
        .def("__getitem__", [](const Vec<int, 2>& v, int i) { return v.data[i]; });
    py::class_<Vec<int, 3>>(m, "Vector3i").def(py::init<>()).def(py::init<int>()).def(py::init<int, int, int>())
        .def("__getitem__", [](const Vec<int, 3>& v, int i) { return v.data[i]; });
    py::class_<Vec<int, 4>>(m, "Vector4i").def(py::init<>()).def(py::init<int, int, int, int>())
        .def("__getitem__", [](const Vec<int, 4>& v, int i) { return v.data[i]; });
    py::class_<Scalar<int>>(m, "Scalari").def(py::init<>());

    py::bind_vector<StdVectorXd>(m, "StdVectorXd");
    py::bind_vector<StdVectorXi>(m, "StdVectorXi");
    py::bind_vector<StdVectorVector2i>(m, "StdVectorVector2i");
    py::bind_vector<StdVectorVector3i>(m, "StdVectorVector3i");
    py::bind_vector<StdVectorVector4i>(m, "StdVectorVector4i");
    py::bind_vector<StdVectorVector3d>(m, "StdVectorVector3d");
    py::class_<EdgeToTri>(m, "StdMapPairiToi").def(py::init<>()).def("__len__", [](const EdgeToTri& e) { return e.size(); });

    py::class_<CsrMatrix>(m, "CSR_MATRIX_D").def(py::init<>())
        .def("rows", [](const CsrMatrix& a) { return a.n; })
        .def("coeff", &CsrMatrix::coeff);

    // ---- Storage.* (FEM/FEM_EXPORTER.h:16-50) ------------------------------------------------------------------------------------
    py::module_ st = m.def_submodule("Storage", "storages the drivers construct and pass back in");
    register_storage<Storage<double>>(st, "SdStorage");
    register_storage<Storage<int>>(st, "SiStorage");
    register_storage<Storage<Vec<double, 2>>>(st, "V2dStorage");
    register_storage<NodeStorage>(st, "V3dStorage")
        .def("get", [](const NodeStorage& s, int i) { return std::get<0>(s.rows.at(i)); });
    register_storage<DbcStorage>(st, "V4dStorage")
        .def("get", [](const DbcStorage& s, int i) { return std::get<0>(s.rows.at(i)); });
    register_storage<Storage<Vec<int, 2>>>(st, "V2iStorage");
    register_storage<TriStorage>(st, "V3iStorage")
        .def("get", [](const TriStorage& s, int i) { return std::get<0>(s.rows.at(i)); });
    register_storage<Storage<Vec<int, 4>>>(st, "V4iStorage");
    register_storage<Storage<Vec<double, 2>, Vec<double, 2>, Vec<double, 2>, double>>(st, "V2dV2dV2dSdStorage");
    register_storage<NodeAttrStorage>(st, "V3dV3dV3dSdStorage");
    register_storage<ElemAttrStorage>(st, "M2dM2dSdStorage");
    register_storage<Storage<Mat<double, 3>, Mat<double, 3>>>(st, "M3dM3dSdStorage");
    register_storage<Storage<Vec<int, 2>, Vec<double, 2>, Vec<double, 2>, Vec<double, 2>, double>>(st, "V2iV2dV2dV2dSdStorage");
    register_storage<DbcMotionStorage>(st, "V2iV3dV3dV3dSdStorage");
    register_storage<Fcr2Storage>(st, "FCR2Storage");
    register_storage<Fcr3Storage>(st, "FCR3Storage");

    // FIXED_COROTATED_{2,3}.Create (Physics/CONSTITUTIVE_MODEL.h:56-69)
    py::module_ f2 = m.def_submodule("FIXED_COROTATED_2");
    f2.def("Create", []() { return std::unique_ptr<Fcr2Storage>(new Fcr2Storage()); });
    py::module_ f3 = m.def_submodule("FIXED_COROTATED_3");
    f3.def("Create", []() { return std::unique_ptr<Fcr3Storage>(new Fcr3Storage()); });
    not_built(f3, "All_Append_FEM");

    // ---- MeshIO.* (Utils/MESHIO.h:1323-1345) -----------------------------------------------------------------------------------
    py::module_ io = m.def_submodule("MeshIO", "mesh files");
    io.def("Append_Attribute", [](const NodeStorage& src, NodeStorage& dst) { dst.rows.insert(dst.rows.end(), src.rows.begin(), src.rows.end()); });
    io.def("Append_Attribute", [](const TriStorage& src, TriStorage& dst) { dst.rows.insert(dst.rows.end(), src.rows.begin(), src.rows.end()); });
    io.def("Read_TriMesh_Obj", &read_trimesh_obj, "read triangle mesh from obj file");
    io.def("Write_TriMesh_Obj", &write_trimesh_obj, "write triangle mesh to obj file");
    // Load_Velocity_X0 (Utils/MESHIO.h:986-1004): velocities that carry the current nodes to frame `lastFrame` within h
    io.def("Load_Velocity_X0", [](const std::string& folder, int lastFrame, double h, NodeStorage& Xcur, NodeAttrStorage& nodeAttr) {
        NodeStorage X1;
        TriStorage E1;
        read_trimesh_obj(folder + "/" + std::to_string(lastFrame) + ".obj", X1, E1);
        if (Xcur.size() != nodeAttr.size() || X1.size() != nodeAttr.size()) {
            printf("node count does not match!\n");
            exit(-1);
        }
        for (int i = 0; i < Xcur.size(); ++i) std::get<1>(nodeAttr.rows[i]) = (std::get<0>(X1.rows[i]) - std::get<0>(Xcur.rows[i])) / h;
    });
    io.def("Zero_Velocity", [](NodeAttrStorage& nodeAttr) { for (auto& r : nodeAttr.rows) std::get<1>(r) = Vec<double, 3>(); }); // :1006-1014
    for (const char* n : {"Transform_Points", "Read_SegMesh_Seg", "Write_SegMesh_Obj", "Read_TetMesh_Vtk", "Find_Surface_TriMesh",
             "Write_Surface_TriMesh_Obj", "Load_Velocity"})
        not_built(io, n);

    // ---- FEM.* (FEM/BOUNDARY_CONDITION.h:268-291) and FEM.DiscreteShell.* (FEM/Shell/DISCRETE_SHELL.h:1087-1126) ---------------------
    py::module_ fem = m.def_submodule("FEM", "finite elements");
    fem.def("Boundary_Dirichlet", &boundary_dirichlet);
    fem.def("Init_Dirichlet",
        [](NodeStorage& X, const Vec<double, 3>& lo, const Vec<double, 3>& hi, const Vec<double, 3>& v, const Vec<double, 3>& c, const Vec<double, 3>& axis,
            double angVelDeg, DbcStorage& DBC, DbcMotionStorage& motion, const Vec<int, 4>& vIndRange) {
            init_dirichlet(X, lo, hi, v, c, axis, angVelDeg, DBC, motion, vIndRange, get_param<double>("Dirichlet_ring", 0.0));
        },
        py::arg("X"), py::arg("relBoxMin"), py::arg("relBoxMax"), py::arg("v"), py::arg("rotCenter"), py::arg("rotAxis"), py::arg("angVelDeg"),
        py::arg("DBC"), py::arg("DBCMotion"), py::arg("vIndRange") = Vec<int, 4>(0, 0, INT_MAX, -1));
    fem.def("Step_Dirichlet", &step_dirichlet);
    fem.def("Turn_Dirichlet", &turn_dirichlet);
    fem.def("Reset_Dirichlet", &reset_dirichlet);
    fem.def("Load_Dirichlet", &load_dirichlet);
    for (const char* n : {"Pop_Back_Dirichlet", "Magnify_Body_Force", "Update_Inv_Basis", "Compute_Vol_And_Inv_Basis",
             "Compute_Mass_And_Init_Velocity_NoAlloc", "Augment_Mass_Matrix_And_Body_Force"})
        not_built(fem, n);

    py::module_ sh = fem.def_submodule("DiscreteShell", "discrete shell simulation hosted on the B200 contact path");
    sh.def("Add_Shell", &add_shell);
    auto init_hinge = [](bool eipc) {
        return [eipc](double rho0, double E, double nu, double thickness, double h, double dHat2, NodeStorage& X, TriStorage& Elem, StdVectorVector2i& seg,
                   EdgeToTri& edge2tri, StdVectorVector4i& edgeStencil, StdVectorVector3d& edgeInfo, NodeAttrStorage& nodeAttr, CsrMatrix& M,
                   const Vec<double, 3>& gravity, StdVectorXd& b, ElemAttrStorage& elemAttr, Fcr2Storage& elasticityAttr, Vec<double, 3>& kappa) {
            return initialize_shell_hinge(rho0, E, nu, thickness, h, dHat2, X, Elem, seg, edge2tri, edgeStencil, edgeInfo, nodeAttr, M, gravity, b, elemAttr,
                elasticityAttr, kappa, eipc);
        };
    };
    sh.def("Initialize_Shell_Hinge_EIPC", init_hinge(true));  // Initialize_Discrete_Shell<double, 3, KL=false, elasticIPC=true>
    sh.def("Initialize_Shell_Hinge", init_hinge(false));      // ... elasticIPC=false
    sh.def("Initialize_EIPC", &initialize_eipc, py::arg("E"), py::arg("nu"), py::arg("thickness"), py::arg("h"), py::arg("M"), py::arg("kappa"),
        py::arg("stiffMult") = 1.0);
    sh.def("Initialize_OIPC_VM", &initialize_oipc_vm, py::arg("dHat2"), py::arg("nodeAttr"), py::arg("kappa"), py::arg("stiffMult") = 1.0);
    sh.def("Initialize_OIPC", &initialize_oipc, py::arg("E"), py::arg("nu"), py::arg("thickness"), py::arg("h"), py::arg("M"), py::arg("kappa"),
        py::arg("stiffMult") = 1.0);
    sh.def("Update_Normal_Flow_Neumann", &update_normal_flow_neumann);
    // Advance_One_Step_IE_Discrete_Shell<double, 3, KL=false, elasticIPC=false, flow> (DISCRETE_SHELL.h:1110, 1113): 42 positional arguments
    auto step = [](bool flow) {
        return [flow](TriStorage& Elem, const StdVectorVector2i& seg, DbcStorage& DBC, const EdgeToTri& edge2tri, const StdVectorVector4i& edgeStencil,
                   const StdVectorVector3d& edgeInfo, double thickness, double bendingStiffMult, const Vec<double, 4>& fiberStiffMult,
                   const Vec<double, 3>& fiberLimit, Vec<double, 2>& s, Vec<double, 2>& sHat, Vec<double, 2>& kappa_s, const StdVectorXd& b, double h,
                   double NewtonTol, bool withCollision, double dHat2, Vec<double, 3>& kappaVec, double mu, double epsv2, int fricIterAmt,
                   const StdVectorXi& compNodeRange, const StdVectorXd& muComp, bool staticSolve, NodeStorage& X, NodeAttrStorage& nodeAttr, CsrMatrix& M,
                   ElemAttrStorage& elemAttr, Fcr2Storage& elasticityAttr, Storage<Vec<int, 4>>& tet, Storage<Mat<double, 3>, Mat<double, 3>>& tetAttr,
                   Fcr3Storage& tetElasticityAttr, const StdVectorVector2i& rod, const StdVectorVector3d& rodInfo, const StdVectorVector3i& rodHinge,
                   const StdVectorVector3d& rodHingeInfo, const StdVectorVector3i& stitchInfo, const StdVectorXd& stitchRatio, double k_stitch,
                   const StdVectorXi& particle, const std::string& outputFolder) {
            (void)edge2tri; (void)fiberLimit; (void)s; (void)sHat; (void)tetAttr; (void)tetElasticityAttr; (void)rodInfo;
            (void)rodHinge; (void)rodHingeInfo; (void)stitchRatio; (void)k_stitch;
            ShellStepInputs in;
            in.flow = flow; in.thickness = thickness; in.bendingStiffMult = bendingStiffMult; in.h = h; in.NewtonTol = NewtonTol; in.dHat2 = dHat2;
            in.mu = mu; in.epsv2 = epsv2; in.fricIterAmt = fricIterAmt;
            in.compNodeRange = compNodeRange; in.muComp = muComp;
            in.withCollision = withCollision; in.staticSolve = staticSolve;
            in.nTet = tet.size(); in.nRod = (int)rod.size(); in.nStitch = (int)stitchInfo.size(); in.nParticle = (int)particle.size();
            in.outputFolder = outputFolder;
#ifdef JGSL_STEP_HOOK
            {
                int hooked = 0;
                if (JGSL_STEP_HOOK(in, Elem, DBC, edgeStencil, edgeInfo, b, kappaVec, X, nodeAttr, elemAttr, elasticityAttr, &hooked)) { fflush(stdout); return hooked; }
            }
#endif
            JGSL_BACKEND_CLASS& be = backend();
            // the linear solve is iterative where the reference factorises: its stopping rule is a parameter of this build
            // (Set_Parameter("B200.pcg_rel_tol", 1e-12), Set_Parameter("B200.pcg_max_iter", 200000))
            be.pcg_rel_tol = get_param<double>("B200.pcg_rel_tol", 1e-12);
            be.pcg_max_iter = get_param<int>("B200.pcg_max_iter", 200000);
            const int it = advance_one_step_ie(be, in, Elem, seg, DBC, edgeStencil, edgeInfo, fiberStiffMult, kappa_s, b, kappaVec, X, nodeAttr, M, elemAttr,
                elasticityAttr);
            fflush(stdout);
            return it;
        };
    };
    sh.def("Advance_One_Step_IE_Flow", step(true));
    sh.def("Advance_One_Step_IE_Hinge", step(false));
    for (const char* n : {"Add_Garment", "Make_Rod", "Make_Rod_Net", "Add_Discrete_Particles", "Initialize_Shell", "Initialize_Garment",
             "Update_Material_With_Tex_Shell", "Initialize_Shell_EIPC", "Initialize_Discrete_Rod",
             "Initialize_Discrete_Particle", "Advance_One_Step_IE",
             "Advance_One_Step_IE_EIPC", "Advance_One_Step_IE_Hinge_EIPC", "Advance_One_Step_SIE", "Advance_One_Step_SIE_Hinge",
             "Advance_One_Step_SIE_EIPC", "Advance_One_Step_SIE_Hinge_EIPC", "Construct_Surface_Mesh", "Compute_Stretch_From_File",
             "XZ_As_Texture", "Adjust_Material"})
        not_built(sh, n);

    // run statistics of the hosted path (not part of the reference's surface)
    m.def("B200_Backend_Name", []() { return std::string(backend().name()); });
    m.def("B200_Release", []() { backend_slot().reset(); });
}
