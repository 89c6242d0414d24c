"""ctypes view of the C ABI in include/idp_contact.h (libidp_contact.so).

This is the harness tests/ and bench.py drive the library with; the product's host side is the C++ header
idp_b200/host/IPC_B200.h, which mirrors the reference's six operators (Library/FEM/IPC.h). There is no CPU
fallback here: if the shared library or a CUDA device is missing, construction raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IDP_LIB_PATH") or os.path.join(_HERE, "libidp_contact.so")  # override: build-variant experiments only

STATUS = {0: "IDP_OK", 1: "IDP_ERR_CUDA", 2: "IDP_ERR_INVALID", 3: "IDP_ERR_NONPOSITIVE_DISTANCE",
          4: "IDP_ERR_CCD_ZERO_STEP", 5: "IDP_ERR_UNSUPPORTED_PRIMITIVE", 6: "IDP_ERR_NCCL",
          7: "IDP_ERR_CCD_ITERATION_CAP", 8: "IDP_ERR_EIGEN_NO_CONVERGENCE"}
STAGES = ["Compute_Constraint_Set_Build_Hash", "Compute_Constraint_Set_PT", "Compute_Constraint_Set_EE",
          "Compute_Constraint_Set_Merge", "Compute_Barrier_EgH", "constructCSRMatrixFromTriplet",
          "Compute_Intersection_Free_StepSize_Build_Hash", "Compute_Intersection_Free_StepSize_PT",
          "Compute_Intersection_Free_StepSize_EE", "Compute_Min_Dist", "upload", "k_barrier", "k_query", "k_accd",
          "k_classify", "nccl_collectives"]

# every symbol include/idp_contact.h declares
EXPORTS = ["idp_create", "idp_destroy", "idp_last_error", "idp_set_stream", "idp_set_mesh", "idp_declare_unsupported",
           "idp_set_positions", "idp_set_rest_positions", "idp_constraint_set", "idp_get_constraints",
           "idp_gather_constraints", "idp_comm_init_local", "idp_comm_abort",
           "idp_set_constraints", "idp_get_candidates", "idp_barrier_energy", "idp_barrier_gradient",
           "idp_barrier_hessian", "idp_barrier_all", "idp_get_hessian_csr", "idp_hessian_csr_device",
           "idp_gradient_device", "idp_get_gradient", "idp_ccd_step", "idp_set_search_direction", "idp_ccd_step_resident", "idp_min_dist2", "idp_comm_unique_id", "idp_comm_init",
           "idp_set_shard", "idp_kernel_launches", "idp_library_calls", "idp_reset_counters", "idp_stage_ms",
           "idp_last_count", "idp_measure_fp64_tflops", "idp_system_set_flow_term", "idp_system_set_mass", "idp_project_dbc",
           "idp_solve_pcg", "idp_set_mesh_from_triangles", "idp_get_surface_primitives", "idp_get_constraints_begin",
           "idp_get_hessian_csr_begin", "idp_transfers_end", "idp_system_set_membrane", "idp_system_set_hinges", "idp_elastic_energy",
           "idp_elastic_gradient", "idp_project_dbc_mask", "idp_friction_update", "idp_friction_set", "idp_friction_energy",
           "idp_friction_gradient", "idp_get_friction", "idp_friction_set_components"]


class IdpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s: %s" % (STATUS.get(code, code), msg))
        self.code = code


def load_library(path=LIB_PATH):
    if not os.path.exists(path):
        raise FileNotFoundError("%s is missing: run __graft_entry__.build() (make -C idp_b200/csrc)" % path)
    L = C.CDLL(path)
    vp, i, d, l = C.c_void_p, C.c_int, C.c_double, C.c_long
    L.idp_create.argtypes = [i, C.POINTER(vp)]
    L.idp_destroy.argtypes = [vp]
    L.idp_destroy.restype = None
    L.idp_last_error.argtypes = [vp]
    L.idp_last_error.restype = C.c_char_p
    L.idp_set_stream.argtypes = [vp, vp]
    L.idp_set_mesh.argtypes = [vp, i, i, vp, i, vp, i, vp, vp]
    L.idp_declare_unsupported.argtypes = [vp, i, i, i]
    L.idp_set_positions.argtypes = [vp, vp, i]
    L.idp_set_rest_positions.argtypes = [vp, vp, i]
    L.idp_constraint_set.argtypes = [vp, d, d, C.POINTER(i)]
    L.idp_get_constraints.argtypes = [vp, vp, vp]
    L.idp_gather_constraints.argtypes = [vp, vp, vp, vp]
    L.idp_comm_init_local.argtypes = [C.POINTER(vp), i]
    L.idp_comm_abort.argtypes = [vp]
    L.idp_set_constraints.argtypes = [vp, i, vp, vp]
    L.idp_get_candidates.argtypes = [vp, i, C.POINTER(l), vp]
    L.idp_barrier_energy.argtypes = [vp, d, d, d, C.POINTER(d)]
    L.idp_barrier_gradient.argtypes = [vp, d, d, d, vp, i]
    L.idp_barrier_hessian.argtypes = [vp, d, d, d, i, C.POINTER(l)]
    L.idp_barrier_all.argtypes = [vp, d, d, d, i, C.POINTER(d), C.POINTER(l)]
    L.idp_get_hessian_csr.argtypes = [vp, vp, vp, vp]
    L.idp_hessian_csr_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(l)]
    L.idp_gradient_device.argtypes = [vp, C.POINTER(vp)]
    L.idp_get_gradient.argtypes = [vp, vp, i]
    L.idp_ccd_step.argtypes = [vp, vp, i, d, C.POINTER(d)]
    L.idp_set_search_direction.argtypes = [vp, vp, i]
    L.idp_ccd_step_resident.argtypes = [vp, d, C.POINTER(d)]
    L.idp_min_dist2.argtypes = [vp, d, vp, C.POINTER(d)]
    L.idp_comm_unique_id.argtypes = [vp]
    L.idp_comm_init.argtypes = [vp, i, i, vp]
    L.idp_set_shard.argtypes = [vp, i, i]
    L.idp_kernel_launches.argtypes = [vp]
    L.idp_kernel_launches.restype = l
    L.idp_library_calls.argtypes = [vp]
    L.idp_library_calls.restype = l
    L.idp_reset_counters.argtypes = [vp]
    L.idp_reset_counters.restype = None
    L.idp_stage_ms.argtypes = [vp, i]
    L.idp_stage_ms.restype = C.c_float
    L.idp_last_count.argtypes = [vp, i]
    L.idp_last_count.restype = l
    L.idp_measure_fp64_tflops.argtypes = [vp, C.POINTER(d)]
    L.idp_system_set_flow_term.argtypes = [vp, i, vp, i, vp, d]
    L.idp_system_set_mass.argtypes = [vp, vp]
    L.idp_project_dbc.argtypes = [vp]
    L.idp_project_dbc_mask.argtypes = [vp, vp]
    L.idp_friction_update.argtypes = [vp, d, d, d, C.POINTER(l)]
    L.idp_friction_set.argtypes = [vp, vp, i, d, d]
    L.idp_friction_energy.argtypes = [vp, C.POINTER(d)]
    L.idp_friction_set_components.argtypes = [vp, i, vp, vp]
    L.idp_friction_gradient.argtypes = [vp, vp, i]
    L.idp_get_friction.argtypes = [vp, C.POINTER(l), vp, vp, vp, vp]
    L.idp_system_set_membrane.argtypes = [vp, i, vp, i, vp, vp, vp, vp, d]
    L.idp_system_set_hinges.argtypes = [vp, i, vp, vp, d, d]
    L.idp_elastic_energy.argtypes = [vp, C.POINTER(d)]
    L.idp_elastic_gradient.argtypes = [vp, vp, i]
    L.idp_get_constraints_begin.argtypes = [vp, vp, vp]
    L.idp_get_hessian_csr_begin.argtypes = [vp, vp, vp, vp]
    L.idp_transfers_end.argtypes = [vp]
    L.idp_solve_pcg.argtypes = [vp, vp, vp, d, i, C.POINTER(i), C.POINTER(d)]
    L.idp_set_mesh_from_triangles.argtypes = [vp, i, i, vp, i, vp, i, vp]
    L.idp_get_surface_primitives.argtypes = [vp, C.POINTER(i), vp, C.POINTER(i), vp, C.POINTER(i), vp, vp, vp, vp]
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class ContactContext:
    """One GPU context. Methods mirror the C entry points one to one."""

    def __init__(self, device=0, lib=None):
        self.L = lib or load_library()
        h = C.c_void_p()
        st = self.L.idp_create(device, C.byref(h))
        if st != 0:
            raise IdpError(st, "idp_create(device=%d) failed (no CUDA device? there is no CPU fallback)" % device)
        self.h = h
        self.nV = 0
        self._keep = []

    def close(self):
        if self.h:
            self.L.idp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st):
        if st != 0:
            raise IdpError(st, self.L.idp_last_error(self.h).decode())

    # ---- inputs ----
    def set_stream(self, cuda_stream):
        self._ck(self.L.idp_set_stream(self.h, C.c_void_p(cuda_stream)))

    def set_mesh(self, nV, bnode, bedge, btri, dbc=None):
        bnode = np.ascontiguousarray(bnode, np.int32)
        bedge = np.ascontiguousarray(bedge, np.int32).reshape(-1, 2)
        btri = np.ascontiguousarray(btri, np.int32).reshape(-1, 3)
        dbc = None if dbc is None else np.ascontiguousarray(dbc, np.uint8)
        self.nV = int(nV)
        self._ck(self.L.idp_set_mesh(self.h, nV, len(bnode), _p(bnode), len(bedge), _p(bedge), len(btri), _p(btri), _p(dbc)))

    def set_surface_mesh(self, mesh):
        """mesh: idp_b200.meshgen.SurfaceMesh"""
        self.set_mesh(mesh.nV, mesh.bnode, mesh.bedge, mesh.btri, mesh.dbc)
        self.set_rest_positions(mesh.X0)
        self.set_positions(mesh.X)

    def set_positions(self, X):
        X = np.ascontiguousarray(X, np.float64)
        self._ck(self.L.idp_set_positions(self.h, _p(X), X.shape[1]))

    def set_rest_positions(self, X0):
        X0 = np.ascontiguousarray(X0, np.float64)
        self._ck(self.L.idp_set_rest_positions(self.h, _p(X0), X0.shape[1]))

    def declare_unsupported(self, n_rod=0, n_particle=0, n_nn=0):
        self._ck(self.L.idp_declare_unsupported(self.h, n_rod, n_particle, n_nn))

    # ---- constraint set ----
    def constraint_set(self, dhat2, thickness=0.0):
        n = C.c_int(0)
        self._ck(self.L.idp_constraint_set(self.h, dhat2, thickness, C.byref(n)))
        return n.value

    def get_constraints(self):
        """The rows this context holds (sharded: this rank's rows only)."""
        n = self.L.idp_last_count(self.h, 9)
        rows = np.empty((n, 4), np.int32)
        info = np.empty((n, 2), np.float64)
        self._ck(self.L.idp_get_constraints(self.h, _p(rows), _p(info)))
        return rows, info

    def gather_constraints(self, want_dist2=False):
        """The global list in the order of the unsharded path (sharded: a collective)."""
        n = self.L.idp_last_count(self.h, 0)
        rows = np.empty((n, 4), np.int32)
        info = np.empty((n, 2), np.float64)
        d = np.empty(n, np.float64) if want_dist2 else None
        self._ck(self.L.idp_gather_constraints(self.h, _p(rows), _p(info), _p(d)))
        return (rows, info, d) if want_dist2 else (rows, info)

    def set_constraints(self, rows, info=None):
        rows = np.ascontiguousarray(rows, np.int32).reshape(-1, 4)
        info = None if info is None else np.ascontiguousarray(info, np.float64).reshape(-1, 2)
        self._ck(self.L.idp_set_constraints(self.h, len(rows), _p(rows), _p(info)))

    def get_candidates(self, which):
        n = C.c_long(0)
        self._ck(self.L.idp_get_candidates(self.h, which, C.byref(n), None))
        out = np.empty((n.value, 2), np.int32)
        if n.value:
            self._ck(self.L.idp_get_candidates(self.h, which, C.byref(n), _p(out)))
        return out

    # ---- the rest of the Newton system (flow term, mass, Project_DBC, PCG, surface extraction) ----
    def set_flow_term(self, elem, vol, h):
        if elem is None or len(elem) == 0:
            return self._ck(self.L.idp_system_set_flow_term(self.h, 0, None, 3, None, 0.0))
        elem = np.ascontiguousarray(elem, np.int32)
        vol = np.ascontiguousarray(vol, np.float64)
        self._ck(self.L.idp_system_set_flow_term(self.h, len(elem), _p(elem), elem.shape[1], _p(vol), float(h)))

    def set_mass(self, m):
        m = None if m is None else np.ascontiguousarray(m, np.float64)
        self._ck(self.L.idp_system_set_mass(self.h, _p(m)))

    def set_membrane(self, elem, ib, vol, lam, mu, h):
        """Membrane triangles: rest first fundamental form ib = (IB00, IB01, IB11), vol, lambda, mu per element; weight h^2."""
        if elem is None or len(elem) == 0:
            return self._ck(self.L.idp_system_set_membrane(self.h, 0, None, 3, None, None, None, None, 0.0))
        elem = np.ascontiguousarray(elem, np.int32)
        n = len(elem)
        ib = np.ascontiguousarray(ib, np.float64).reshape(n, 3)
        vol, lam, mu = (np.ascontiguousarray(np.broadcast_to(a, n), np.float64) for a in (vol, lam, mu))
        self._ck(self.L.idp_system_set_membrane(self.h, n, _p(elem), elem.shape[1], _p(ib), _p(vol), _p(lam), _p(mu), float(h)))

    def set_hinges(self, stencil, info, k, h):
        """Bending hinges: stencil (v0; v1, v2; v3), info = (rest angle, rest edge length, rest height); energy h^2 k (..)^2 e/h."""
        if stencil is None or len(stencil) == 0:
            return self._ck(self.L.idp_system_set_hinges(self.h, 0, None, None, 0.0, 0.0))
        stencil = np.ascontiguousarray(stencil, np.int32).reshape(-1, 4)
        info = np.ascontiguousarray(info, np.float64).reshape(-1, 3)
        self._ck(self.L.idp_system_set_hinges(self.h, len(stencil), _p(stencil), _p(info), float(k), float(h)))

    def elastic_energy(self, E0=0.0):
        E = C.c_double(E0)
        self._ck(self.L.idp_elastic_energy(self.h, C.byref(E)))
        return E.value

    def elastic_gradient(self, g_accum=None):
        g = np.zeros((self.nV, 3), np.float64) if g_accum is None else g_accum
        self._ck(self.L.idp_elastic_gradient(self.h, _p(g), g.shape[1]))
        return g

    def friction_update(self, dhat2, kappa, thickness=0.0):
        n = C.c_long(0)
        self._ck(self.L.idp_friction_update(self.h, dhat2, kappa, thickness, C.byref(n)))
        return n.value

    def friction_set(self, Xn, epsv2_h2, mu):
        Xn = None if Xn is None else np.ascontiguousarray(Xn, np.float64)
        self._ck(self.L.idp_friction_set(self.h, _p(Xn), 3 if Xn is None else Xn.shape[1], float(epsv2_h2), float(mu)))

    def friction_set_components(self, comp_node_range, mu_comp):
        if comp_node_range is None or len(comp_node_range) == 0:
            return self._ck(self.L.idp_friction_set_components(self.h, 0, None, None))
        r = np.ascontiguousarray(comp_node_range, np.int32)
        m = np.ascontiguousarray(mu_comp, np.float64).reshape(len(r), len(r))
        self._ck(self.L.idp_friction_set_components(self.h, len(r), _p(r), _p(m)))

    def friction_energy(self, E0=0.0):
        E = C.c_double(E0)
        self._ck(self.L.idp_friction_energy(self.h, C.byref(E)))
        return E.value

    def friction_gradient(self, g_accum=None):
        g = np.zeros((self.nV, 3), np.float64) if g_accum is None else g_accum
        self._ck(self.L.idp_friction_gradient(self.h, _p(g), g.shape[1]))
        return g

    def get_friction(self):
        n = C.c_long(0)
        self._ck(self.L.idp_get_friction(self.h, C.byref(n), None, None, None, None))
        n = n.value
        rows = np.zeros((n, 4), np.int32); cp = np.zeros((n, 2)); basis = np.zeros((n, 6)); nf = np.zeros(n)
        if n:
            self._ck(self.L.idp_get_friction(self.h, None, _p(rows), _p(cp), _p(basis), _p(nf)))
        return rows, cp, basis, nf

    def project_dbc(self, mask=None):
        if mask is None:
            return self._ck(self.L.idp_project_dbc(self.h))
        mask = np.ascontiguousarray(mask, np.uint8)
        self._ck(self.L.idp_project_dbc_mask(self.h, _p(mask)))

    def solve_pcg(self, rhs, rel_tol=1e-10, max_iter=10000):
        rhs = np.ascontiguousarray(rhs, np.float64).reshape(-1)
        sol = np.empty_like(rhs)
        it = C.c_int(0)
        res = C.c_double(0.0)
        self._ck(self.L.idp_solve_pcg(self.h, _p(rhs), _p(sol), rel_tol, max_iter, C.byref(it), C.byref(res)))
        return sol, it.value, res.value

    def set_mesh_from_triangles(self, nV, tri, X=None, dbc=None):
        tri = np.ascontiguousarray(tri, np.int32)
        X = None if X is None else np.ascontiguousarray(X, np.float64)
        dbc = None if dbc is None else np.ascontiguousarray(dbc, np.uint8)
        self.nV = nV
        self._ck(self.L.idp_set_mesh_from_triangles(self.h, nV, len(tri), _p(tri), tri.shape[1], _p(X), 0 if X is None else X.shape[1], _p(dbc)))

    def get_surface_primitives(self, areas=False):
        n = [C.c_int(0) for _ in range(3)]
        self._ck(self.L.idp_get_surface_primitives(self.h, C.byref(n[0]), None, C.byref(n[1]), None, C.byref(n[2]), None, None, None, None))
        nN, nE, nT = (k.value for k in n)
        out = dict(bnode=np.empty(nN, np.int32), bedge=np.empty((nE, 2), np.int32), btri=np.empty((nT, 3), np.int32))
        ar = dict(BNArea=np.empty(nN), BEArea=np.empty(nE), BTArea=np.empty(nT)) if areas else dict(BNArea=None, BEArea=None, BTArea=None)
        self._ck(self.L.idp_get_surface_primitives(self.h, None, _p(out["bnode"]), None, _p(out["bedge"]), None, _p(out["btri"]),
                                                   _p(ar["BNArea"]), _p(ar["BEArea"]), _p(ar["BTArea"])))
        if areas:
            out.update(ar)
        return out

    # ---- barrier ----
    def barrier_energy(self, dhat2, kappa, thickness=0.0, E0=0.0):
        E = C.c_double(E0)
        self._ck(self.L.idp_barrier_energy(self.h, dhat2, kappa, thickness, C.byref(E)))
        return E.value

    def barrier_gradient(self, dhat2, kappa, thickness=0.0, g_accum=None):
        g = np.zeros((self.nV, 3), np.float64) if g_accum is None else g_accum
        self._ck(self.L.idp_barrier_gradient(self.h, dhat2, kappa, thickness, _p(g), g.shape[1]))
        return g

    def barrier_hessian(self, dhat2, kappa, thickness=0.0, project_spd=True, fetch=True):
        nnz = C.c_long(0)
        self._ck(self.L.idp_barrier_hessian(self.h, dhat2, kappa, thickness, int(project_spd), C.byref(nnz)))
        return self.get_hessian_csr() if fetch else nnz.value

    def barrier_all(self, dhat2, kappa, thickness=0.0, project_spd=True):
        E = C.c_double(0.0)
        nnz = C.c_long(0)
        self._ck(self.L.idp_barrier_all(self.h, dhat2, kappa, thickness, int(project_spd), C.byref(E), C.byref(nnz)))
        return E.value, nnz.value

    def get_hessian_csr(self):
        nnz = self.L.idp_last_count(self.h, 6)
        ptr = np.empty(3 * self.nV + 1, np.int32)
        col = np.empty(nnz, np.int32)
        val = np.empty(nnz, np.float64)
        self._ck(self.L.idp_get_hessian_csr(self.h, _p(ptr), _p(col), _p(val)))
        return ptr, col, val

    # ---- CCD / min distance ----
    def ccd_step(self, direction, alpha=1.0, thickness=0.0):
        direction = np.ascontiguousarray(direction, np.float64)
        a = C.c_double(alpha)
        self._ck(self.L.idp_ccd_step(self.h, _p(direction), direction.shape[1], thickness, C.byref(a)))
        return a.value

    def set_search_direction(self, direction):
        direction = np.ascontiguousarray(direction, np.float64)
        self._ck(self.L.idp_set_search_direction(self.h, _p(direction), direction.shape[1]))

    def ccd_step_resident(self, alpha=1.0, thickness=0.0):
        a = C.c_double(alpha)
        self._ck(self.L.idp_ccd_step_resident(self.h, thickness, C.byref(a)))
        return a.value

    def min_dist2(self, thickness=0.0, want_all=True):
        n = self.L.idp_last_count(self.h, 9)
        d = np.empty(n, np.float64) if want_all else None
        m = C.c_double(np.nan)
        self._ck(self.L.idp_min_dist2(self.h, thickness, _p(d), C.byref(m)))
        return d, m.value

    # ---- sharding / instrumentation ----
    def set_shard(self, rank, nranks):
        self._ck(self.L.idp_set_shard(self.h, rank, nranks))

    @staticmethod
    def comm_init_local(ctxs):
        """Make ctxs[r] rank r of an in-process group (each context must then be driven by its own thread)."""
        arr = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
        st = ctxs[0].L.idp_comm_init_local(arr, len(ctxs))
        if st != 0:
            raise IdpError(st, ctxs[0].L.idp_last_error(ctxs[0].h).decode())

    def comm_abort(self):
        self.L.idp_comm_abort(self.h)

    def comm_init(self, rank, nranks, uid_bytes):
        buf = C.create_string_buffer(bytes(uid_bytes), 128)
        self._ck(self.L.idp_comm_init(self.h, rank, nranks, buf))

    def unique_id(self):
        buf = C.create_string_buffer(128)
        st = self.L.idp_comm_unique_id(buf)
        if st != 0:
            raise IdpError(st, "ncclGetUniqueId failed")
        return buf.raw

    def launches(self):
        return self.L.idp_kernel_launches(self.h), self.L.idp_library_calls(self.h)

    def reset_counters(self):
        self.L.idp_reset_counters(self.h)

    def stage_ms(self):
        return {STAGES[i]: float(self.L.idp_stage_ms(self.h, i)) for i in range(len(STAGES))}

    def count(self, what):
        return self.L.idp_last_count(self.h, what)

    def fp64_tflops(self):
        t = C.c_double(0)
        self._ck(self.L.idp_measure_fp64_tflops(self.h, C.byref(t)))
        return t.value
