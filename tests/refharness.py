"""ctypes view of oracle/_ref/libnlps2d_ref.so (the reference's own 2D TUs + oracle/ref_harness.c).

Test infrastructure only.  One reference "simulation" per process (the reference keeps its
state in process globals), so golden generation runs each deck in a fresh subprocess.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "..", "oracle", "_ref", "libnlps2d_ref.so")

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def available(path: str = REF_SO) -> bool:
    return os.path.exists(path)


class RefHarness:
    def __init__(self, deck: str, so: str = REF_SO, threads: int = 1):
        self.lib = ctypes.CDLL(os.path.abspath(so))
        L = self.lib
        for f in ("refh_delta_x", "refh_cfl", "refh_cel", "refh_dt", "refh_gamma_lme",
                  "refh_tol_zero_lme", "refh_tol_wrapper_lme", "refh_tol_radial", "refh_thickness"):
            getattr(L, f).restype = ctypes.c_double
        L.refh_material_type.restype = ctypes.c_char_p
        L.refh_set_threads(threads)
        # the reference builds every path relative to the deck ("./" + dirs, Parser.c:44-62)
        cwd = os.getcwd()
        os.chdir(os.path.dirname(os.path.abspath(deck)))
        try:
            rc = L.refh_init(os.path.basename(deck).encode())
        finally:
            os.chdir(cwd)
        if rc != 0:
            raise RuntimeError(f"refh_init failed: {rc}")
        self.ndim = L.refh_ndim()
        self.nn = L.refh_num_nodes()
        self.ne = L.refh_num_elems()
        self.np_ = L.refh_num_particles()
        self.nsteps = L.refh_num_steps()

    # ---- mesh
    def coords(self):
        out = np.zeros((self.nn, self.ndim))
        self.lib.refh_get_coords(out.ctypes.data_as(_dp))
        return out

    def h_avg(self):
        out = np.zeros(self.nn)
        self.lib.refh_get_h_avg(out.ctypes.data_as(_dp))
        return out

    def table(self, which: int):
        n = {0: self.ne, 1: self.nn, 2: self.nn, 3: self.nn, 4: self.np_}[which]
        tot = self.lib.refh_table_total(which)
        ptr = np.zeros(n + 1, dtype=np.int32)
        idx = np.zeros(max(tot, 1), dtype=np.int32)
        self.lib.refh_table_csr(which, ptr.ctypes.data_as(_ip), idx.ctypes.data_as(_ip))
        return ptr, idx[:tot]

    def active(self):
        out = np.zeros(self.nn, dtype=np.uint8)
        self.lib.refh_get_active(out.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)))
        return out

    def bounds(self):
        res = []
        for b in range(self.lib.refh_num_bounds()):
            nb = self.lib.refh_bound_num_nodes(b)
            dim = self.lib.refh_bound_dim(b)
            nodes = np.zeros(nb, dtype=np.int32)
            self.lib.refh_bound_nodes(b, nodes.ctypes.data_as(_ip))
            dirs = np.zeros((dim, self.nsteps), dtype=np.int32)
            vals = np.zeros((dim, self.nsteps))
            self.lib.refh_bound_table(b, dirs.ctypes.data_as(_ip), vals.ctypes.data_as(_dp))
            res.append(dict(nodes=nodes, dir=dirs, val=vals))
        return res

    def neumann(self):
        res = []
        for b in range(self.lib.refh_num_neumann()):
            nb = self.lib.refh_neumann_num_nodes(b)
            dim = self.lib.refh_neumann_dim(b)
            nodes = np.zeros(nb, dtype=np.int32)
            self.lib.refh_neumann_nodes(b, nodes.ctypes.data_as(_ip))
            dirs = np.zeros((dim, self.nsteps), dtype=np.int32)
            vals = np.zeros((dim, self.nsteps))
            self.lib.refh_neumann_table(b, dirs.ctypes.data_as(_ip), vals.ctypes.data_as(_dp))
            res.append(dict(nodes=nodes, dir=dirs, val=vals))
        return res

    def gravity(self):
        g = np.zeros((self.ndim, self.nsteps))
        self.lib.refh_gravity_table(g.ctypes.data_as(_dp))
        return g

    def material(self, m=0):
        p = np.zeros(16)
        self.lib.refh_material_params(m, p.ctypes.data_as(_dp))
        t = self.lib.refh_material_type(m).decode()
        if t == "Von-Mises" and hasattr(self.lib, "refh_material_voce"):
            v = np.zeros(4)
            self.lib.refh_material_voce(m, v.ctypes.data_as(_dp))
            p = np.concatenate([p, v])
        return t, p

    # ---- particles
    def field(self, name: str):
        c = self.lib.refh_field_cols(name.encode())
        if c < 0:
            raise KeyError(name)
        out = np.zeros((self.np_, c))
        self.lib.refh_get_field(name.encode(), out.ctypes.data_as(_dp))
        return out if c > 1 else out[:, 0].copy()

    def set_field(self, name: str, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        assert self.lib.refh_set_field(name.encode(), a.ctypes.data_as(_dp)) == 0

    def ints(self, name: str):
        out = np.zeros(self.np_, dtype=np.int32)
        self.lib.refh_get_ints(name.encode(), out.ctypes.data_as(_ip))
        return out

    def shape(self, p: int):
        N = np.zeros(256)
        dN = np.zeros(256 * self.ndim)
        n = self.lib.refh_shape(p, N.ctypes.data_as(_dp), dN.ctypes.data_as(_dp))
        return N[:n].copy(), dN[: n * self.ndim].reshape(n, self.ndim).copy()

    def local_search(self):
        return self.lib.refh_local_search()

    def step(self, k: int):
        return self.lib.refh_step(k)

    def nodal(self, which: int):
        out = np.zeros((self.nn, self.ndim))
        self.lib.refh_get_nodal(which, out.ctypes.data_as(_dp))
        return out

    def stage_times(self):
        t = np.zeros(5)
        self.lib.refh_stage_times(t.ctypes.data_as(_dp))
        return t

    def scalars(self):
        L = self.lib
        return dict(delta_x=L.refh_delta_x(), cfl=L.refh_cfl(), cel=L.refh_cel(), dt=L.refh_dt(),
                    gamma_lme=L.refh_gamma_lme(), tol_zero=L.refh_tol_zero_lme(),
                    tol_wrapper=L.refh_tol_wrapper_lme(), max_iter_lme=L.refh_max_iter_lme(),
                    tol_radial=L.refh_tol_radial(), maxiter_radial=L.refh_maxiter_radial(),
                    thickness=L.refh_thickness())

    def stress_point(self, p, DF, F_n1, J_n1, b_e_n, eps_n, kappa_n):
        T = 5 if self.ndim == 2 else 9
        d = self.ndim
        arr = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        DF, F_n1, b_e_n = arr(DF), arr(F_n1), arr(b_e_n)
        stress, be1, cep = np.zeros(T), np.zeros(T), np.zeros(d * d)
        e1, k1, W = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        st = self.lib.refh_stress_point(
            p, DF.ctypes.data_as(_dp), F_n1.ctypes.data_as(_dp), ctypes.c_double(J_n1),
            b_e_n.ctypes.data_as(_dp), ctypes.c_double(eps_n), ctypes.c_double(kappa_n),
            stress.ctypes.data_as(_dp), be1.ctypes.data_as(_dp), ctypes.byref(e1),
            ctypes.byref(k1), ctypes.byref(W), cep.ctypes.data_as(_dp))
        return dict(status=st, stress=stress, b_e_n1=be1, eps_n1=e1.value, kappa_n1=k1.value,
                    W=W.value, C_ep=cep)


# -- the reference's tangent blocks (implicit scheme, K5): pure functions, no deck needed -----------------------
def _arr(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def stiffness_ep(u, v, b_e, stress, c_ep, so: str = REF_SO):
    """compute_stiffness_elastoplastic__Constitutive__ (Elastoplastic-Tangent-Matrix.c:42-160), 2D build."""
    L = ctypes.CDLL(os.path.abspath(so))
    u, v, b_e, stress, c_ep = _arr(u), _arr(v), _arr(b_e).copy(), _arr(stress).copy(), _arr(c_ep).copy()
    out = np.zeros(4)
    rc = L.refh_stiffness_ep(out.ctypes.data_as(_dp), u.ctypes.data_as(_dp), v.ctypes.data_as(_dp),
                             b_e.ctypes.data_as(_dp), stress.ctypes.data_as(_dp), c_ep.ctypes.data_as(_dp))
    assert rc == 0
    return out


def stiffness_nh(u, v, un, vn, F_n, J, E, nu, so: str = REF_SO):
    """compute_stiffness_density_Neo_Hookean (Neo-Hookean.c:89-141), 2D build."""
    L = ctypes.CDLL(os.path.abspath(so))
    u, v, un, vn, F_n = _arr(u), _arr(v), _arr(un), _arr(vn), _arr(F_n).copy()
    out = np.zeros(4)
    rc = L.refh_stiffness_nh(out.ctypes.data_as(_dp), u.ctypes.data_as(_dp), v.ctypes.data_as(_dp),
                             un.ctypes.data_as(_dp), vn.ctypes.data_as(_dp), F_n.ctypes.data_as(_dp),
                             ctypes.c_double(J), ctypes.c_double(E), ctypes.c_double(nu))
    assert rc == 0
    return out
