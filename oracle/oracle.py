"""ctypes wrapper of oracle/liboracle.so (oracle/nlps_oracle.c).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "liboracle.so")
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)
_up = ctypes.POINTER(ctypes.c_ubyte)

MAT = {"Neo-Hookean-Wriggers": 0, "Drucker-Prager": 1, "Matsuoka-Nakai": 2, "Von-Mises": 3, "Hencky": 4, "Lade-Duncan": 5}
STAGES = dict(search=0, p2g_mass_disp=1, grid_disp=2, kin_stress=3, force=4, grid_acc=5, g2p=6)


def build(force=False):
    if force or not os.path.exists(SO) or any(
            os.path.getmtime(os.path.join(HERE, f)) > os.path.getmtime(SO)
            for f in ("nlps_oracle.c", "mini_lapack.c")):
        subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(SO)
        L.orc_create.restype = ctypes.c_void_p
        L.orc_locality_build.restype = ctypes.c_void_p
        L.orc_dt.restype = ctypes.c_double
        _lib = L
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def build_locality(ndim, coords, conn):
    """Restated adjacency (ring1/ring2 CSR in chain order, h_avg, DeltaX)."""
    L = lib()
    coords, conn = _d(coords), _i(conn)
    nn, (ne, nne) = coords.shape[0], conn.shape
    h = ctypes.c_void_p(L.orc_locality_build(ndim, nn, ne, nne, conn.ctypes.data_as(_ip),
                                             coords.ctypes.data_as(_dp)))
    n1, n2 = L.orc_locality_size(h, 1), L.orc_locality_size(h, 2)
    r1p, r2p = np.zeros(nn + 1, np.int32), np.zeros(nn + 1, np.int32)
    r1i, r2i = np.zeros(n1, np.int32), np.zeros(n2, np.int32)
    h_avg = np.zeros(nn)
    dx = ctypes.c_double()
    L.orc_locality_get(h, r1p.ctypes.data_as(_ip), r1i.ctypes.data_as(_ip), r2p.ctypes.data_as(_ip),
                       r2i.ctypes.data_as(_ip), h_avg.ctypes.data_as(_dp), ctypes.byref(dx))
    L.orc_locality_free(h)
    return r1p, r1i, r2p, r2i, h_avg, dx.value


class Oracle:
    def __init__(self, prob, threads=1):
        L = lib()
        self.L = L
        L.orc_set_threads(threads)
        self.prob = prob
        self.d = prob.ndim
        self.h = ctypes.c_void_p(L.orc_create(prob.ndim))
        h = self.h
        c, r1p, r1i, r2p, r2i, ha = (_d(prob.coords), _i(prob.r1p), _i(prob.r1i), _i(prob.r2p),
                                     _i(prob.r2i), _d(prob.h_avg))
        L.orc_set_mesh(h, prob.nn, c.ctypes.data_as(_dp), r1p.ctypes.data_as(_ip),
                       r1i.ctypes.data_as(_ip), r2p.ctypes.data_as(_ip), r2i.ctypes.data_as(_ip),
                       ha.ctypes.data_as(_dp), ctypes.c_double(prob.dx))
        s = prob.solver
        L.orc_set_solver(h, ctypes.c_double(s["cfl"]), ctypes.c_double(s["cel"]), int(s["nsteps"]),
                         ctypes.c_double(s["gamma_lme"]), ctypes.c_double(s["tol_zero"]),
                         ctypes.c_double(s["tol_wrapper"]), int(s["max_iter_lme"]),
                         ctypes.c_double(s["tol_radial"]), int(s["maxiter_radial"]),
                         ctypes.c_double(s.get("thickness", 1.0)))
        for b in prob.bounds:
            n, di, v = _i(b["nodes"]), _i(b["dir"]), _d(b["val"])
            L.orc_add_bound(h, len(n), n.ctypes.data_as(_ip), di.shape[0], di.ctypes.data_as(_ip),
                            v.ctypes.data_as(_dp))
        for b in prob.neumann:
            n, di, v = _i(b["nodes"]), _i(b["dir"]), _d(b["val"])
            L.orc_add_neumann(h, len(n), n.ctypes.data_as(_ip), di.shape[0], di.ctypes.data_as(_ip),
                              v.ctypes.data_as(_dp))
        g = _d(prob.gravity)
        L.orc_set_gravity(h, g.ctypes.data_as(_dp))
        for j, (t, p) in enumerate(prob.materials):
            p = _d(p)
            L.orc_add_material(h, MAT[t], p.ctypes.data_as(_dp))
            if len(p) >= 20:  # Voce hardening of Von-Mises (slots 16..19)
                v = _d(p[16:20])
                L.orc_set_material_voce(h, j, v.ctypes.data_as(_dp))
        self.np_ = prob.np_
        # GramsShapeFun (Type=aLME): the field "Beta" is the d x d metric, "Cut_off_Ellipsoid" the neighbour test
        self.alme = bool(s.get("alme", 0))
        if self.alme:
            L.orc_set_shape_alme(h, 1)
        L.orc_set_num_particles(h, self.np_)
        for k, v in prob.fields.items():
            self.set_field(k, v)
        self.set_ints("I0", prob.I0)
        self.set_ints("MatIdx", prob.MatIdx)
        self.lcap = L.orc_list_capacity(h)

    def set_flags(self, quirk_transposed, compute_cep):
        self.L.orc_set_flags(self.h, int(quirk_transposed), int(compute_cep))

    def _fname(self, name):
        return "Beta_tensor" if (self.alme and name == "Beta") else name

    def set_field(self, name, arr):
        a = _d(arr)
        name = self._fname(name)
        assert self.L.orc_set_field(self.h, name.encode(), a.ctypes.data_as(_dp)) == 0, name

    def field(self, name):
        name = self._fname(name)
        c = self.L.orc_field_cols(self.h, name.encode())
        out = np.zeros((self.np_, c))
        self.L.orc_get_field(self.h, name.encode(), out.ctypes.data_as(_dp))
        return out if c > 1 else out[:, 0].copy()

    def set_ints(self, name, arr):
        a = _i(arr)
        self.L.orc_set_ints(self.h, name.encode(), a.ctypes.data_as(_ip))

    def ints(self, name):
        out = np.zeros(self.np_, np.int32)
        self.L.orc_get_ints(self.h, name.encode(), out.ctypes.data_as(_ip))
        return out

    def lists(self):
        out = np.zeros((self.np_, self.lcap), np.int32)
        self.L.orc_get_lists(self.h, out.ctypes.data_as(_ip))
        return out

    def set_lists(self, lists, counts):
        self.set_ints("NumberNodes", counts)
        a = np.full((self.np_, self.lcap), -1, np.int32)
        a[:, :lists.shape[1]] = lists[:, :self.lcap]
        self.L.orc_set_lists(self.h, a.ctypes.data_as(_ip))

    def active(self):
        out = np.zeros(self.prob.nn, np.uint8)
        self.L.orc_get_active(self.h, out.ctypes.data_as(_up))
        return out

    def set_active(self, a):
        a = np.ascontiguousarray(a, dtype=np.uint8)
        self.L.orc_set_active(self.h, a.ctypes.data_as(_up))

    def nodal(self, which):
        out = np.zeros((self.prob.nn, self.d))
        self.L.orc_get_nodal(self.h, which, out.ctypes.data_as(_dp))
        return out

    def set_nodal(self, which, arr):
        a = _d(arr)
        self.L.orc_set_nodal(self.h, which, a.ctypes.data_as(_dp))

    def search_closest(self):
        return self.L.orc_search_closest(self.h)

    def search_lists(self):
        return self.L.orc_search_lists(self.h, None)

    # ---- implicit Newmark-beta restatement (pinned in 2D to the reference's compiled scheme run against oracle/minipetsc)
    def newmark_setup(self, beta=0.25, gamma=0.5, tol=1e-10, max_iter=10, explicit_trial=False):
        self.L.orc_newmark_setup(self.h, ctypes.c_double(beta), ctypes.c_double(gamma), ctypes.c_double(tol),
                                 int(max_iter), int(explicit_trial))

    def lme_point(self, l, lam, beta):
        """Newton for lambda from `lam`, then N and grad N, for one particle given l = x_p - x_a (n x d)."""
        l, lam = _d(l), _d(lam).copy()
        n = l.shape[0]
        N, dN = np.zeros(n), np.zeros((n, self.d))
        st = self.L.orc_lme_point(self.h, n, l.ctypes.data_as(_dp), lam.ctypes.data_as(_dp), ctypes.c_double(beta),
                                  N.ctypes.data_as(_dp), dN.ctypes.data_as(_dp))
        return st, lam, N, dN

    def static_setup(self, tol=1e-10, max_iter=10):
        """U_Static: the implicit loop without inertia (U-Static.c)."""
        self.L.orc_static_setup(self.h, ctypes.c_double(tol), int(max_iter))

    def newmark_begin(self, step):
        return self.L.orc_newmark_begin(self.h, int(step))

    def newmark_begin_after_search(self, step):
        return self.L.orc_newmark_begin_after_search(self.h, int(step))

    def newmark_set(self, which, arr):
        a = _d(arr)
        self.L.orc_newmark_set(self.h, dict(Vn=0, An=1, dU=2)[which], a.ctypes.data_as(_dp))

    def newmark_coeffs(self):
        out = np.zeros(6)
        self.L.orc_newmark_coeffs(out.ctypes.data_as(_dp))
        return out

    def newmark_finish(self):
        return self.L.orc_newmark_finish(self.h)

    def newmark_residual(self, step, dU):
        dU = _d(dU)
        R = np.zeros_like(dU)
        st = self.L.orc_newmark_residual(self.h, int(step), dU.ctypes.data_as(_dp), R.ctypes.data_as(_dp))
        return st, R

    def newmark_tangent(self):
        nd = self.prob.nn * self.d
        K = np.zeros((nd, nd))
        st = self.L.orc_newmark_tangent(self.h, K.ctypes.data_as(_dp))
        return st, K

    def newmark_step(self, step):
        return self.L.orc_newmark_step(self.h, int(step))

    def newmark_iters(self):
        return self.L.orc_newmark_iters()

    def newmark_get(self, which):
        out = np.zeros((self.prob.nn, self.d))
        self.L.orc_newmark_get(self.h, dict(Vn=0, An=1, dU=2, R=3)[which], out.ctypes.data_as(_dp))
        return out

    def fixed(self):
        out = np.zeros((self.prob.nn, self.d), np.uint8)
        self.L.orc_get_fixed(self.h, out.ctypes.data_as(_up))
        return out

    def init_lme(self):
        return self.L.orc_init_lme(self.h)

    def local_search(self):
        it = ctypes.c_int()
        st = self.L.orc_local_search(self.h, ctypes.byref(it))
        return st, it.value

    def stage(self, name, step):
        return self.L.orc_stage(self.h, STAGES[name], step)

    def step(self, k):
        return self.L.orc_step(self.h, k)

    def error(self):
        p = ctypes.c_int()
        c = self.L.orc_error(self.h, ctypes.byref(p))
        return c, p.value

    def shape(self, p):
        N = np.zeros(1024)
        dN = np.zeros(1024 * self.d)
        n = self.L.orc_shape(self.h, p, N.ctypes.data_as(_dp), dN.ctypes.data_as(_dp))
        return N[:n].copy(), dN[:n * self.d].reshape(n, self.d).copy()

    def stress_point(self, p, DF, F_n1, J_n1, b_e_n, eps_n, kappa_n):
        T, d = (5 if self.d == 2 else 9), self.d
        DF, F_n1, b_e_n = _d(DF), _d(F_n1), _d(b_e_n)
        stress, be1, cep = np.zeros(T), np.zeros(T), np.zeros(d * d)
        e1, k1, W = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        st = self.L.orc_stress_point(
            self.h, p, DF.ctypes.data_as(_dp), F_n1.ctypes.data_as(_dp), ctypes.c_double(J_n1),
            b_e_n.ctypes.data_as(_dp), ctypes.c_double(eps_n), ctypes.c_double(kappa_n),
            stress.ctypes.data_as(_dp), be1.ctypes.data_as(_dp), ctypes.byref(e1), ctypes.byref(k1),
            ctypes.byref(W), cep.ctypes.data_as(_dp))
        return dict(status=st, stress=stress, b_e_n1=be1, eps_n1=e1.value, kappa_n1=k1.value,
                    W=W.value, C_ep=cep)


def stiffness_ep(d, u, v, b_e, stress, c_ep):
    """Elastoplastic tangent block of one node pair (restates Elastoplastic-Tangent-Matrix.c:42-160);
    b_e, stress: d x d row-major."""
    out = np.zeros(d * d)
    u, v, b_e, stress, c_ep = _d(u), _d(v), _d(b_e), _d(stress), _d(c_ep)
    lib().orc_stiffness_ep(d, out.ctypes.data_as(_dp), u.ctypes.data_as(_dp), v.ctypes.data_as(_dp),
                           b_e.ctypes.data_as(_dp), stress.ctypes.data_as(_dp), c_ep.ctypes.data_as(_dp))
    return out


def stiffness_nh(d, u, v, un, vn, F_n, J, E, nu):
    """Neo-Hookean tangent block of one node pair (restates Neo-Hookean.c:89-141)."""
    out = np.zeros(d * d)
    u, v, un, vn, F_n = _d(u), _d(v), _d(un), _d(vn), _d(F_n)
    lib().orc_stiffness_nh(d, out.ctypes.data_as(_dp), u.ctypes.data_as(_dp), v.ctypes.data_as(_dp),
                           un.ctypes.data_as(_dp), vn.ctypes.data_as(_dp), F_n.ctypes.data_as(_dp),
                           ctypes.c_double(J), ctypes.c_double(E), ctypes.c_double(nu))
    return out
