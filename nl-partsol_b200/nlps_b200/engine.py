"""ctypes binding of include/nlps_b200.h."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build
from .problem import MATERIAL_TYPES, Problem, material_params

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

STAGES = dict(search=0, p2g_mass_disp=1, grid_disp=2, kin_stress=3, force=4, grid_acc=5, g2p=6)


class Mesh(C.Structure):
    _fields_ = [("ndim", C.c_int), ("n_nodes", C.c_int), ("coords", _dp), ("ring1_ptr", _ip),
                ("ring1_idx", _ip), ("ring2_ptr", _ip), ("ring2_idx", _ip), ("h_avg", _dp),
                ("delta_x", C.c_double)]


class Load(C.Structure):
    _fields_ = [("n_ids", C.c_int), ("dim", C.c_int), ("ids", _ip), ("dir", _ip), ("val", _dp)]


class Material(C.Structure):
    _fields_ = [("type", C.c_int), ("rho", C.c_double), ("E", C.c_double), ("nu", C.c_double),
                ("reference_pressure", C.c_double), ("kappa_0", C.c_double),
                ("hardening_modulus", C.c_double), ("plastic_strain_0", C.c_double),
                ("phi_frictional", C.c_double), ("psi_frictional", C.c_double),
                ("exponent_hardening_ortiz", C.c_double), ("cohesion", C.c_double),
                ("alpha_hardening_borja", C.c_double), ("a_hardening_borja", C.c_double * 3),
                ("theta_hardening_voce", C.c_double), ("k_0_hardening_voce", C.c_double),
                ("k_inf_hardening_voce", C.c_double), ("delta_hardening_voce", C.c_double)]


class Solver(C.Structure):
    _fields_ = [("cfl", C.c_double), ("cel", C.c_double), ("initial_step", C.c_int),
                ("num_steps", C.c_int), ("gamma_lme", C.c_double), ("tol_zero_lme", C.c_double),
                ("tol_wrapper_lme", C.c_double), ("max_iter_lme", C.c_int),
                ("tol_radial_returning", C.c_double), ("max_iter_radial_returning", C.c_int),
                ("thickness", C.c_double), ("quirk_transposed_eigvec", C.c_int), ("compute_c_ep", C.c_int),
                ("shape_function", C.c_int)]


_PFIELDS = ("x_GC", "dis", "D_dis", "vel", "acc", "F_n", "F_n1", "DF", "b_e_n", "b_e_n1", "Stress", "C_ep",
            "J_n", "J_n1", "mass", "rho", "Vol_0", "W", "EPS_n", "EPS_n1", "Kappa_n", "Kappa_n1", "lambda",
            "Beta")


class Particles(C.Structure):
    _fields_ = [("n", C.c_int)] + [(k, _dp) for k in _PFIELDS] + [("I0", _ip), ("NumberNodes", _ip),
                                                                  ("MatIdx", _ip), ("Area_0", _dp),
                                                                  ("Back_stress", _dp), ("Cut_off_Ellipsoid", _dp)]


class Msg(C.Structure):
    _fields_ = [("peer", C.c_int), ("send", C.c_void_p), ("send_bytes", C.c_ulonglong), ("recv", C.c_void_p),
                ("recv_bytes", C.c_ulonglong)]


class Slab(C.Structure):
    _fields_ = [("rank", C.c_int), ("world", C.c_int), ("axis", C.c_int), ("cuts", _dp), ("band_cells", C.c_int),
                ("migrate_every", C.c_int), ("capacity_factor", C.c_double), ("n_global", C.c_int),
                ("global_id", _ip), ("node_id_offset", C.c_int), ("comm", C.c_void_p)]


class Newmark(C.Structure):
    _fields_ = [("beta", C.c_double), ("gamma", C.c_double), ("tol", C.c_double), ("max_iter", C.c_int),
                ("use_explicit_trial", C.c_int), ("pcg_rtol", C.c_double), ("pcg_max_iter", C.c_int),
                ("quasi_static", C.c_int)]


class NewmarkStats(C.Structure):
    _fields_ = [("newton_iters", C.c_int), ("n_rows", C.c_int), ("nnz_blocks", C.c_int),
                ("pcg_iters_total", C.c_longlong), ("assemblies_total", C.c_longlong),
                ("residual_evals_total", C.c_longlong), ("residual0", C.c_double), ("residual", C.c_double),
                ("ms_assemble", C.c_double), ("ms_pcg", C.c_double), ("ms_residual", C.c_double)]


EXCHANGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(Msg), C.c_void_p)

_lib = None


def lib():
    """Load libnlps_b200.so, building it in-tree if needed.  Raises if it cannot be loaded:
    the product path has no fallback."""
    global _lib
    if _lib is None:
        so = os.environ.get("NLPS_LIB") or _build.build()  # NLPS_LIB: A/B a differently built library
        L = C.CDLL(so)
        L.nlps_b200_create.restype = C.c_void_p
        L.nlps_b200_dt.restype = C.c_double
        L.nlps_b200_version.restype = C.c_char_p
        L.nlps_b200_launch_count.restype = C.c_longlong
        L.nlps_b200_create_slab.restype = C.c_void_p
        L.nlps_b200_comm_create_nccl.restype = C.c_void_p
        L.nlps_b200_comm_create_custom.restype = C.c_void_p
        L.nlps_b200_comm_create_custom.argtypes = [C.c_int, C.c_int, EXCHANGE_FN, C.c_void_p]
        L.nlps_b200_comm_destroy.argtypes = [C.c_void_p]
        L.nlps_b200_memcpy_d2d.argtypes = [C.c_void_p, C.c_void_p, C.c_ulonglong, C.c_void_p]
        L.nlps_b200_stream_sync.argtypes = [C.c_void_p]
        L.nlps_b200_migrated_count.restype = C.c_longlong
        L.nlps_b200_transport.restype = C.c_char_p
        _lib = L
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _material(t, p):
    m = Material()
    m.type = MATERIAL_TYPES[t]
    (m.rho, m.E, m.nu, m.reference_pressure, m.kappa_0, m.hardening_modulus, m.plastic_strain_0,
     m.phi_frictional, m.psi_frictional, m.exponent_hardening_ortiz, m.cohesion,
     m.alpha_hardening_borja) = [float(v) for v in p[:12]]
    for k in range(3):
        m.a_hardening_borja[k] = float(p[12 + k])
    q = material_params(p)
    m.theta_hardening_voce, m.k_0_hardening_voce, m.k_inf_hardening_voce, m.delta_hardening_voce = [float(v) for v in q[16:20]]
    return m


def build_locality(ndim, coords, conn):
    """nlps_b200_build_locality: scalable adjacency with the reference's chain orders."""
    L = lib()
    coords, conn = _d(coords), _i(conn)
    nn, (ne, nne) = coords.shape[0], conn.shape
    r1p, r2p = np.zeros(nn + 1, np.int32), np.zeros(nn + 1, np.int32)
    h_avg = np.zeros(nn)
    dx = C.c_double()
    args = (ndim, nn, ne, nne, conn.ctypes.data_as(_ip), coords.ctypes.data_as(_dp))
    assert L.nlps_b200_build_locality(*args, r1p.ctypes.data_as(_ip), None, r2p.ctypes.data_as(_ip), None,
                                      h_avg.ctypes.data_as(_dp), C.byref(dx)) == 0
    r1i, r2i = np.zeros(r1p[-1], np.int32), np.zeros(r2p[-1], np.int32)
    assert L.nlps_b200_build_locality(*args, r1p.ctypes.data_as(_ip), r1i.ctypes.data_as(_ip),
                                      r2p.ctypes.data_as(_ip), r2i.ctypes.data_as(_ip),
                                      h_avg.ctypes.data_as(_dp), C.byref(dx)) == 0
    return r1p, r1i, r2p, r2i, h_avg, dx.value


def _mesh_struct(prob: Problem):
    arrs = dict(coords=_d(prob.coords), r1p=_i(prob.r1p), r1i=_i(prob.r1i), r2p=_i(prob.r2p),
                r2i=_i(prob.r2i), h=_d(prob.h_avg))
    m = Mesh()
    m.ndim, m.n_nodes = prob.ndim, prob.nn
    m.coords = arrs["coords"].ctypes.data_as(_dp)
    m.ring1_ptr, m.ring1_idx = arrs["r1p"].ctypes.data_as(_ip), arrs["r1i"].ctypes.data_as(_ip)
    m.ring2_ptr, m.ring2_idx = arrs["r2p"].ctypes.data_as(_ip), arrs["r2i"].ctypes.data_as(_ip)
    m.h_avg, m.delta_x = arrs["h"].ctypes.data_as(_dp), float(prob.dx)
    return m, arrs


# ---- spatial slabs: host planning, identical on every rank ------------------------
def slab_cuts(prob: Problem, world, axis=-1, I0=None):
    """nlps_b200_slab_cuts: (axis, cuts[world-1]) balancing the particle counts."""
    m, keep = _mesh_struct(prob)
    I0 = _i(prob.I0 if I0 is None else I0)
    ax = C.c_int()
    cuts = np.zeros(max(world - 1, 1))
    rc = lib().nlps_b200_slab_cuts(C.byref(m), len(I0), I0.ctypes.data_as(_ip), int(world), int(axis), C.byref(ax),
                                   cuts.ctypes.data_as(_dp))
    if rc != 0:
        raise RuntimeError(f"nlps_b200_slab_cuts failed ({rc})")
    return ax.value, cuts[:world - 1].copy()


def slab_owner(prob: Problem, axis, cuts, I0=None):
    """Owner slab of every particle (vectorised twin of nlps_b200_slab_owner)."""
    I0 = prob.I0 if I0 is None else I0
    return np.searchsorted(np.asarray(cuts), prob.coords[I0, axis], side="right").astype(np.int32)


def slab_halo_nodes(prob: Problem, axis, cut, band_cells=6):
    m, keep = _mesh_struct(prob)
    L = lib()
    n = L.nlps_b200_slab_halo_nodes(C.byref(m), int(axis), C.c_double(cut), int(band_cells), None)
    ids = np.zeros(max(n, 1), np.int32)
    L.nlps_b200_slab_halo_nodes(C.byref(m), int(axis), C.c_double(cut), int(band_cells), ids.ctypes.data_as(_ip))
    return ids[:n]


class NcclComm:
    """NCCL transport; the id travels through torch.distributed (any backend) -- plumbing only."""

    def __init__(self, rank, world, device):
        import torch
        import torch.distributed as dist
        L = lib()
        buf = C.create_string_buffer(128)
        if rank == 0:
            assert L.nlps_b200_comm_unique_id(buf) == 0
        t = torch.tensor(list(buf.raw), dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            t = t.cuda(device)
        dist.broadcast(t, 0)
        raw = bytes(t.cpu().tolist())
        self.h = C.c_void_p(L.nlps_b200_comm_create_nccl(raw, int(rank), int(world), int(device)))
        if not self.h:
            raise RuntimeError("nlps_b200_comm_create_nccl failed")

    def close(self):
        if self.h:
            lib().nlps_b200_comm_destroy(self.h)
            self.h = None


class ThreadComm:
    """Loopback transport for several slab engines inside ONE process (one thread per slab, any
    number of GPUs -- also all on the same one): device-to-device copies behind a barrier.
    Used by the single-GPU tests of the multi-slab path."""

    class _Shared:
        def __init__(self, world):
            import threading
            self.world = world
            self.barrier = threading.Barrier(world)
            self.box = {}

    @staticmethod
    def group(world):
        sh = ThreadComm._Shared(world)
        return [ThreadComm(sh, r) for r in range(world)]

    def __init__(self, shared, rank):
        L = lib()
        self.sh, self.rank = shared, rank

        def xchg(user, n, msgs, stream):
            try:
                sh = self.sh
                L.nlps_b200_stream_sync(stream)
                for k in range(n):
                    sh.box[(rank, msgs[k].peer)] = (msgs[k].send, msgs[k].send_bytes)
                sh.barrier.wait(timeout=120)
                for k in range(n):
                    src, nb = sh.box[(msgs[k].peer, rank)]
                    if nb != msgs[k].recv_bytes:
                        return 1
                    if L.nlps_b200_memcpy_d2d(msgs[k].recv, src, nb, stream) != 0:
                        return 1
                L.nlps_b200_stream_sync(stream)
                sh.barrier.wait(timeout=120)
                return 0
            except Exception:  # a broken barrier must not hang the other slabs
                try:
                    self.sh.barrier.abort()
                except Exception:
                    pass
                return 1

        self._cb = EXCHANGE_FN(xchg)
        self.h = C.c_void_p(L.nlps_b200_comm_create_custom(rank, shared.world, self._cb, None))

    def close(self):
        if self.h:
            lib().nlps_b200_comm_destroy(self.h)
            self.h = None


class _Marshal:
    """Keeps the numpy buffers alive that the C structs point into."""

    def __init__(self, prob: Problem, quirk=-1, compute_c_ep=0, initial_step=0, copy=True):
        self.keep = []
        k = self.keep
        self.mesh = Mesh()
        arrs = dict(coords=_d(prob.coords), r1p=_i(prob.r1p), r1i=_i(prob.r1i), r2p=_i(prob.r2p),
                    r2i=_i(prob.r2i), h=_d(prob.h_avg))
        k.append(arrs)
        m = self.mesh
        m.ndim, m.n_nodes = prob.ndim, prob.nn
        m.coords = arrs["coords"].ctypes.data_as(_dp)
        m.ring1_ptr, m.ring1_idx = arrs["r1p"].ctypes.data_as(_ip), arrs["r1i"].ctypes.data_as(_ip)
        m.ring2_ptr, m.ring2_idx = arrs["r2p"].ctypes.data_as(_ip), arrs["r2i"].ctypes.data_as(_ip)
        m.h_avg, m.delta_x = arrs["h"].ctypes.data_as(_dp), float(prob.dx)
        s = prob.solver
        self.solver = Solver(float(s["cfl"]), float(s["cel"]), int(initial_step), int(s["nsteps"]),
                             float(s["gamma_lme"]), float(s["tol_zero"]), float(s["tol_wrapper"]),
                             int(s["max_iter_lme"]), float(s["tol_radial"]), int(s["maxiter_radial"]),
                             float(s.get("thickness", 1.0)), int(quirk), int(compute_c_ep),
                             1 if s.get("alme", 0) else 0)   # GramsShapeFun (Type=aLME)
        self.bounds = self._loads(prob.bounds)
        self.neumann = self._loads(prob.neumann)
        self.gravity = _d(prob.gravity) if prob.gravity is not None else None
        self.materials = (Material * len(prob.materials))(*[_material(t, p) for t, p in prob.materials])
        self.state, self.host = self.particles(prob, copy)

    def _loads(self, lst):
        arr = (Load * max(len(lst), 1))()
        for j, b in enumerate(lst):
            ids, di, v = _i(b["nodes"]), _i(b["dir"]), _d(b["val"])
            self.keep.append((ids, di, v))
            arr[j] = Load(len(ids), di.shape[0], ids.ctypes.data_as(_ip), di.ctypes.data_as(_ip),
                          v.ctypes.data_as(_dp))
        return arr

    @staticmethod
    def particles(prob: Problem, copy=True):
        """copy=False: the C side reads and writes the Problem's own arrays, as the reference's scheme functions
        do with the driver's buffers (arrays must already be contiguous float64 / int32)."""
        cp = (lambda a: a.copy()) if copy else (lambda a: a)
        host = {k: cp(_d(prob.fields[k])) for k in _PFIELDS if k in prob.fields}
        host["I0"] = cp(_i(prob.I0))
        host["MatIdx"] = cp(_i(prob.MatIdx))
        host["NumberNodes"] = np.zeros(prob.np_, np.int32)
        st = Particles()
        st.n = prob.np_
        for kname in _PFIELDS:
            setattr(st, kname, host[kname].ctypes.data_as(_dp) if kname in host else None)
        for kname in ("I0", "MatIdx", "NumberNodes"):
            setattr(st, kname, host[kname].ctypes.data_as(_ip))
        if "Area_0" in prob.fields:  # 3D Neumann loads (Phi.Area_0); constant, never downloaded
            host["Area_0"] = cp(_d(prob.fields["Area_0"]))
            st.Area_0 = host["Area_0"].ctypes.data_as(_dp)
        if "Back_stress" in prob.fields:  # Von-Mises kinematic hardening (Phi.Back_stress, n x 3)
            host["Back_stress"] = cp(_d(prob.fields["Back_stress"]))
            st.Back_stress = host["Back_stress"].ctypes.data_as(_dp)
        if "Cut_off_Ellipsoid" in prob.fields:  # aLME: the metric of the neighbour test (n x d*d); Beta is n x d*d too
            host["Cut_off_Ellipsoid"] = cp(_d(prob.fields["Cut_off_Ellipsoid"]))
            st.Cut_off_Ellipsoid = host["Cut_off_Ellipsoid"].ctypes.data_as(_dp)
        return st, host


class Engine:
    """Device-resident explicit NPC-FS engine (nlps_b200_create .. destroy)."""

    def __init__(self, prob: Problem, device=0, quirk=-1, compute_c_ep=0, slab=None):
        """slab: None, or dict(rank, world, axis, cuts, comm[, band_cells, migrate_every, capacity_factor,
        global_id, n_global]) -- this engine then keeps the particles of `prob` its slab owns."""
        L = lib()
        self.L = L
        self.prob = prob
        self.m = _Marshal(prob, quirk, compute_c_ep)
        err = C.create_string_buffer(256)
        m = self.m
        args = (C.byref(m.mesh), C.byref(m.solver), len(prob.bounds), m.bounds, len(prob.neumann), m.neumann,
                m.gravity.ctypes.data_as(_dp) if m.gravity is not None else None, len(prob.materials), m.materials,
                C.byref(m.state))
        self.slab = None
        if slab is None:
            self.h = L.nlps_b200_create(*args, device, err, 256)
        else:
            self.slab = make_slab(slab, prob.np_, m.keep)
            self.h = L.nlps_b200_create_slab(*args, C.byref(self.slab), device, err, 256)
        if not self.h:
            raise RuntimeError("nlps_b200_create failed: " + err.value.decode())
        self.h = C.c_void_p(self.h)
        self.d, self.np_, self.nn = prob.ndim, prob.np_, prob.nn
        self.compact = slab is not None and slab.get("global_id") is not None

    def close(self):
        if getattr(self, "h", None):
            self.L.nlps_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def initialize_lme(self):
        return self.L.nlps_b200_initialize_lme(self.h)

    def step(self, k):
        return self.L.nlps_b200_step(self.h, int(k))

    def run(self, first, count):
        return self.L.nlps_b200_run(self.h, int(first), int(count))

    def timed_run(self, first, count):
        ms = C.c_double()
        rc = self.L.nlps_b200_timed_run(self.h, int(first), int(count), C.byref(ms))
        return rc, ms.value

    def stage(self, name, k):
        return self.L.nlps_b200_stage(self.h, STAGES[name], int(k))

    def download(self):
        """All particle fields as a dict of numpy arrays (host layout of the reference)."""
        st, host = self.m.state, self.m.host
        assert self.L.nlps_b200_download(self.h, C.byref(st)) == 0
        out = {}
        for k, v in host.items():
            out[k] = v.copy()
        return out

    def local_count(self):
        return int(self.L.nlps_b200_local_count(self.h))

    def download_local(self):
        """Compact rows of the particles this slab holds + their global ids."""
        n = self.local_count()
        host = {k: np.zeros((max(n, 1),) + v.shape[1:], v.dtype) for k, v in self.m.host.items() if k != "Area_0"}
        st = Particles()
        st.n = max(n, 1)
        for k in _PFIELDS:
            setattr(st, k, host[k].ctypes.data_as(_dp) if k in host else None)
        for k in ("I0", "MatIdx", "NumberNodes"):
            setattr(st, k, host[k].ctypes.data_as(_ip))
        if "Back_stress" in host:
            st.Back_stress = host["Back_stress"].ctypes.data_as(_dp)
        if "Cut_off_Ellipsoid" in host:
            st.Cut_off_Ellipsoid = host["Cut_off_Ellipsoid"].ctypes.data_as(_dp)
        ids = np.zeros(max(n, 1), np.int32)
        assert self.L.nlps_b200_download_local(self.h, C.byref(st), ids.ctypes.data_as(_ip)) == 0
        return {k: v[:n] for k, v in host.items()}, ids[:n].copy()

    def migrate(self):
        return self.L.nlps_b200_migrate(self.h)

    def transport(self):
        """The data plane of the per-step halo sums (peer-memory stores, ncclSend/ncclRecv, custom, none)."""
        return self.L.nlps_b200_transport(self.h).decode()

    def migrated_count(self):
        return int(self.L.nlps_b200_migrated_count(self.h))

    # ---- implicit Newmark-beta (nlps_b200_newmark_*)
    def newmark_setup(self, beta=0.25, gamma=0.5, tol=1e-10, max_iter=10, explicit_trial=False, pcg_rtol=0.0,
                      pcg_max_iter=0, quasi_static=False):
        prm = Newmark(beta, gamma, tol, int(max_iter), int(explicit_trial), pcg_rtol, int(pcg_max_iter),
                      int(quasi_static))
        return self.L.nlps_b200_newmark_setup(self.h, C.byref(prm))

    def run_async(self, first, count):
        return self.L.nlps_b200_run_async(self.h, int(first), int(count))

    def sync(self):
        return self.L.nlps_b200_sync(self.h)

    def download_begin(self):
        """Snapshot of the current step; returns the host dict that download_end() completes."""
        if self.L.nlps_b200_download_begin(self.h, C.byref(self.m.state)) != 0:
            raise RuntimeError("nlps_b200_download_begin failed")
        return self.m.host

    def download_end(self):
        if self.L.nlps_b200_download_end(self.h) != 0:
            raise RuntimeError("nlps_b200_download_end failed")
        return {k: v.copy() for k, v in self.m.host.items()}

    def newmark_step(self, k):
        return self.L.nlps_b200_newmark_step(self.h, int(k))

    def newmark_run(self, first, count):
        return self.L.nlps_b200_newmark_run(self.h, int(first), int(count))

    def newmark_stats(self):
        st = NewmarkStats()
        assert self.L.nlps_b200_newmark_stats(self.h, C.byref(st)) == 0
        return {k: getattr(st, k) for k, _ in NewmarkStats._fields_}

    def newmark_begin(self, k):
        return self.L.nlps_b200_newmark_begin(self.h, int(k))

    def newmark_get(self, which):
        out = np.zeros((self.nn, self.d))
        assert self.L.nlps_b200_newmark_get(self.h, dict(Vn=0, An=1, dU=2, R=3)[which], out.ctypes.data_as(_dp)) == 0
        return out

    def newmark_residual(self, k, dU=None):
        R = np.zeros((self.nn, self.d))
        du = _d(dU) if dU is not None else None
        rc = self.L.nlps_b200_newmark_residual(self.h, int(k), du.ctypes.data_as(_dp) if du is not None else None,
                                               R.ctypes.data_as(_dp))
        return rc, R

    def newmark_tangent(self):
        """(row_nodes, row_ptr, col_nodes, vals[nnz, d, d]) of the tangent at the last residual state."""
        nr, nz = C.c_int(), C.c_int()
        assert self.L.nlps_b200_newmark_tangent(self.h, C.byref(nr), C.byref(nz), None, None, None, None) == 0
        rows = np.zeros(nr.value, np.int32)
        rp = np.zeros(nr.value + 1, np.int32)
        cols = np.zeros(max(nz.value, 1), np.int32)
        vals = np.zeros((max(nz.value, 1), self.d, self.d))
        assert self.L.nlps_b200_newmark_tangent(self.h, C.byref(nr), C.byref(nz), rows.ctypes.data_as(_ip),
                                                rp.ctypes.data_as(_ip), cols.ctypes.data_as(_ip),
                                                vals.ctypes.data_as(_dp)) == 0
        return rows, rp, cols[:nz.value], vals[:nz.value]

    def upload(self, fields: dict):
        st, host = self.m.state, self.m.host
        for k, v in fields.items():
            if k in host:
                host[k][...] = np.asarray(v).reshape(host[k].shape)
        assert self.L.nlps_b200_upload(self.h, C.byref(st)) == 0

    def nodal(self, which):
        out = np.zeros((self.nn, self.d))
        assert self.L.nlps_b200_get_nodal(self.h, which, out.ctypes.data_as(_dp)) == 0
        return out

    def active(self):
        out = np.zeros(self.nn, np.uint8)
        assert self.L.nlps_b200_get_active(self.h, out.ctypes.data_as(C.POINTER(C.c_ubyte))) == 0
        return out

    def lists(self):
        cap = self.L.nlps_b200_list_capacity(self.h)
        counts = np.zeros(self.np_, np.int32)
        lists = np.zeros((self.np_, cap), np.int32)
        assert self.L.nlps_b200_get_lists(self.h, counts.ctypes.data_as(_ip), lists.ctypes.data_as(_ip), cap) == 0
        return counts, lists

    def error(self):
        c, p = C.c_int(), C.c_int()
        self.L.nlps_b200_last_error(self.h, C.byref(c), C.byref(p))
        return c.value, p.value

    def dt(self):
        return self.L.nlps_b200_dt(self.h)

    def profile(self, on=True):
        self.L.nlps_b200_profile(self.h, int(on))

    def kernel_times(self, reset=False):
        names = (C.c_char_p * 32)()
        ms = (C.c_double * 32)()
        n = (C.c_int * 32)()
        k = self.L.nlps_b200_kernel_times(self.h, 32, names, ms, n)
        out = {names[i].decode(): (ms[i], n[i]) for i in range(k)}
        if reset:
            self.L.nlps_b200_reset_kernel_times(self.h)
        return out

    def launch_count(self):
        return int(self.L.nlps_b200_launch_count(self.h))


def make_slab(slab: dict, n_state, keep: list):
    cuts = _d(slab.get("cuts", np.zeros(0)))
    gid = slab.get("global_id")
    gid = _i(gid) if gid is not None else None
    keep.append((cuts, gid))
    comm = slab.get("comm")
    return Slab(int(slab["rank"]), int(slab["world"]), int(slab["axis"]),
                cuts.ctypes.data_as(_dp) if len(cuts) else None, int(slab.get("band_cells", 0)),
                int(slab.get("migrate_every", 0)), float(slab.get("capacity_factor", 0.0)),
                int(slab.get("n_global", n_state)), gid.ctypes.data_as(_ip) if gid is not None else None,
                int(slab.get("node_offset", 0)), comm.h if comm is not None else None)


RESULTS_CB = C.CFUNCTYPE(None, C.c_int, C.c_void_p)


def u_verlet(prob: Problem, run_initialize=False, results_every=0, device=0, quirk=-1, initial_step=0, slab=None,
             inplace=False, callback=None):
    """The whole scheme call with HOST buffers (nlps_b200_u_verlet[_slab]).  Returns the final fields
    (slab engines: rows of other slabs keep their input values; with global_id: compact rows).
    callback(step, fields): called for every results step with the host buffers holding that step."""
    L = lib()
    m = _Marshal(prob, quirk, 0, initial_step, copy=not inplace)
    cb = RESULTS_CB(lambda k, _u: callback(k, m.host)) if callback is not None else None
    args = (C.byref(m.mesh), C.byref(m.solver), len(prob.bounds), m.bounds, len(prob.neumann),
            m.neumann, m.gravity.ctypes.data_as(_dp) if m.gravity is not None else None,
            len(prob.materials), m.materials, C.byref(m.state))
    if slab is None:
        rc = L.nlps_b200_u_verlet(*args, int(run_initialize), int(results_every), cb, None, device)
    else:
        sl = make_slab(slab, prob.np_, m.keep)
        ids = np.zeros(max(prob.np_, 1), np.int32)
        rc = L.nlps_b200_u_verlet_slab(*args, C.byref(sl), ids.ctypes.data_as(_ip), int(run_initialize),
                                       int(results_every), cb, None, device)
    if rc != 0:
        raise RuntimeError("nlps_b200_u_verlet failed")
    n = m.state.n
    out = {k: (v[:n] if inplace else v[:n].copy()) for k, v in m.host.items()}
    if slab is not None and slab.get("global_id") is not None:
        out["_ids"] = ids[:n].copy()
    return out


def stress_points(ndim, mat_type, mat_params, tol_radial, maxiter_radial, DF, F_n1, J_n1, b_e_n, eps_n, kappa_n,
                  quirk=-1, device=0):
    """nlps_b200_stress_points: the constitutive update on arrays of material points."""
    L = lib()
    DF, F_n1, J_n1, b_e_n, eps_n, kappa_n = (_d(a) for a in (DF, F_n1, J_n1, b_e_n, eps_n, kappa_n))
    n, T, dd = DF.shape[0], DF.shape[1], ndim * ndim
    out = dict(stress=np.zeros((n, T)), b_e_n1=np.zeros((n, T)), eps_n1=np.zeros(n), kappa_n1=np.zeros(n),
               W=np.zeros(n), C_ep=np.zeros((n, dd)), status=np.zeros(n, np.int32))
    m = _material(mat_type, mat_params)
    rc = L.nlps_b200_stress_points(ndim, C.byref(m), C.c_double(tol_radial), int(maxiter_radial), int(quirk), n,
                                   DF.ctypes.data_as(_dp), F_n1.ctypes.data_as(_dp), J_n1.ctypes.data_as(_dp),
                                   b_e_n.ctypes.data_as(_dp), eps_n.ctypes.data_as(_dp), kappa_n.ctypes.data_as(_dp),
                                   out["stress"].ctypes.data_as(_dp), out["b_e_n1"].ctypes.data_as(_dp),
                                   out["eps_n1"].ctypes.data_as(_dp), out["kappa_n1"].ctypes.data_as(_dp),
                                   out["W"].ctypes.data_as(_dp), out["C_ep"].ctypes.data_as(_dp),
                                   out["status"].ctypes.data_as(_ip), device)
    if rc != 0:
        raise RuntimeError("nlps_b200_stress_points failed")
    return out
