"""Plain-array description of one NL-PartSol `-u` explicit problem.

This is the host-side mirror of what the reference's setup hands to a scheme
function (`Mesh`, `Particle`, `Time_Int_Params` + the process globals of
`Globals.h`, see `Types.h:548-865`), flattened to numpy arrays: linked-list
adjacency (`NodalLocality_0`, `NodalLocality`) becomes CSR in CHAIN order,
`Load` tables become dense (dim x NumTimeStep) arrays.
"""
from __future__ import annotations

import io
from dataclasses import dataclass, field

import numpy as np

MATERIAL_TYPES = {"Neo-Hookean-Wriggers": 0, "Drucker-Prager": 1, "Matsuoka-Nakai": 2, "Von-Mises": 3, "Hencky": 4, "Lade-Duncan": 5}
# order of the material parameter block (oracle/ref_harness.c refh_material_params; slots 16..19 = the Voce hardening
# parameters of Von-Mises, refh_material_voce; blocks of 16 are padded with their defaults theta = 1, 0, 0, 0)
MATERIAL_SLOTS = ("rho", "E", "nu", "ReferencePressure", "kappa_0", "Hardening_modulus",
                  "Plastic_Strain_0", "phi_Frictional", "psi_Frictional", "Exponent_Hardening_Ortiz",
                  "Cohesion", "alpha_Hardening_Borja", "a1", "a2", "a3", "J2_degradated",
                  "theta_Hardening_Voce", "K_0_Hardening_Voce", "K_inf_Hardening_Voce", "delta_Hardening_Voce")


def material_params(p):
    """The 20-slot block of a material given 16 or 20 numbers."""
    p = [float(v) for v in p]
    if len(p) < 20:
        p = (p + [0.0] * 16)[:16] + [1.0, 0.0, 0.0, 0.0]
    return np.array(p[:20])

VECTOR_FIELDS = ("x_GC", "dis", "D_dis", "vel", "acc", "lambda")
TENSOR_FIELDS = ("F_n", "F_n1", "DF", "b_e_n", "b_e_n1", "Stress")
SCALAR_FIELDS = ("J_n", "J_n1", "mass", "rho", "Vol_0", "W", "EPS_n", "EPS_n1", "Kappa_n",
                 "Kappa_n1", "Beta")
ALL_FIELDS = VECTOR_FIELDS + TENSOR_FIELDS + SCALAR_FIELDS + ("C_ep",)


@dataclass
class Problem:
    ndim: int
    coords: np.ndarray
    r1p: np.ndarray
    r1i: np.ndarray
    r2p: np.ndarray
    r2i: np.ndarray
    h_avg: np.ndarray
    dx: float
    solver: dict
    gravity: np.ndarray                      # (ndim, nsteps)
    bounds: list = field(default_factory=list)   # dict(nodes, dir, val)
    neumann: list = field(default_factory=list)  # dict(nodes=particle ids, dir, val)
    materials: list = field(default_factory=list)  # (type string, params[16])
    fields: dict = field(default_factory=dict)    # name -> float64 array
    I0: np.ndarray | None = None
    MatIdx: np.ndarray | None = None
    conn: np.ndarray | None = None

    @property
    def nn(self):
        return int(self.coords.shape[0])

    @property
    def np_(self):
        return int(self.fields["x_GC"].shape[0])

    @property
    def T(self):
        return 5 if self.ndim == 2 else 9

    @property
    def nsteps(self):
        return int(self.solver["nsteps"])

    def dt(self):
        return self.solver["cfl"] * self.dx / self.solver["cel"]

    # ---- defaults for a freshly seeded particle set (allocate_U_vars__Fields__, U-Analisys.c:5-170)
    def init_fields(self, x, vol, matidx):
        n, d, T = x.shape[0], self.ndim, self.T
        f = {}
        for k in VECTOR_FIELDS:
            f[k] = np.zeros((n, d))
        ident = np.zeros(T)
        for i in range(d):
            ident[i * d + i] = 1.0
        if d == 2:
            ident[4] = 1.0
        for k in TENSOR_FIELDS:
            f[k] = np.tile(ident, (n, 1)) if k != "Stress" else np.zeros((n, T))
        for k in SCALAR_FIELDS:
            f[k] = np.zeros(n)
        f["C_ep"] = np.zeros((n, d * d))
        f["x_GC"] = np.ascontiguousarray(x, dtype=np.float64)
        f["J_n"][:] = 1.0
        f["J_n1"][:] = 1.0
        f["Vol_0"] = np.ascontiguousarray(vol, dtype=np.float64)
        self.MatIdx = np.ascontiguousarray(matidx, dtype=np.int32)
        rho = np.array([m[1][0] for m in self.materials])[self.MatIdx]
        f["rho"] = rho.copy()
        f["mass"] = rho * f["Vol_0"]
        # Generate-One-Phase-Analysis.c:621-626
        f["Kappa_n"] = np.array([m[1][4] for m in self.materials])[self.MatIdx].copy()
        is_mn = np.array([m[0] == "Matsuoka-Nakai" for m in self.materials])[self.MatIdx]
        eps0 = np.array([m[1][6] for m in self.materials])[self.MatIdx]
        f["EPS_n"] = np.where(is_mn, eps0, 0.0)
        if any(m[0] == "Von-Mises" for m in self.materials):
            f["Back_stress"] = np.zeros((n, 3))  # Phi.Back_stress (U-Analisys.c:152), principal components
        self.fields = f

    # ---- npz (golden fixtures)
    def to_npz_dict(self):
        out = dict(ndim=self.ndim, coords=self.coords, r1p=self.r1p, r1i=self.r1i, r2p=self.r2p,
                   r2i=self.r2i, h_avg=self.h_avg, dx=self.dx, gravity=self.gravity,
                   I0=self.I0, MatIdx=self.MatIdx,
                   solver_keys=np.array(list(self.solver.keys())),
                   solver_vals=np.array([float(v) for v in self.solver.values()]),
                   n_bounds=len(self.bounds), n_neumann=len(self.neumann),
                   mat_types=np.array([m[0] for m in self.materials]),
                   mat_params=np.array([material_params(m[1]) for m in self.materials]))
        if self.conn is not None:
            out["conn"] = self.conn
        for i, b in enumerate(self.bounds):
            for k in ("nodes", "dir", "val"):
                out[f"bound{i}_{k}"] = b[k]
        for i, b in enumerate(self.neumann):
            for k in ("nodes", "dir", "val"):
                out[f"neumann{i}_{k}"] = b[k]
        for k, v in self.fields.items():
            out["f_" + k] = v
        return out

    @staticmethod
    def from_npz(z):
        solver = {str(k): float(v) for k, v in zip(z["solver_keys"], z["solver_vals"])}
        for k in ("nsteps", "max_iter_lme", "maxiter_radial"):
            solver[k] = int(solver[k])
        p = Problem(ndim=int(z["ndim"]), coords=z["coords"], r1p=z["r1p"], r1i=z["r1i"], r2p=z["r2p"],
                    r2i=z["r2i"], h_avg=z["h_avg"], dx=float(z["dx"]), solver=solver,
                    gravity=z["gravity"])
        p.I0 = z["I0"]
        p.MatIdx = z["MatIdx"]
        p.conn = z["conn"] if "conn" in z else None
        p.bounds = [dict(nodes=z[f"bound{i}_nodes"], dir=z[f"bound{i}_dir"], val=z[f"bound{i}_val"])
                    for i in range(int(z["n_bounds"]))]
        p.neumann = [dict(nodes=z[f"neumann{i}_nodes"], dir=z[f"neumann{i}_dir"],
                          val=z[f"neumann{i}_val"]) for i in range(int(z["n_neumann"]))]
        p.materials = [(str(t), np.array(q)) for t, q in zip(z["mat_types"], z["mat_params"])]
        p.fields = {k[2:]: z[k] for k in z.files if k.startswith("f_")}
        return p

    def copy(self):
        buf = io.BytesIO()
        np.savez(buf, **self.to_npz_dict())
        buf.seek(0)
        return Problem.from_npz(np.load(buf))
