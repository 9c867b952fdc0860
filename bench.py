#!/usr/bin/env python
"""bench.py -- particle-updates/s of the NL-PartSol explicit (NPC-FS) hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2|c4] [--scale S]

Default workload: BASELINE.json configs[2] ("c3") -- 3D Neo-Hookean cube compression, 126^3 particle cells x GPxElement 8
= 16,003,008 particles, LME gamma = 6, explicit NPC-FS; at N > 1 the SAME cube is split into N z-slabs (strong scaling,
halo sums + migration over NVLink).  --workload c2: configs[1], 2D Drucker-Prager column, 10^6 particles (weak scaling
over y-slabs); --workload c4: configs[3], 3D Matsuoka-Nakai slope, 8 M particles, slabs + migration (strong scaling);
--workload c5: configs[4], implicit Newmark-beta 3D beam, 2 M particles, device block-CSR tangent + PCG.
A "step" is one full time step over all particles.  Prints ONE JSON line.

`--impl reference` times the reference's CPU implementation of the same path on a bounded sample: the reference's own
compiled 2D translation units (oracle/_ref) for c2, the C port of the reference (oracle/, the 3D reference does not
compile: SURVEY F3) for c3 / c4 -- all host threads.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "nl-partsol_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "particle-updates/sec"
UNIT = "particle-updates/s"

# algorithmic bytes per particle per step (compulsory reads + writes of the per-particle SoA fields).
#   2D plastic: K0 76+4n, K1 112, K2 348, K3 128, K4 168;  3D: K0 100+4n, K1 160, K2 380 (NH) / 556 (plastic), K3 208, K4 248.
# A kernel that performs several stages is charged the SUM of their bytes; a stage that is split over several kernels
# is charged once, against the sum of their times (per_stage).
WORKLOADS = {
    "c2": dict(name="BASELINE configs[1]: 2D granular column collapse, Drucker-Prager, explicit NPC-FS, LME gamma=3, GPxElement 4",
               scaling="weak", step_alg=lambda n: 832 + 4 * n,
               kernel_alg={"lme_p2g_mass_disp": lambda n: 76 + 4 * n + 112.0, "kin_stress_p2g_force": lambda n: 348.0 + 128.0,
                           "g2p_update": lambda n: 168.0},
               stages={"K0+K1 lme+p2g_mass_disp+grid_disp": (("lme_p2g_mass_disp", "grid_disp_bc"), lambda n: 188 + 4 * n),
                       "K2+K3 kin_stress+p2g_force+grid_acc": (("traction", "kin_gather", "stress_update", "kin_stress_p2g_force", "grid_acc"), lambda n: 476.0),
                       "K4 g2p_update": (("g2p_update",), lambda n: 168.0)}),
    "c3": dict(name="BASELINE configs[2]: 3D Neo-Hookean cube compression, explicit NPC-FS, LME gamma=6, GPxElement 8",
               scaling="strong", step_alg=lambda n: 1096 + 4 * n,
               kernel_alg={"lme_p2g_mass_disp": lambda n: 100 + 4 * n + 160.0, "kin_stress_p2g_force": lambda n: 380.0 + 208.0,
                           "g2p_update": lambda n: 248.0},
               stages={"K0+K1 lme+p2g_mass_disp+grid_disp": (("lme_p2g_mass_disp", "grid_disp_bc"), lambda n: 260 + 4 * n),
                       "K2+K3 kin_stress+p2g_force+grid_acc": (("traction", "kin_gather", "stress_update", "kin_stress_p2g_force", "grid_acc"), lambda n: 588.0),
                       "K4 g2p_update": (("g2p_update",), lambda n: 248.0)}),
    "c4": dict(name="BASELINE configs[3]: 3D 45-degree slope, Matsuoka-Nakai (cohesion 1e3), explicit NPC-FS, LME gamma=6, GPxElement 8, gravity ramp",
               scaling="strong", step_alg=lambda n: 1272 + 4 * n,
               # the plastic cloud runs gather -> stress -> force sums: each kernel is charged the stage bytes it moves
               kernel_alg={"lme_p2g_mass_disp": lambda n: 100 + 4 * n + 160.0, "kin_stress_p2g_force": lambda n: 208.0,
                           "kin_gather": lambda n: 56.0 + 72.0, "stress_update": lambda n: 556.0 - 128.0,
                           "g2p_update": lambda n: 248.0},
               stages={"K0+K1 lme+p2g_mass_disp+grid_disp": (("lme_p2g_mass_disp", "grid_disp_bc"), lambda n: 260 + 4 * n),
                       "K2+K3 kin_stress+p2g_force+grid_acc": (("traction", "kin_gather", "stress_update", "kin_stress_p2g_force", "grid_acc"), lambda n: 764.0),
                       "K4 g2p_update": (("g2p_update",), lambda n: 248.0)}),
}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def library_hash():
    """sha256 of the kernel sources: ncu figures stored under profiles/ are only quoted for the build they belong to."""
    h = hashlib.sha256()
    for f in ("nlps_engine.cu", "nlps_cellwarp.cu", "nlps_device.cuh", "nlps_types.cuh"):
        with open(os.path.join(ROOT, "nl-partsol_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def ncu_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu capture of this
    workload -- or None when no capture exists for exactly this build (profiles/ncu_traffic.json: {workload: {hash, sass,
    particles, kernels: {name: bytes}}}).  "Exactly this build" = the MACHINE CODE of the captured kernels: `sass` holds the
    sha256 of the SASS of every kernel the capture covers, the build leaves the same hashes of the library it linked in
    nl-partsol_b200/libnlps_b200.sass.json (nlps_b200/build.py:sass_stamp); a record without `sass` is keyed by the hash of
    the source files instead."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            rec = json.load(f).get(workload)
        if not rec or kernel not in rec.get("kernels", {}):
            return None
        if rec.get("sass"):
            with open(os.path.join(ROOT, "nl-partsol_b200", "libnlps_b200.sass.json")) as f:
                stamp = json.load(f)
            if all(stamp.get(k) == v for k, v in rec["sass"].items()):
                return rec
        elif rec.get("hash") == library_hash():
            return rec
    except Exception:
        pass
    return None


class NvmlSampler:
    """SM clock and clock-event (throttle) reasons read through NVML every few milliseconds during the timed
    region -- the same quantities as the recipe's nvidia-smi line, fine enough for a region of tens of ms."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu=0, period=0.004):
        self.gpu, self.period, self.sm, self.mask, self.ok = gpu, period, [], 0, False
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[gpu]) if vis and vis.split(",")[gpu].isdigit() else gpu
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def start(self):
        self.stop_flag = False
        if not self.ok:
            return
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self.stop_flag = True
        if not self.ok:
            return None
        self.t.join(1.0)
        if not self.sm:
            return None
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max,
                "reasons": sorted(v for k, v in self.REASONS.items() if self.mask & k), "samples": len(self.sm),
                "source": "nvml"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu=0):
        self.rows, self.proc, self.gpu = [], None, gpu

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].startswith("Active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU path on a bounded sample of the workload
def reference_sample(workload, steps, warmup, threads=None):
    threads = threads or os.cpu_count() or 1
    nsteps = steps + warmup
    if workload == "c2":
        return _reference_c2(steps, warmup, threads)
    # 3D: the reference's 3D build does not compile; its C port (oracle/nlps_oracle.c, pinned stage by stage
    # to the reference functions that do compile in 3D) runs the same step on a reduced cloud of the same shape
    import oracle
    from nlps_b200 import synthetic
    if workload == "c3":
        cells = 12 if threads == 1 else 20
        P = synthetic.cube_3d(cells=cells, nsteps=nsteps)
        what = f"3D Neo-Hookean cube, {cells}^3 particle cells x 8"
    else:
        cells, width = (20, 6) if threads == 1 else (32, 10)
        P, _ = synthetic.slope_slab_3d(0, 1, cells=cells, width=width, nsteps=nsteps)
        what = f"3D Matsuoka-Nakai slope, {cells} x {width} x {cells} / 2 particle cells x 8"
    o = oracle.Oracle(P, threads=threads)
    assert o.init_lme() == 0
    for k in range(warmup):
        assert o.step(k) == 0
    t0 = time.perf_counter()
    for k in range(warmup, nsteps):
        assert o.step(k) == 0
    dt = time.perf_counter() - t0
    return dict(value=P.np_ * steps / dt, unit=UNIT, cores=threads, kind="port", ms_per_step=1e3 * dt / steps,
                sample=f"{what} = {P.np_} particles, {steps} steps, C port of the reference (the 3D reference does not "
                       f"compile, SURVEY F3; its setup is O(Nn*Ne): the full size is out of its reach)",
                stage_seconds_last_step=None)


def _reference_c2(steps, warmup, threads, cells=(64, 128)):
    """oracle/_ref (the reference's own 2D translation units driven by the restated step loop) on cells[0] x cells[1]
    particle cells x 4 particles of the C2 deck; falls back to the C port where oracle/_ref was not built."""
    so = os.path.join(ROOT, "oracle", "_ref", "libnlps2d_ref.so")
    bx, by = cells if threads > 1 else (32, 64)
    nsteps = steps + warmup
    if os.path.exists(so):
        import deckgen
        import refharness
        spec = deckgen.DeckSpec(nx=bx * 6, ny=by + by // 4, h=0.2 / bx, pnx=bx, pny=by, ph=0.2 / bx,
                                porigin=(0.0, 0.0), nsteps=nsteps, cfl=0.5, cel=(1e7 / 2000.0) ** 0.5 * 1.3)
        spec.material = deckgen.Material("Drucker-Prager", {
            "rho": 2000.0, "E": 1e7, "nu": 0.3, "m": 1.0, "Hardening-modulus": 1.0,
            "Reference-plastic-strain": 1e-2, "kappa-0": 1e4, "Friction-angle": 30.0, "Dilatancy-angle": 0.0})
        tmp = tempfile.mkdtemp(prefix="nlps_bench_ref_")
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(1)
        os.dup2(devnull, 1)  # the reference parser is chatty on stdout
        try:
            h = refharness.RefHarness(deckgen.write_deck(spec, tmp), threads=threads)
        finally:
            import ctypes
            ctypes.CDLL(None).fflush(None)
            os.dup2(saved, 1)
        npart = h.np_
        for k in range(warmup):
            assert h.step(k) == 0
        t0 = time.perf_counter()
        for k in range(warmup, nsteps):
            assert h.step(k) == 0
        dt = time.perf_counter() - t0
        kind = "reference"
        stage = h.stage_times().tolist()
    else:
        import oracle
        from nlps_b200 import synthetic
        sys.stderr.write("bench: oracle/_ref missing, timing the C port instead\n")
        P = synthetic.column_collapse_2d(scale=bx / 354.0, nsteps=nsteps)
        o = oracle.Oracle(P, threads=threads)
        assert o.init_lme() == 0
        npart = P.np_
        for k in range(warmup):
            assert o.step(k) == 0
        t0 = time.perf_counter()
        for k in range(warmup, nsteps):
            assert o.step(k) == 0
        dt = time.perf_counter() - t0
        kind = "port"
        stage = None
    return dict(value=npart * steps / dt, unit=UNIT, cores=threads, kind=kind, ms_per_step=1e3 * dt / steps,
                sample=f"2D DP column, {bx}x{by} particle cells x4 = {npart} particles, {steps} steps "
                       f"(reference setup is O(Nn*Ne): full 10^6 size is out of its reach)",
                stage_seconds_last_step=stage)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = reference_sample(args.workload, args.steps, args.warmup, threads=args.threads or None)
    wl = WORKLOADS[args.workload]
    line = {"metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": wl["scaling"],
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": wl["name"] + " (bounded CPU sample; throughput is per particle)", "sample": cb["sample"]},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
def build_workload(name, rank, world, scale, nsteps, synthetic):
    """(Problem of this rank, slab dict or None, particles of the whole job, extra config)."""
    if name == "c4":
        cells, width = max(16 * world, int(round(160 * scale))), max(8, int(round(78 * scale)))
        # band of 5 cells: a slab must be wider than two bands, and the quantile cuts make the slabs at the tall end of
        # the wedge thin; a particle's closest node may then sit one node layer beyond the cut between two migrations
        P, slab = synthetic.slope_slab_3d(rank, world, cells=cells, width=width, nsteps=nsteps, band_cells=5)
        if world == 1:
            return P, None, P.np_, dict(cells=cells, width=width)
        total = slab["n_particles"]
        return P, {k: v for k, v in slab.items() if k != "n_particles"}, total, dict(cells=cells, width=width)
    if name == "c3":
        cells = max(8 * world, int(round(126 * scale)))
        if world == 1:
            P = synthetic.cube_3d(cells=cells, nsteps=nsteps)
            return P, None, P.np_, dict(cells=cells)
        P, slab = synthetic.cube_slab_3d(rank, world, cells=cells, nsteps=nsteps)
        return P, slab, slab["n_global"], dict(cells=cells)
    if world == 1:
        P = synthetic.column_collapse_2d(scale=scale, nsteps=nsteps)
        return P, None, P.np_, {}
    # weak scaling over spatial slabs: the column is `world` times taller, one slab of by rows per GPU; every rank
    # builds only its sub-mesh and particles
    P, slab = synthetic.column_slab_2d(rank, world, scale=scale, nsteps=nsteps)
    return P, slab, slab["n_global"], {}


def parity_small(name, rank, world, local, comm, engine, synthetic):
    """N > 1: the slab data plane that has just been timed (same transport, same kernels) against ONE engine holding
    the whole problem, on a reduced cloud of the same workload with a velocity kick so that particles cross the cuts.
    Every rank compares the particles it ends up holding.  Returns (max relative error over the compared fields,
    integer outputs identical, particles that migrated in on this rank)."""
    from util import field_scales
    nsteps = 16
    if name == "c3":
        cz, xy = 12 * world, 8
        G = synthetic.cube_3d(cells=cz, nsteps=nsteps, xy=xy)
        Pr, sl = synthetic.cube_slab_3d(rank, world, cells=cz, nsteps=nsteps, xy=xy, band_cells=4)
        for Q in (G, Pr):
            Q.fields["vel"][:, 2] = -0.15 * Q.solver["cel"]
        gid_G = np.arange(G.np_)
    elif name == "c4":
        cells, width = 16 * world, 4
        G, _ = synthetic.slope_slab_3d(0, 1, cells=cells, width=width, nsteps=nsteps, ramp_steps=8, band_cells=5,
                                       material=synthetic.NH_C1)
        Pr, sl = synthetic.slope_slab_3d(rank, world, cells=cells, width=width, nsteps=nsteps, ramp_steps=8, band_cells=5,
                                         material=synthetic.NH_C1)
        sl = {k: v for k, v in sl.items() if k != "n_particles"}
        for Q in (G, Pr):
            Q.fields["vel"][:, 2] = -0.15 * Q.solver["cel"] * np.clip((Q.fields["x_GC"][:, 2] - 0.2) / 0.4, 0.0, 1.0)
        gid_G = (G.kept_cells.astype(np.int64)[:, None] * 8 + np.arange(8)[None, :]).ravel()
    else:
        scale = 0.04
        bx = max(4, int(round(354 * scale)))
        by = 2 * bx
        G = synthetic.structured_problem(2, (6 * bx, by * world + by // 4), 0.2 / bx, (bx, by * world), (0, 0),
                                         synthetic.DP_C2, nsteps, 0.5, (1e7 / 2000.0) ** 0.5 * 1.3, (0.0, -9.81))
        Pr, sl = synthetic.column_slab_2d(rank, world, scale=scale, nsteps=nsteps)
        for Q in (G, Pr):
            Q.fields["vel"][:, 1] = -0.12 * Q.solver["cel"]
        gid_G = np.arange(G.np_)
    eng = engine.Engine(Pr, device=local, slab=dict(sl, comm=comm, migrate_every=3))
    assert eng.initialize_lme() == 0, eng.error()
    assert eng.run(0, nsteps) == 0, eng.error()
    f, ids = eng.download_local()
    moved = eng.migrated_count()
    eng.close()
    e1 = engine.Engine(G, device=local)
    assert e1.initialize_lme() == 0 and e1.run(0, nsteps) == 0, e1.error()
    f1 = e1.download()
    e1.close()
    rows = np.searchsorted(gid_G, ids)
    assert np.array_equal(gid_G[rows], ids)
    ints_ok = bool(np.array_equal(f["I0"] + sl["node_offset"], f1["I0"][rows]) and
                   np.array_equal(f["NumberNodes"], f1["NumberNodes"][rows]))
    sc = field_scales(G)
    worst = 0.0
    for k in ("x_GC", "dis", "D_dis", "vel", "acc", "F_n", "DF", "Stress", "rho", "J_n", "W", "b_e_n", "EPS_n", "Kappa_n",
              "lambda", "Beta"):
        a, b = np.asarray(f[k], float), np.asarray(f1[k][rows], float)
        s = max(float(np.abs(b).max()) if b.size else 0.0, sc.get(k) or 0.0, 1e-300)
        worst = max(worst, float(np.abs(a - b).max()) / s if a.size else 0.0)
    return worst, ints_ok, moved


def run_ours(args):
    import torch
    import torch.distributed as dist

    from nlps_b200 import engine, synthetic
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "NONE"  # VERSION and WARN both print "NCCL version ..." to stdout: keep it to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    name = args.workload
    wl = WORKLOADS[name]
    K, Wm = args.steps, max(args.warmup, 3)
    trace = int(os.environ.get("NLPS_BENCH_TRACE", "0"))  # diagnostic: per-step wall times of `trace` extra steps on stderr
    nsteps_total = Wm + K + 2 + trace
    t_setup = time.perf_counter()
    comm = None
    P, slab, total_particles, extra = build_workload(name, rank, world, args.scale, nsteps_total, synthetic)
    if world > 1:
        comm = engine.NcclComm(rank, world, local)
        slab = dict(slab, comm=comm, migrate_every=10)
    eng = engine.Engine(P, device=local, slab=slab)
    assert eng.initialize_lme() == 0, eng.error()
    setup_s = time.perf_counter() - t_setup
    npart = eng.local_count() if world > 1 else P.np_
    npart_max = npart
    if world > 1:
        tn = torch.tensor([npart], dtype=torch.int64, device="cuda")
        dist.all_reduce(tn, op=dist.ReduceOp.MAX)
        npart_max = int(tn.item())
    # warm-up
    assert eng.run(0, Wm) == 0, eng.error()
    sampler, sampler2 = NvmlSampler(local), ClockSampler(local)
    sampler2.start()
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = eng.launch_count()
    rc, ms = eng.timed_run(Wm, K)
    assert rc == 0, eng.error()
    torch.cuda.synchronize()
    launches = eng.launch_count() - l0
    clocks = sampler.stop()
    clocks2 = sampler2.stop()
    if clocks is None:
        clocks = clocks2
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = total_particles * K / (ms_max * 1e-3)

    # neighbours per particle (NumberNodes of the particles this rank holds)
    n_avg = float(eng.download_local()[0]["NumberNodes"].mean()) if world > 1 else float(eng.download()["NumberNodes"].mean())
    # per-kernel device times (CUDA events on the engine's stream, serialised per launch)
    eng.profile(True)
    eng.kernel_times(reset=True)
    assert eng.run(Wm + K, 2) == 0
    kt = eng.kernel_times()
    eng.profile(False)
    if trace:
        per = []
        for k in range(trace):
            t0 = time.perf_counter()
            assert eng.run(Wm + K + 2 + k, 1) == 0
            torch.cuda.synchronize()
            per.append(round((time.perf_counter() - t0) * 1e3, 2))
        print(f"[trace rank {rank}] steps {Wm + K + 2}..: {per}", file=sys.stderr)
    transport = eng.transport() if world > 1 else "none (single GPU)"
    peak, peak_src = measured_peak()
    per_kernel = {}
    for kname, (kms, kn) in kt.items():
        if kn == 0:
            continue
        avg = kms / kn
        d = {"ms": round(avg, 4), "launches_per_step": round(kn / 2, 2)}
        if kname in wl["kernel_alg"]:
            gbs = wl["kernel_alg"][kname](n_avg) * npart / (avg * 1e-3) / 1e9
            d.update(alg_gbs=round(gbs, 1), frac=round(gbs / peak, 4))
        per_kernel[kname] = d
    groups = {}
    for gname, (members, fn) in wl["stages"].items():
        tms = sum(per_kernel[k]["ms"] * per_kernel[k]["launches_per_step"] for k in members if k in per_kernel)
        if tms > 0:
            gbs = fn(n_avg) * npart / (tms * 1e-3) / 1e9
            groups[gname] = {"ms": round(tms, 4), "alg_gbs": round(gbs, 1), "frac": round(gbs / peak, 4)}
    dom = max((k for k in per_kernel if "alg_gbs" in per_kernel[k]), key=lambda k: per_kernel[k]["ms"])
    step_bytes = wl["step_alg"](n_avg) * total_particles
    traffic, traffic_note = None, "no ncu capture of this workload for this build (profiles/ncu_traffic.json keyed by the source hash)"
    rec = ncu_traffic(name, dom) if world == 1 else None
    if rec and abs(rec.get("particles", 0) - npart) <= 0:
        traffic = round(rec["kernels"][dom] / (per_kernel[dom]["ms"] * 1e-3) / 1e9, 1)
        traffic_note = f"ncu dram bytes per launch ({rec.get('source')}) / live launch duration"
    roofline = {"bound": "hbm", "kernel": dom, "achieved": per_kernel[dom]["alg_gbs"], "peak": peak,
                "unit": "GB/s", "frac": per_kernel[dom]["frac"], "traffic": traffic, "traffic_note": traffic_note,
                "fp64_roof": {"dfma_per_s": 1.711e13, "source": "profiles/fp64_peak.cu (58.8 DFMA / clk / SM at 1965 MHz)"},
                "peak_source": peak_src,
                "step_achieved_gbs": round(step_bytes * K / (ms_max * 1e-3) / 1e9, 1),
                "step_frac": round(step_bytes * K / (ms_max * 1e-3) / 1e9 / (peak * world), 4),
                "neighbours_per_particle": round(n_avg, 2), "per_kernel": per_kernel, "per_stage": groups,
                "build": library_hash()}
    eng.close()

    # ---- N > 1: parity of the data plane that was just timed, on a reduced cloud of the same workload
    parity = None
    if world > 1:
        worst, ints_ok, moved = parity_small(name, rank, world, local, comm, engine, synthetic)
        tv = torch.tensor([worst, 0.0 if ints_ok else 1.0, float(moved)], dtype=torch.float64, device="cuda")
        tmax = tv.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = tv.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ok = float(tmax[0].item()) <= 1e-10 and float(tmax[1].item()) == 0.0
        parity = {"parity_n": "ok" if ok else "FAILED", "max_rel_err": float(tmax[0].item()),
                  "integers_identical": float(tmax[1].item()) == 0.0, "particles_migrated": int(tsum[2].item()),
                  "against": "one engine holding the whole (reduced) cloud, 16 steps, migration every 3 steps, tolerance 1e-10",
                  "transport": transport}
        if not ok:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "value": None, "error": "slab parity check failed", "parity": parity}))
            sys.exit(3)

    e2e = None
    if not args.no_e2e and name != "c4":
        e2e = measure_e2e(args, name, rank, world, local, comm, total_particles, K, engine, synthetic, torch, dist)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(name)
    if rank == 0:
        multi = "single GPU"
        if world > 1:
            multi = (f"{wl['scaling']} scaling over {world} spatial slabs "
                     + ("along z of the fixed problem" if name != "c2" else f"along y (column {world}x taller)")
                     + f": halo sums of occupancy / mass+momentum / forces every step over {transport}, "
                       "particle migration every 10 steps (ncclSend/ncclRecv)")
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
                "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": dict({"workload": wl["name"], "particles": total_particles, "particles_per_gpu": npart,
                                "particles_max_rank": npart_max, "background_nodes": P.nn, "scale": args.scale,
                                "l2": "inputs larger than L2 (particle state of one GPU >> 126 MB)" if npart * 400 > 2e8
                                      else "particle state near L2 size: reduced --scale run, not a bench value",
                                "multi_gpu": multi, "transport": transport, "setup_seconds": round(setup_s, 2),
                                "plastic_3d_eigenvector_form": ("column form (quirk_transposed_eigvec = 0): deliberate deviation from the "
                                                                "reference's compiled row form, DESIGN.md section 6, deviation 3") if name == "c4" else None},
                               **extra),
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu, "parity": parity}
        print(json.dumps(line))
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(name):
    """The reference arm on a bounded sample, all host threads and one thread (reported baseline only, never gating)."""
    out = None
    for label, extra in (("all", []), ("one", ["--threads", "1"])):
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", name,
                                "--steps", "3", "--warmup", "1"] + extra, capture_output=True, text=True, timeout=600)
            cb = json.loads(r.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as ex:
            cb = {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": f"failed: {ex}"}
        if label == "all":
            out = cb
        else:
            out["one_thread"] = {k: cb.get(k) for k in ("value", "cores", "sample")}
    return out


def measure_e2e(args, name, rank, world, local, comm, total_particles, K, engine, synthetic, torch, dist):
    """The same metric through the scheme call with HOST buffers (create + H2D of mesh and state, steps, D2H of every
    field at each results step, destroy); host wall clock, max over ranks.  The FIRST call of the process is reported
    as `cold` (empty memory pool, no peer-buffer cache), the median of the following ones as the warm value."""
    c3 = name == "c3"
    e2e_steps = max(K, 200) if not c3 else max(K, 40)   # the scheme call amortises its set-up over the run
    every = 50 if not c3 else 20
    P2, slab2, _, _ = build_workload(name, rank, world, args.scale, e2e_steps, synthetic)
    if world > 1:
        slab2 = dict(slab2, comm=comm, migrate_every=10)
    # initialise lambda / beta once (setup, as the driver does with initialise_shapefun__MeshTools__ before the scheme)
    eng0 = engine.Engine(P2, device=local, slab=slab2)
    assert eng0.initialize_lme() == 0
    if world == 1:
        f0 = eng0.download()
        for k in ("lambda", "Beta"):
            P2.fields[k] = f0[k]
    else:
        f0, ids0 = eng0.download_local()
        order = np.argsort(slab2["global_id"])
        rows = order[np.searchsorted(slab2["global_id"][order], ids0)]
        for k in ("lambda", "Beta"):
            P2.fields[k][rows] = f0[k]
    eng0.close()
    del f0

    # the caller's buffers live in pinned host memory (bench contract): mesh tables and every particle field
    def _pin(a):
        a = np.asarray(a)
        a = np.ascontiguousarray(a, dtype=np.int32 if a.dtype.kind in "iub" else np.float64)
        return torch.from_numpy(a).pin_memory().numpy()
    if not os.environ.get("NLPS_BENCH_PAGEABLE"):
        for nm in ("coords", "r1p", "r1i", "r2p", "r2i", "h_avg", "I0", "MatIdx"):
            setattr(P2, nm, _pin(getattr(P2, nm)))
        P2.fields = {k: _pin(v) for k, v in P2.fields.items()}
    mesh_bytes = sum(a.nbytes for a in (P2.coords, P2.r1p, P2.r1i, P2.r2p, P2.r2i, P2.h_avg))
    state_bytes = sum(v.nbytes for v in P2.fields.values()) + P2.I0.nbytes + P2.MatIdx.nbytes
    # the scheme call works in place on the caller's buffers: keep the initial state to repeat the measurement
    reps = 3 if (world == 1 and not c3) else 2
    init = {k: v.copy() for k, v in P2.fields.items()} if (reps > 1 and state_bytes < 4e9) else None
    if init is None:
        reps = 1
    init_I0 = P2.I0.copy()
    engine.lib().nlps_b200_trim(local)  # the first call starts from an empty pool, like a fresh process
    samples = []
    for rep in range(reps):
        if rep > 0:
            for k, v in init.items():
                P2.fields[k][...] = v
            P2.I0[...] = init_I0
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        engine.u_verlet(P2, run_initialize=False, results_every=every, device=local, slab=slab2, inplace=True)
        samples.append(time.perf_counter() - t0)
    cold_s = samples[0]
    warm = sorted(samples[1:]) if len(samples) > 1 else [samples[0]]
    e2e_s = warm[len(warm) // 2]
    te = torch.tensor([e2e_s, cold_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s, cold_s = float(te[0].item()), float(te[1].item())
    n_dl = sum(1 for k in range(e2e_steps) if k % every == 0) + 1
    return {"value": total_particles * e2e_steps / e2e_s, "unit": UNIT,
            "h2d_bytes_per_step": int(world * (mesh_bytes + state_bytes) / e2e_steps),
            "d2h_bytes_per_step": int(world * state_bytes * n_dl / e2e_steps),
            "steps": e2e_steps, "results_every": every, "seconds": round(e2e_s, 4),
            "cold": {"value": total_particles * e2e_steps / cold_s, "seconds": round(cold_s, 4),
                     "what": "first scheme call after nlps_b200_trim(): empty memory pool, no cached peer buffers"},
            "seconds_all_calls": [round(x, 4) for x in samples],
            "host_memory": "pageable" if os.environ.get("NLPS_BENCH_PAGEABLE") else "pinned",
            "call": f"nlps_b200_u_verlet[_slab] (create + H2D of mesh and state, {e2e_steps} steps, D2H of all fields every {every} "
                    "steps overlapped with the following steps, destroy), host wall clock"}


def run_c5(args):
    """--workload c5: BASELINE configs[4], implicit Newmark-beta finite-strain 3D beam (8 x 1 x 1, Neo-Hookean, LME gamma 6,
    dt = 10 x the explicit limit), 2,097,152 particles, device block-CSR tangent + Jacobi-PCG.  A step = one converged
    time step (Newton to TOL 1e-10).  N > 1: slabs along the beam (strong scaling)."""
    import torch
    import torch.distributed as dist

    from nlps_b200 import engine, synthetic
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "NONE")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    K, Wm = args.steps, max(1, min(args.warmup, 2))
    c = max(4, int(round(32 * args.scale)))
    t0 = time.perf_counter()
    P = synthetic.beam_3d(cells_per_unit=c, nsteps=Wm + K + 1)
    # N > 1: the beam is cut into N slabs along its axis (strong scaling); every slab assembles the tangent of its own
    # particles, the Krylov vectors are summed over the band nodes and the dot products over the slabs
    comm, slab = None, None
    if world > 1:
        comm = engine.NcclComm(rank, world, local)
        axis, cuts = engine.slab_cuts(P, world)
        slab = dict(rank=rank, world=world, axis=axis, cuts=cuts, comm=comm, migrate_every=10)
    eng = engine.Engine(P, device=local, slab=slab)
    assert eng.initialize_lme() == 0
    assert eng.newmark_setup(tol=1e-10, max_iter=10, pcg_rtol=1e-6) == 0
    setup_s = time.perf_counter() - t0
    for k in range(Wm):
        assert eng.newmark_step(k) == 0, eng.error()
    s0 = eng.newmark_stats()
    sampler = NvmlSampler(local)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ta = time.perf_counter()
    newton = []
    for k in range(Wm, Wm + K):
        assert eng.newmark_step(k) == 0, eng.error()     # every step ends with a device synchronisation (Newton's test)
        newton.append(eng.newmark_stats()["newton_iters"])
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - ta)
    clocks = sampler.stop()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    s1 = eng.newmark_stats()
    pcg = s1["pcg_iters_total"] - s0["pcg_iters_total"]
    asm = s1["assemblies_total"] - s0["assemblies_total"]
    ms_pcg_iter = (s1["ms_pcg"] - s0["ms_pcg"]) / max(1, pcg)
    # block-CSR SpMV: 9 doubles + one 4-byte column index per 3x3 block, plus the vectors of one PCG iteration
    spmv_bytes = s1["nnz_blocks"] * (72 + 4) + 5 * 3 * s1["n_rows"] * 8
    peak, peak_src = measured_peak()
    gbs = spmv_bytes / 1e9 / (ms_pcg_iter * 1e-3)
    transport = eng.transport() if world > 1 else "none (single GPU)"
    f1, ids1 = eng.download_local() if world > 1 else (eng.download(), None)
    eng.close()
    # end to end with host buffers: create + H2D of mesh and state, the same steps, D2H of every field at the end
    P2 = synthetic.beam_3d(cells_per_unit=c, nsteps=Wm + K + 1)
    state_bytes = sum(v.nbytes for v in P2.fields.values())
    mesh_bytes = sum(a.nbytes for a in (P2.coords, P2.r1p, P2.r1i, P2.r2p, P2.r2i, P2.h_avg))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    tb = time.perf_counter()
    e2 = engine.Engine(P2, device=local, slab=slab)
    assert e2.initialize_lme() == 0 and e2.newmark_setup(tol=1e-10, max_iter=10, pcg_rtol=1e-6) == 0
    for k in range(Wm + K):
        assert e2.newmark_step(k) == 0
    f2, ids2 = e2.download_local() if world > 1 else (e2.download(), None)
    e2e_s = time.perf_counter() - tb
    e2.close()
    # the two runs are the same computation (the tangent is assembled with floating-point atomics and the Krylov solve
    # stops at a tolerance: equal to solver accuracy, not bit for bit)
    xa = f1["x_GC"] if ids1 is None else f1["x_GC"][np.argsort(ids1)]
    xb = f2["x_GC"] if ids2 is None else f2["x_GC"][np.argsort(ids2)]
    assert xa.shape == xb.shape and np.abs(xa - xb).max() <= 1e-7 * np.abs(xa).max(), float(np.abs(xa - xb).max())
    parity = None
    if world > 1:
        te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te.item())
        # parity of the slab data plane: a reduced beam on the slabs against one engine holding all of it
        Ps = synthetic.beam_3d(cells_per_unit=max(6, (13 * world + 7) // 8), nsteps=5)  # a slab must span two halo bands
        ax, cu = engine.slab_cuts(Ps, world)
        es = engine.Engine(Ps, device=local, slab=dict(rank=rank, world=world, axis=ax, cuts=cu, comm=comm, migrate_every=2))
        assert es.initialize_lme() == 0 and es.newmark_setup(tol=1e-12, max_iter=25, pcg_rtol=1e-13) == 0
        assert es.newmark_run(0, 4) == 0, es.error()
        fs, ids = es.download_local()
        es.close()
        e1 = engine.Engine(Ps, device=local)
        assert e1.initialize_lme() == 0 and e1.newmark_setup(tol=1e-12, max_iter=25, pcg_rtol=1e-13) == 0
        assert e1.newmark_run(0, 4) == 0, e1.error()
        fw = e1.download()
        e1.close()
        worst = 0.0
        for kf in ("x_GC", "vel", "Stress", "F_n"):
            den = max(float(np.abs(fw[kf]).max()), 1e-300)
            worst = max(worst, float(np.abs(fs[kf] - fw[kf][ids]).max()) / den)
        tv = torch.tensor([worst], dtype=torch.float64, device="cuda")
        dist.all_reduce(tv, op=dist.ReduceOp.MAX)
        worst = float(tv.item())
        parity = {"parity_n": "ok" if worst <= 1e-8 else "FAILED", "max_rel_err": worst, "transport": transport,
                  "against": "one engine holding the whole (reduced) beam, 4 converged steps, migration every 2, tolerance 1e-8 "
                             "(Newton stopped at 1e-12 |R0|)"}
        if worst > 1e-8:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "value": None, "error": "slab parity check failed", "parity": parity}))
            sys.exit(3)
    if rank == 0:
        value = P.np_ * K / (ms_max * 1e-3)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
                "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": "BASELINE configs[4]: implicit Newmark-beta finite-strain 3D beam (Neo-Hookean, LME gamma=6, "
                                       "dt = 10 x explicit limit), device block-CSR tangent + Jacobi-PCG replacing PETSc KSP",
                           "particles": P.np_, "background_nodes": P.nn, "scale": args.scale,
                           "multi_gpu": "single GPU" if world == 1 else
                           (f"strong scaling over {world} slabs along the beam: per PCG iteration the band sums of K p over {transport}, "
                            "two all-reduces of the dot-product partials; particle migration every 10 steps"),
                           "transport": transport,
                           "newton_iters_per_step": newton, "pcg_iters": int(pcg), "block_rows": s1["n_rows"],
                           "nnz_blocks": s1["nnz_blocks"], "setup_seconds": round(setup_s, 2),
                           "ms_assemble_per_newton": round((s1["ms_assemble"] - s0["ms_assemble"]) / max(1, asm), 3),
                           "ms_per_pcg_iter": round(ms_pcg_iter, 4), "l2": "tangent (7 GB) larger than L2"},
                "clocks": clocks,
                "e2e": {"value": P.np_ * (Wm + K) / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int((mesh_bytes + state_bytes) / (Wm + K)),
                        "d2h_bytes_per_step": int(state_bytes / (Wm + K)), "steps": Wm + K, "seconds": round(e2e_s, 3),
                        "call": "create (H2D of mesh and state) + initialize + newmark steps + download of every field, host wall clock"},
                "gpu_launches": None,
                "roofline": {"bound": "hbm", "kernel": "k_bsr_spmv (one PCG iteration: SpMV + vector updates)", "achieved": round(gbs, 1),
                             "peak": peak, "unit": "GB/s", "frac": round(gbs / peak, 4), "traffic": None, "peak_source": peak_src,
                             "bytes_per_iteration": int(spmv_bytes),
                             "note": "76 bytes per 3x3 block (72 values + one column index) + 5 vectors of 3 x rows doubles"},
                "cpu_baseline": None, "parity": parity}
        print(json.dumps(line))
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--scale", type=float, default=1.0, help="linear scale of the workload (1.0 = the BASELINE size)")
    ap.add_argument("--threads", type=int, default=0, help="--impl reference: host threads (0 = all)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer scheme call")
    ap.add_argument("--workload", default="c3", choices=("c2", "c3", "c4", "c5"),
                    help="c3 (default, the driver's bench line): BASELINE configs[2], 3D cube, 16 M particles, strong scaling; "
                         "c2: configs[1], 2D column, weak scaling; c4: configs[3], 3D Matsuoka-Nakai slope, slabs + migration "
                         "(no e2e leg)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c5":
        run_c5(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
