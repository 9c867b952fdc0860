"""oracle/mini_lapack.c against a real LAPACK (OpenBLAS bundled in the opencv wheel) when present."""
import ctypes
import glob
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
OB_DIR = "/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs"
dp = ctypes.POINTER(ctypes.c_double)
ip = ctypes.POINTER(ctypes.c_int)


def _mini():
    so = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])
    return ctypes.CDLL(so)


def _openblas():
    libs = glob.glob(OB_DIR + "/libopenblas*")
    if not libs:
        pytest.skip("no OpenBLAS in this image")
    for f in glob.glob(OB_DIR + "/libquadmath*") + glob.glob(OB_DIR + "/libgfortran*"):
        ctypes.CDLL(f, mode=ctypes.RTLD_GLOBAL)
    return ctypes.CDLL(libs[0])


def _dsyev(lib, A):
    a = np.array(A, dtype=np.float64, order="C")
    w = np.zeros(a.shape[0])
    info = lib.LAPACKE_dsyev(101, ctypes.c_char(b"V"), ctypes.c_char(b"U"), a.shape[0],
                             a.ctypes.data_as(dp), a.shape[0], w.ctypes.data_as(dp))
    assert info == 0
    return w, a


def test_dsyev_2x2_bit_exact_including_signs():
    ob, mi = _openblas(), _mini()
    rng = np.random.default_rng(1)
    for t in range(20000):
        kind = t % 4
        if kind == 0:
            F = np.eye(2) + 0.3 * rng.standard_normal((2, 2))
            A = F @ F.T
        elif kind == 1:
            A = rng.standard_normal((2, 2))
            A = A + A.T
        elif kind == 2:
            A = np.diag(rng.uniform(0.5, 2, 2))
            A[0, 1] = A[1, 0] = rng.standard_normal() * 10.0 ** rng.integers(-20, 0)
        else:
            A = np.eye(2) * rng.uniform(0.9, 1.1)
            A[0, 0] += rng.standard_normal() * 1e-12
            A[0, 1] = A[1, 0] = rng.standard_normal() * 1e-13
        w1, z1 = _dsyev(ob, A)
        w2, z2 = _dsyev(mi, A)
        assert np.array_equal(w1, w2) and np.array_equal(z1, z2), A


def test_lu_solve_5x5_matches_lapack():
    ob, mi = _openblas(), _mini()
    rng = np.random.default_rng(2)
    for _ in range(500):
        A = rng.standard_normal((5, 5))
        b = rng.standard_normal(5)
        outs = []
        for lib in (ob, mi):
            a, x, piv = A.copy(), b.copy(), np.zeros(5, np.int32)
            assert lib.LAPACKE_dgetrf(101, 5, 5, a.ctypes.data_as(dp), 5, piv.ctypes.data_as(ip)) == 0
            assert lib.LAPACKE_dgetrs(101, ctypes.c_char(b"N"), 5, 1, a.ctypes.data_as(dp), 5,
                                      piv.ctypes.data_as(ip), x.ctypes.data_as(dp), 1) == 0
            outs.append((x, piv))
        assert np.array_equal(outs[0][1], outs[1][1])
        assert np.abs(outs[0][0] - outs[1][0]).max() <= 1e-12 * max(1.0, np.abs(outs[0][0]).max()) * np.linalg.cond(A)


def test_jacobi_3x3_eigen():
    mi = _mini()
    rng = np.random.default_rng(3)
    for _ in range(200):
        F = np.eye(3) + 0.3 * rng.standard_normal((3, 3))
        A = F @ F.T
        w, z = _dsyev(mi, A)
        assert np.all(np.diff(w) >= 0)
        assert np.abs(z @ np.diag(w) @ z.T - A).max() < 1e-13 * np.abs(A).max()
        assert np.abs(z.T @ z - np.eye(3)).max() < 1e-14
