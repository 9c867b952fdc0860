"""`bench.py --impl reference` (the CPU arm the driver runs beside the GPU line): contract of its JSON line, alone and under
torchrun with two ranks (rank 0 alone works and prints, the other exits 0).  CPU only, small step counts."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
REF2D = os.path.join(ROOT, "oracle", "_ref", "libnlps2d_ref.so")


def _lines(cmd):
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900,
                         env=dict(os.environ, OMP_NUM_THREADS="2"))
    assert out.returncode == 0, out.stderr[-3000:]
    return [json.loads(ln) for ln in out.stdout.splitlines() if ln.startswith("{")]


def _check(line, n_gpus, steps, warmup):
    assert line["impl"] == "reference" and "unavailable" not in line
    assert line["metric"] == "particle-updates/sec" and line["unit"] == "particle-updates/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["ms_per_step"] > 0
    assert (line["n_gpus"], line["steps"], line["warmup"]) == (n_gpus, steps, warmup)
    assert line["dtype"] == "f64" and line["data"] == "synthetic" and line["vs_baseline"] is None
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == line["value"]
    e = line["e2e"]
    assert e["value"] == line["value"] and e["unit"] == line["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and line.get("gpu_launches", 0) == 0


def test_reference_arm_default_workload():
    """BASELINE configs[2] (3D): the reference's 3D build does not compile, so the arm is the C port, labelled as such."""
    (line,) = _lines([sys.executable, "bench.py", "--impl", "reference", "--steps", "2", "--warmup", "1"])
    _check(line, 1, 2, 1)
    assert line["cpu_baseline"]["kind"] == "port" and "configs[2]" in line["config"]["workload"]


def test_reference_arm_2d_uses_the_compiled_reference():
    if not os.path.exists(REF2D):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    # (--threads 1: the one-thread sample is 8,192 particles -- the reference's set-up is quadratic)
    (line,) = _lines([sys.executable, "bench.py", "--impl", "reference", "--workload", "c2", "--steps", "2", "--warmup", "1",
                      "--threads", "1"])
    _check(line, 1, 2, 1)
    assert line["cpu_baseline"]["cores"] == 1
    assert line["cpu_baseline"]["kind"] == "reference" and "configs[1]" in line["config"]["workload"]


def test_reference_arm_under_torchrun_prints_one_line():
    lines = _lines([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                    "127.0.0.1", "--master-port", "29777", "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "2",
                    "--warmup", "1"])
    assert len(lines) == 1
    _check(lines[0], 2, 2, 1)
