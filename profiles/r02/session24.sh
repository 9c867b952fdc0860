#!/bin/bash
# round 2, GPU session 24 (1 GPU): the implicit scheme against the goldens of the reference's own compiled U_Newmark_Beta /
# U_Static (oracle/minipetsc), the Von-Mises / Hencky tangents, the Von-Mises C_ep of the explicit golden trace
cd "$(dirname "$0")/../.."
O=gpurun_out/r02_s24; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_implicit.py "tests/test_gpu_parity.py::test_steps_match_golden_reference" -m gpu -q -rf > $O/pytest.log 2>&1; echo "rc=$?"
tail -40 $O/pytest.log
