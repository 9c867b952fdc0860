#!/bin/bash
# round 2, GPU session 30 (1 GPU, the last 30 s of the round's budget): the aLME deck through the reference's own driver with the
# B200 U_Verlet shim, after the shims' "LME only" guard was relaxed to aLME in 2D
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r02_s30
timeout 20 python -m pytest "tests/test_dropin_driver.py::test_reference_driver_with_b200_scheme[almenh]" -m gpu -q -rf > gpurun_out/r02_s30/pytest.log 2>&1; echo rc=$?
