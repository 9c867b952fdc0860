#!/bin/bash
# round 2, GPU session 29 (1 GPU): the drop-in driver tests with the binary rebuilt against the final ABI header (session 28 ran
# them with a stale binary: the shim objects predated nlps_solver.shape_function)
cd "$(dirname "$0")/../.."
O=gpurun_out/r02_s29; mkdir -p $O
timeout 60 python -m pytest tests/test_dropin_driver.py -m gpu -q -rfs > $O/pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^SKIPPED" $O/pytest.log | head
