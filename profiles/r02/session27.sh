#!/bin/bash
# round 2, GPU session 27 (1 GPU): aLME on the device -- initialisation, explicit traces, implicit steps against the goldens of
# the reference's compiled Nodes/aLME.c; with them every other test of the two files (the LME paths share the kernels)
cd "$(dirname "$0")/../.."
O=gpurun_out/r02_s27; mkdir -p $O
timeout 150 python -m pytest tests/test_gpu_parity.py tests/test_gpu_implicit.py -m gpu -q -rf > $O/pytest.log 2>&1; echo "rc=$?"
grep -E "passed|failed|^FAILED|Error|rel err" $O/pytest.log | head -40
