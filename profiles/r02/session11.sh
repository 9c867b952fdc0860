#!/bin/bash
# round 2, GPU session 11 (1 GPU): implicit scheme over slabs (thread loopback), then the whole suite
cd "$(dirname "$0")/../.."
O=gpurun_out/r02_s11; mkdir -p $O
echo "== implicit slabs"; timeout 900 python -m pytest tests/test_gpu_slabs.py -m gpu -q --timeout 600 -k implicit > $O/pytest_imp.log 2>&1; echo "rc=$?"; tail -40 $O/pytest_imp.log
echo "== implicit single"; timeout 900 python -m pytest tests/test_gpu_implicit.py -m gpu -q --timeout 600 > $O/pytest_imp1.log 2>&1; echo "rc=$?"; tail -5 $O/pytest_imp1.log
