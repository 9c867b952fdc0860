#!/bin/bash
# round 2, GPU session 13 (1 GPU): whole GPU suite after Lade-Duncan and the implicit slabs
cd "$(dirname "$0")/../.."
O=gpurun_out/r02_s13; mkdir -p $O
echo "== pytest gpu"; timeout 1700 python -m pytest tests -m gpu -q --timeout 900 > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest.log
