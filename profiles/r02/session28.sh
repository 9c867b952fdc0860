#!/bin/bash
# round 2, GPU session 28 (1 GPU): the whole GPU suite and smoke() with the final build (aLME in the 2D kernels)
cd "$(dirname "$0")/../.."
O=gpurun_out/r02_s28; mkdir -p $O
timeout 200 python -m pytest tests -m gpu -q -rfs > $O/pytest.log 2>&1; echo "pytest rc=$?"
tail -8 $O/pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
