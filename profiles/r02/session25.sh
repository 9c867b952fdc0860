#!/bin/bash
# round 2, GPU session 25 (1 GPU): the whole GPU suite and smoke() with the final build (Von-Mises / Hencky tangents, goldens of
# the reference's compiled implicit schemes, loads and mixed-material fixtures)
cd "$(dirname "$0")/../.."
O=gpurun_out/r02_s25; mkdir -p $O
timeout 280 python -m pytest tests -m gpu -q -rfs > $O/pytest.log 2>&1; echo "pytest rc=$?"
tail -15 $O/pytest.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
