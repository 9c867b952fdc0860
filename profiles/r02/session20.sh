#!/bin/bash
# round 2, GPU session 20 (1 GPU): cell-major cell sums in the 2D (block-per-cell-group) kernels -- GPU suite, C2 A/B
cd "$(dirname "$0")/../.."
O=gpurun_out/r02_s20; mkdir -p $O
echo "== pytest gpu"; timeout 1700 python -m pytest tests -m gpu -q --timeout 900 > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.log
for spec in "NLPS_PART_SLOTMAJOR=1" "NLPS_PART_SLOTMAJOR=0"; do
  echo "-- c2 $spec"; env $spec timeout 600 python bench.py --workload c2 --steps 50 --no-cpu --no-e2e > $O/bench_c2_$spec.json 2> $O/err.txt; python -c "
import json;l=json.loads(open('$O/bench_c2_$spec.json').read().strip().splitlines()[-1]);print(l['ms_per_step'],l['value'],{k:v['ms'] for k,v in l['roofline']['per_kernel'].items() if v['ms']>0.02})"
done
