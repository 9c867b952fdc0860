#!/bin/bash
# round 2, GPU session 1: parity of the warp-per-cell kernels (default) + A/B against the round-1 kernels
cd "$(dirname "$0")/../.."
O=gpurun_out/r02_s1; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.log
echo "== pytest gpu (new kernels)"; timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest_new.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_new.log
echo "== 3D A/B (64^3 cube, 2.1M particles)"
for spec in "NLPS_KERNELS=1" "NLPS_KERNELS=2" "NLPS_KERNELS=2 NLPS_CW_NC=64" "NLPS_KERNELS=2 NLPS_CW_WARPS=2" "NLPS_KERNELS=2 NLPS_SPLIT_NH=1"; do
  echo "-- $spec"; env $spec timeout 300 python profiles/bench_3d.py 64 10 2>&1 | tail -1 | python -c "
import json,sys
try:
  d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), d['kernel_ms'])
except Exception as ex: print('failed', ex)"
done
echo "== 3D gamma 3 (n~100)"; for spec in "NLPS_KERNELS=1" "NLPS_KERNELS=2"; do echo "-- $spec"; env $spec timeout 300 python profiles/bench_3d.py 48 6 3.0 2>&1 | tail -1 | cut -c1-600; done
echo "== 2D A/B (C2, 1.0M particles, DP)"
for spec in "NLPS_KERNELS=1" "NLPS_KERNELS=2" "NLPS_KERNELS=2 NLPS_CW_CPW=2" "NLPS_KERNELS=2 NLPS_CW_CPW=8"; do
  echo "-- $spec"; env $spec timeout 300 python bench.py --workload c2 --steps 20 --no-cpu --no-e2e 2>$O/err.txt | tail -1 | python -c "
import json,sys
try:
  d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), {k:v['ms'] for k,v in d['roofline']['per_kernel'].items()})
except Exception as ex: print('failed', ex)"
done
echo "== C4 (MN slope) half scale"
for spec in "NLPS_KERNELS=1" "NLPS_KERNELS=2"; do
  echo "-- $spec"; env $spec timeout 400 python bench.py --workload c4 --scale 0.5 --steps 10 --no-cpu --no-e2e 2>$O/err4.txt | tail -1 | python -c "
import json,sys
try:
  d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), d['config']['particles_per_gpu'], {k:v['ms'] for k,v in d['roofline']['per_kernel'].items()})
except Exception as ex: print('failed', ex)"
done
echo "== ncu (3D new kernels)"
timeout 300 python profiles/prof_run3d.py 64 4 > $O/plain3d.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cw_ -s 3 -c 3 -o $O/prof3d python profiles/prof_run3d.py 64 4 > $O/ncu3d.log 2>&1
echo "ncu rc=$?"
ls -la $O
