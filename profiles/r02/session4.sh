#!/bin/bash
# round 2, GPU session 3: v3 of the warp-per-cell kernels (compact lists in HBM, conflict-free list build, prefetch)
cd "$(dirname "$0")/../.."
O=gpurun_out/r02_s4; mkdir -p $O
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest.log
echo "== 3D A/B (64^3 cube, 2.1M particles)"
for spec in "NLPS_KERNELS=1" "NLPS_KERNELS=2" "NLPS_KERNELS=2 NLPS_CW_GRID=592" "NLPS_KERNELS=2 NLPS_CW_GRID=2368"; do
  echo "-- $spec"; env $spec timeout 300 python profiles/bench_3d.py 64 10 2>&1 | tail -1 | python -c "
import json,sys
try:
  d=json.loads(sys.stdin.read()); k=d['kernel_ms']; print(round(d['ms_per_step'],3), {n:k[n] for n in k if k[n]>0.15})
except Exception as ex: print('failed', ex)"
done
echo "== 3D gamma 3 (n~100)"; for spec in "NLPS_KERNELS=1" "NLPS_KERNELS=2"; do echo "-- $spec"; env $spec timeout 300 python profiles/bench_3d.py 48 6 3.0 2>&1 | tail -1 | python -c "
import json,sys
try:
  d=json.loads(sys.stdin.read()); k=d['kernel_ms']; print(round(d['ms_per_step'],3), {n:k[n] for n in k if k[n]>0.15})
except Exception as ex: print('failed', ex)"; done
echo "== C4 (MN slope) half scale"
for spec in "NLPS_KERNELS=1" "NLPS_KERNELS=2"; do
  echo "-- $spec"; env $spec timeout 400 python bench.py --workload c4 --scale 0.5 --steps 10 --no-cpu --no-e2e 2>$O/err4.txt | tail -1 | python -c "
import json,sys
try:
  d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), d['config']['particles_per_gpu'], {k:v['ms'] for k,v in d['roofline']['per_kernel'].items() if v['ms']>0.1})
except Exception as ex: print('failed', ex)"
done
echo "== ncu (3D new kernels)"
timeout 300 python profiles/prof_run3d.py 64 4 > $O/plain3d.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cw_ -s 3 -c 3 -o $O/prof3d python profiles/prof_run3d.py 64 4 > $O/ncu3d.log 2>&1
echo "ncu rc=$?"
