#!/bin/bash
# round 2, final evidence run (1 GPU): plain bench (exit 0) -> ncu launch list of the same command -> ncu --set full of the
# three hot kernels inside the same command.  A number printed by a run under ncu is never a bench value.
cd "$(dirname "$0")/../.."
O=gpurun_out/r02_final; mkdir -p $O
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e"
echo "== plain"; timeout 600 $CMD > $O/plain_c3.json 2> $O/plain_c3.err; echo "rc=$?"
if [ -s $O/plain_c3.json ]; then
  echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_c3.csv $CMD > $O/ncu_list.log 2>&1; echo "rc=$?"
  echo "== full capture"; timeout 1500 ncu --set full --clock-control none --import-source on -k regex:cw_ -s 12 -c 6 -o $O/full_c3 $CMD > $O/ncu_full.log 2>&1; echo "rc=$?"
  ncu -i $O/full_c3.ncu-rep --page raw --csv > $O/full_c3_raw.csv 2>/dev/null
  ls -la $O
fi
