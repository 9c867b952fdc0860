#!/bin/bash
# round 2, GPU session 23 (1 GPU): the driver's commands with the final build -- smoke, default bench line, reference arm
cd "$(dirname "$0")/../.."
O=gpurun_out/r02_s23; mkdir -p $O
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?"; tail -2 $O/smoke.log
echo "== bench default"; ( time timeout 1200 python bench.py > $O/bench_c3_n1.json 2> $O/bench_c3_n1.err ) 2>&1 | grep real; python -c "
import json;l=json.loads(open('$O/bench_c3_n1.json').read().strip().splitlines()[-1]);r=l['roofline'];print(l['ms_per_step'],l['value'],l['e2e']['value'],l['e2e']['seconds'],l['cpu_baseline']['value'],r['frac'],r['step_frac'],r['traffic'],l['gpu_launches'],l['clocks'])"
echo "== reference arm"; timeout 600 python bench.py --impl reference > $O/bench_c3_ref.json 2> $O/bench_c3_ref.err; tail -c 300 $O/bench_c3_ref.json
