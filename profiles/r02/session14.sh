#!/bin/bash
# round 2, GPU session 14 (1 GPU): GPU suite + the default bench line (exactly what the driver runs) + the reference arm
cd "$(dirname "$0")/../.."
O=gpurun_out/r02_s14; mkdir -p $O
echo "== pytest gpu"; timeout 1700 python -m pytest tests -m gpu -q --timeout 900 > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?"; tail -2 $O/smoke.log
echo "== bench default"; ( time timeout 1200 python bench.py > $O/bench_c3_n1.json 2> $O/bench_c3_n1.err ); python -c "
import json;l=json.loads(open('$O/bench_c3_n1.json').read().strip().splitlines()[-1]);print(l['ms_per_step'],l['value'],{k:v['ms'] for k,v in l['roofline']['per_kernel'].items() if v['ms']>1}, l['e2e']['value'], l['cpu_baseline'] and l['cpu_baseline']['value'], l['roofline']['traffic'])"; tail -3 $O/bench_c3_n1.err
echo "== reference arm"; ( time timeout 1200 python bench.py --impl reference > $O/bench_c3_ref.json 2> $O/bench_c3_ref.err ); tail -c 600 $O/bench_c3_ref.json
