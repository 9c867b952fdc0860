#!/bin/bash
# round 2, GPU session 18 (1 GPU): cell-major partial sums in the 3D kernels -- GPU suite, 64^3 timing, default bench line
cd "$(dirname "$0")/../.."
O=gpurun_out/r02_s18; mkdir -p $O
echo "== pytest gpu"; timeout 1700 python -m pytest tests -m gpu -q --timeout 900 > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest.log
echo "== 64^3"; timeout 300 python profiles/bench_3d.py 64 10 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms']; print(round(d['ms_per_step'],3), {n:k[n] for n in k if k[n]>0.1})"
echo "== bench default"; timeout 1200 python bench.py --no-cpu > $O/bench_c3_n1.json 2> $O/bench_c3_n1.err; python -c "
import json;l=json.loads(open('$O/bench_c3_n1.json').read().strip().splitlines()[-1]);print(l['ms_per_step'],l['value'],{k:v['ms'] for k,v in l['roofline']['per_kernel'].items() if v['ms']>0.9}, l['e2e']['value'])"; tail -3 $O/bench_c3_n1.err
