#!/bin/bash
# round 2, GPU session 19 (1 GPU): quick A/B of a kernel change -- 3D parity subset, 64^3 timing, default bench (device-timed only)
cd "$(dirname "$0")/../.."
O=gpurun_out/r02_s19; mkdir -p $O
echo "== parity subset"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slabs.py -m gpu -q --timeout 600 -k "3d or cube or synthetic or implicit or slope" > $O/pytest.log 2>&1; echo "rc=$?"; tail -3 $O/pytest.log
echo "== bench default"; timeout 1200 python bench.py --no-cpu --no-e2e > $O/bench_c3_n1.json 2> $O/bench_c3_n1.err; python -c "
import json;l=json.loads(open('$O/bench_c3_n1.json').read().strip().splitlines()[-1]);print(l['ms_per_step'],l['value'],{k:v['ms'] for k,v in l['roofline']['per_kernel'].items() if v['ms']>0.9})"; tail -3 $O/bench_c3_n1.err
