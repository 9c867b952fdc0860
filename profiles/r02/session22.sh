#!/bin/bash
# round 2, GPU session 22 (8 GPUs): final build -- c3, c4 and c5 at N = 8
cd "$(dirname "$0")/../.."
O=gpurun_out/r02_s22; mkdir -p $O
run() { # name nproc args...
  local nm=$1 n=$2; shift 2
  echo "== $nm"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29800 + RANDOM % 100)) bench.py --gpus $n "$@" > $O/$nm.json 2> $O/$nm.err
  echo "rc=$?"
  python - <<PY
import json
try:
    l=json.loads(open('$O/$nm.json').read().strip().splitlines()[-1])
    print('ms/step', l['ms_per_step'], 'value %.4g' % l['value'], 'setup', l['config']['setup_seconds'], 'parity', (l.get('parity') or {}).get('parity_n'), 'e2e', (l.get('e2e') or {}).get('seconds'))
except Exception as e: print('no line', e)
PY
  tail -2 $O/$nm.err | cut -c1-300
}
run bench_c3_n8 8
run bench_c4_n8 8 --workload c4 --steps 20
run bench_c5_n8 8 --workload c5 --steps 3 --warmup 1
