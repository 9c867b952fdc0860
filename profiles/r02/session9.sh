#!/bin/bash
# round 2, GPU session 9 (8 GPUs): scaling lines at N = 8 (c3, c2, c4) and N = 4 (c3), with the set-up marks of the engine
cd "$(dirname "$0")/../.."
O=gpurun_out/r02_s9; mkdir -p $O
nvidia-smi -L > $O/gpus.txt; nvidia-smi topo -m > $O/topo.txt 2>&1
run() { # name nproc args...
  local nm=$1 n=$2; shift 2
  echo "== $nm"
  NLPS_TIMING=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29800 + RANDOM % 100)) bench.py --gpus $n "$@" > $O/$nm.json 2> $O/$nm.err
  echo "rc=$?"
  python - <<PY
import json
try:
    l=json.loads(open('$O/$nm.json').read().strip().splitlines()[-1])
    print('ms/step', l['ms_per_step'], 'value %.4g' % l['value'], 'setup', l['config']['setup_seconds'], 'parity', (l.get('parity') or {}).get('parity_n'), 'e2e', (l.get('e2e') or {}).get('seconds'), 'launches', l['gpu_launches'])
    print({k:(v['ms'],v['launches_per_step']) for k,v in l['roofline']['per_kernel'].items()})
except Exception as e: print('no line', e)
PY
  tail -3 $O/$nm.err
}
run bench_c3_n8 8
run bench_c2_n8 8 --workload c2 --steps 50
run bench_c4_n8 8 --workload c4 --steps 20
run bench_c3_n4 4
