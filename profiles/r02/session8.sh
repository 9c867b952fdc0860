#!/bin/bash
# round 2, GPU session 8 (2 GPUs): N = 2 bench lines after the collective warm-up at slab create, with a per-step trace
cd "$(dirname "$0")/../.."
O=gpurun_out/r02_s8; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== bench c2 N=2"; NLPS_BENCH_TRACE=24 timeout 900 $TR --master-port 29754 bench.py --gpus 2 --workload c2 --steps 50 --no-e2e > $O/bench_c2_n2.json 2> $O/bench_c2_n2.err; python -c "
import json;l=json.loads(open('$O/bench_c2_n2.json').read().strip().splitlines()[-1]);print(l['ms_per_step'],l['value'],l['config']['setup_seconds'])"; grep trace $O/bench_c2_n2.err
echo "== bench c3 N=2"; NLPS_BENCH_TRACE=24 timeout 1500 $TR --master-port 29753 bench.py --gpus 2 > $O/bench_c3_n2.json 2> $O/bench_c3_n2.err; python -c "
import json;l=json.loads(open('$O/bench_c3_n2.json').read().strip().splitlines()[-1]);print(l['ms_per_step'],l['value'],l['config']['setup_seconds'],l['e2e'])"; grep trace $O/bench_c3_n2.err
