#!/bin/bash
# round 2, GPU session 21 (2 GPUs): final build -- whole GPU suite (2-GPU tests included), NCCL slab worker on both halo
# transports, c3 / c5 at N = 2, c4 and c2 at N = 1
cd "$(dirname "$0")/../.."
O=gpurun_out/r02_s21; mkdir -p $O
echo "== pytest gpu (all)"; timeout 1700 python -m pytest tests -m gpu -q --timeout 900 > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== NCCL slab worker: peer memory"; timeout 600 $TR --master-port 29751 tests/workers/slab_nccl_worker.py > $O/slab_nccl_p2p.log 2>&1; echo "rc=$?"; grep -h "NCCL slabs OK\|unavailable\|Error" $O/slab_nccl_p2p.log | head -3
echo "== NCCL slab worker: ncclSend/ncclRecv"; NLPS_P2P=0 timeout 600 $TR --master-port 29752 tests/workers/slab_nccl_worker.py > $O/slab_nccl_sendrecv.log 2>&1; echo "rc=$?"; grep -h "NCCL slabs OK\|unavailable\|Error" $O/slab_nccl_sendrecv.log | head -3
show() { python -c "
import json;l=json.loads(open('$1').read().strip().splitlines()[-1]);print(l['ms_per_step'],'%.4g'%l['value'],(l.get('parity') or {}).get('parity_n'),{k:v['ms'] for k,v in l['roofline'].get('per_kernel',{}).items() if v['ms']>0.25})"; }
echo "== c3 N=2"; timeout 900 $TR --master-port 29753 bench.py --gpus 2 > $O/bench_c3_n2.json 2> $O/bench_c3_n2.err; show $O/bench_c3_n2.json
echo "== c5 N=2"; timeout 900 $TR --master-port 29754 bench.py --gpus 2 --workload c5 --steps 3 --warmup 1 > $O/bench_c5_n2.json 2> $O/bench_c5_n2.err; show $O/bench_c5_n2.json
echo "== c4 N=1"; timeout 900 python bench.py --workload c4 --no-cpu > $O/bench_c4_n1.json 2> $O/bench_c4_n1.err; show $O/bench_c4_n1.json
echo "== c2 N=1"; timeout 900 python bench.py --workload c2 --steps 50 --no-cpu > $O/bench_c2_n1.json 2> $O/bench_c2_n1.err; show $O/bench_c2_n1.json
echo "== c5 N=1"; timeout 900 python bench.py --workload c5 --steps 3 --warmup 1 > $O/bench_c5_n1.json 2> $O/bench_c5_n1.err; show $O/bench_c5_n1.json
