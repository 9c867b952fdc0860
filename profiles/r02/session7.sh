#!/bin/bash
# round 2, GPU session 7 (2 GPUs): the whole GPU test suite (NCCL slab test, C-host two-GPU driver run, 100k drop-in deck),
# both halo transports of the NCCL worker, bench lines at N = 2 with the slab parity check
cd "$(dirname "$0")/../.."
O=gpurun_out/r02_s7; mkdir -p $O
nvidia-smi -L | tee $O/gpus.txt
echo "== pytest gpu (all)"; timeout 1700 python -m pytest tests -m gpu -q --timeout 900 > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== NCCL slab worker: peer-memory halos"; timeout 600 $TR --master-port 29751 tests/workers/slab_nccl_worker.py > $O/slab_nccl_p2p.log 2>&1; echo "rc=$?"; grep -h "NCCL slabs OK\|unavailable\|Error" $O/slab_nccl_p2p.log | head -3
echo "== NCCL slab worker: ncclSend/ncclRecv halos (NLPS_P2P=0)"; NLPS_P2P=0 timeout 600 $TR --master-port 29752 tests/workers/slab_nccl_worker.py > $O/slab_nccl_sendrecv.log 2>&1; echo "rc=$?"; grep -h "NCCL slabs OK\|unavailable\|Error" $O/slab_nccl_sendrecv.log | head -3
echo "== bench c3 N=2"; ( time timeout 1500 $TR --master-port 29753 bench.py --gpus 2 > $O/bench_c3_n2.json 2> $O/bench_c3_n2.err ); tail -c 1500 $O/bench_c3_n2.json; tail -4 $O/bench_c3_n2.err
echo "== bench c2 N=2"; timeout 900 $TR --master-port 29754 bench.py --gpus 2 --workload c2 --steps 50 > $O/bench_c2_n2.json 2> $O/bench_c2_n2.err; tail -c 1200 $O/bench_c2_n2.json; tail -4 $O/bench_c2_n2.err
echo "== bench c4 N=2 (scale 0.6)"; timeout 900 $TR --master-port 29755 bench.py --gpus 2 --workload c4 --scale 0.6 --steps 20 > $O/bench_c4_n2.json 2> $O/bench_c4_n2.err; tail -c 1200 $O/bench_c4_n2.json; tail -4 $O/bench_c4_n2.err
