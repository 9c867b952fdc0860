#!/bin/bash
# round 2, GPU session 12 (2 GPUs): implicit scheme over slabs -- loopback tests, then configs[4] at N = 1 and N = 2 (NCCL)
cd "$(dirname "$0")/../.."
O=gpurun_out/r02_s12; mkdir -p $O
echo "== implicit slabs (loopback)"; timeout 900 python -m pytest tests/test_gpu_slabs.py -m gpu -q --timeout 600 -k implicit > $O/pytest_imp.log 2>&1; echo "rc=$?"; tail -5 $O/pytest_imp.log
echo "== bench c5 N=1"; timeout 900 python bench.py --workload c5 --steps 3 --warmup 1 > $O/bench_c5_n1.json 2> $O/bench_c5_n1.err; echo "rc=$?"; tail -c 1800 $O/bench_c5_n1.json; tail -3 $O/bench_c5_n1.err
echo "== bench c5 N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29781 bench.py --gpus 2 --workload c5 --steps 3 --warmup 1 > $O/bench_c5_n2.json 2> $O/bench_c5_n2.err; echo "rc=$?"; tail -c 1800 $O/bench_c5_n2.json; tail -5 $O/bench_c5_n2.err
