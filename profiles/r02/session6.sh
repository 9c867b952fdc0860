#!/bin/bash
# round 2, GPU session 6: the driver's bench line (configs[2], 16 M particles) at N=1 + reference arm + c2 / c4 lines
cd "$(dirname "$0")/../.."
O=gpurun_out/r02_s6; mkdir -p $O
free -g | head -2; nproc
echo "== pytest (new tests only)"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "config or row_form" > $O/pytest.log 2>&1; tail -4 $O/pytest.log
echo "== bench c3 N=1 (default flags)"; ( time timeout 1500 python bench.py > $O/bench_c3_n1.json 2> $O/bench_c3_n1.err ); tail -c 3000 $O/bench_c3_n1.json; tail -5 $O/bench_c3_n1.err
echo "== reference arm c3"; ( time timeout 900 python bench.py --impl reference > $O/bench_c3_ref.json 2> $O/bench_c3_ref.err ); tail -c 1200 $O/bench_c3_ref.json; tail -3 $O/bench_c3_ref.err
echo "== bench c4 N=1"; timeout 1200 python bench.py --workload c4 --steps 20 > $O/bench_c4_n1.json 2> $O/bench_c4_n1.err; tail -c 1500 $O/bench_c4_n1.json; tail -3 $O/bench_c4_n1.err
echo "== bench c2 N=1"; timeout 900 python bench.py --workload c2 --steps 50 > $O/bench_c2_n1.json 2> $O/bench_c2_n1.err; tail -c 1500 $O/bench_c2_n1.json; tail -3 $O/bench_c2_n1.err
