#!/bin/bash
# round 2, GPU session 15 (2 GPUs): band-only second pass of the node kernels -- slab tests, then c2 / c3 at N = 2
cd "$(dirname "$0")/../.."
O=gpurun_out/r02_s15; mkdir -p $O
echo "== slab tests"; timeout 1500 python -m pytest tests/test_gpu_slabs.py tests/test_dropin_driver.py -m gpu -q --timeout 900 > $O/pytest.log 2>&1; echo "rc=$?"; tail -5 $O/pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for w in c2 c3; do
  extra=""; [ $w = c2 ] && extra="--steps 50"
  echo "== bench $w N=2"; timeout 900 $TR --master-port $((29700 + RANDOM % 200)) bench.py --gpus 2 --workload $w $extra --no-e2e > $O/bench_${w}_n2.json 2> $O/bench_${w}_n2.err; python -c "
import json;l=json.loads(open('$O/bench_${w}_n2.json').read().strip().splitlines()[-1]);print(l['ms_per_step'],l['value'],l['parity']['parity_n'],{k:(v['ms'],v['launches_per_step']) for k,v in l['roofline']['per_kernel'].items() if 'grid' in k or 'halo' in k})"; tail -2 $O/bench_${w}_n2.err
done
