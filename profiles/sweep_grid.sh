# resident blocks of the persistent cell-block kernels: does throughput follow occupancy?
for cfg in "1 148" "1 296" "1 444" "1 592" "0 592" "0 888" "0 1184"; do
  set -- $cfg
  echo "== CACHE=$1 GRID=$2"
  NLPS_CACHE_PA=$1 NLPS_GRID=$2 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['per_kernel']; print(round(d['ms_per_step'],4), {n:k[n]['ms'] for n in ('lme_p2g_mass_disp','kin_stress_p2g_force','g2p_update')})"
done
