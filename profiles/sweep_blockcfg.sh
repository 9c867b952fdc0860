for cfg in "32 128" "16 64" "8 32" "8 64" "16 128"; do
  set -- $cfg
  echo "== C=$1 T=$2"
  NLPS_CELLS_PER_BLOCK=$1 NLPS_THREADS=$2 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['per_kernel']; print(round(d['ms_per_step'],4), {n:k[n]['ms'] for n in ('lme_p2g_mass_disp','kin_stress_p2g_force','g2p_update')})"
done
