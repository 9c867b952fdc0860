for cfg in "1 32 128 128" "0 32 128 128" "0 16 64 64" "0 64 128 256" "0 24 96 96"; do
  set -- $cfg
  echo "== CACHE=$1 C=$2 T=$3 PCAP=$4"
  NLPS_CACHE_PA=$1 NLPS_CELLS_PER_BLOCK=$2 NLPS_THREADS=$3 NLPS_PCAP=$4 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['per_kernel']; print(round(d['ms_per_step'],4), {n:k[n]['ms'] for n in ('lme_p2g_mass_disp','kin_stress_p2g_force','g2p_update')})"
done
