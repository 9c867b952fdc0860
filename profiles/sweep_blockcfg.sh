# sweep cells per block / threads / particle chunk of the cell-block kernels (env overrides of the engine)
for cfg in "32 128 160" "32 128 128" "24 96 96" "16 64 64" "16 64 80" "8 32 32" "48 128 192" "64 128 256" "32 96 128" "20 96 96"; do
  set -- $cfg
  echo "== C=$1 T=$2 PCAP=$3"
  NLPS_CELLS_PER_BLOCK=$1 NLPS_THREADS=$2 NLPS_PCAP=$3 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['per_kernel']; print(round(d['ms_per_step'],4), {n:k[n]['ms'] for n in ('lme_p2g_mass_disp','kin_stress_p2g_force','g2p_update')})"
done
