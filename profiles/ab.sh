# A/B two builds of the library on the same box: ab.sh <libA.so> <libB.so> [steps]
STEPS=${3:-40}
for rep in 1 2; do
for lib in "$1" "$2"; do
  echo "== $lib"
  NLPS_LIB=$lib python bench.py --steps $STEPS --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['per_kernel']; print(round(d['ms_per_step'],4), {n:k[n]['ms'] for n in ('lme_p2g_mass_disp','kin_stress_p2g_force','g2p_update')}, d['clocks'])"
done
done
