# A/B of 3D builds / settings on one box: ab3d.sh "<ENV=.. lib>" ...   each argument: "VAR=val,VAR2=val2:lib.so"
for spec in "$@"; do
  envs=${spec%%:*}; lib=${spec##*:}
  echo "== $spec"
  env $(echo $envs | tr ',' ' ') NLPS_LIB=$lib python profiles/bench_3d.py 64 10 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernel_ms']; print(round(d['ms_per_step'],3), {n:k[n] for n in ('lme_p2g_mass_disp','kin_stress_p2g_force','g2p_update')})"
done
