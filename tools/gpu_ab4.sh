cd $GRAFT_REPO_ROOT
for v in "" _b192; do
  for n in 65536 16384; do
    REXSIM_LIB=$PWD/rex_gym_b200/librexsim$v.so timeout 300 python bench.py --no-extras --envs-per-gpu $n > gpurun_out/ab4$v.$n.json 2> gpurun_out/ab4$v.$n.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab4$v.$n.json"))
    print("variant '$v' n=$n: %.4f ms  %.1f M/s err=%s" % (d["ms_per_step"], d["value"]/1e6, d["config"].get("error_flags_or")))
except Exception as e:
    print("variant '$v' n=$n failed", e)
PY
  done
done
