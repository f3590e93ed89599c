# A/B of library variants (REXSIM_LIB): full bench line per variant
cd $GRAFT_REPO_ROOT
for v in "" _old _occ3 _repl; do
  if [ -f rex_gym_b200/librexsim$v.so ]; then
    REXSIM_LIB=$PWD/rex_gym_b200/librexsim$v.so timeout 600 python bench.py > gpurun_out/ab$v.json 2> gpurun_out/ab$v.err
    python - <<PY
import json
d=json.load(open("gpurun_out/ab$v.json"))
print("variant '$v': 4096 %.4f ms  e2e %.1fM" % (d["ms_per_step"], d["e2e"]["value"]/1e6), " | ".join("%s %.4f" % (k[:14], x.get("ms_per_step") or x.get("ms_per_control_step") or 0) for k,x in d["config"]["extras"].items()))
PY
  fi
done
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02e_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02e_tests.log
tail -15 gpurun_out/r02e_tests.log
