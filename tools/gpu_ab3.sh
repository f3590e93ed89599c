cd $GRAFT_REPO_ROOT
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d=json.load(open(sys.argv[1]))
    ex=d["config"].get("extras",{})
    print(sys.argv[2], "%d envs %.4f ms %.1f M/s e2e %.1fM |" % (d["config"]["envs_per_gpu"], d["ms_per_step"], d["value"]/1e6, d["e2e"]["value"]/1e6), " | ".join("%s %.4f" % (k[:16], x.get("ms_per_step") or x.get("ms_per_control_step") or 0) for k,x in ex.items()))
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
}
timeout 600 python bench.py > gpurun_out/ab3_default.json 2> gpurun_out/ab3_default.err; show gpurun_out/ab3_default.json default
REXSIM_FORCE_BUILD=big timeout 600 python bench.py > gpurun_out/ab3_forcebig.json 2> gpurun_out/ab3_forcebig.err; show gpurun_out/ab3_forcebig.json forcebig
for v in big512 big128; do
  REXSIM_LIB=$PWD/rex_gym_b200/librexsim_$v.so timeout 300 python bench.py --no-extras --envs-per-gpu 65536 > gpurun_out/ab3_$v.json 2> gpurun_out/ab3_$v.err; show gpurun_out/ab3_$v.json $v
done
REXSIM_FORCE_BUILD=big REXSIM_LIB=$PWD/rex_gym_b200/librexsim_big512.so timeout 300 python bench.py --no-extras --envs-per-gpu 16384 > gpurun_out/ab3_b512_16k.json 2>/dev/null; show gpurun_out/ab3_b512_16k.json big512@16k
REXSIM_FORCE_BUILD=big timeout 300 python bench.py --no-extras --envs-per-gpu 16384 > gpurun_out/ab3_b256_16k.json 2>/dev/null; show gpurun_out/ab3_b256_16k.json big256@16k
timeout 300 python bench.py --no-extras --envs-per-gpu 16384 > gpurun_out/ab3_small_16k.json 2>/dev/null; show gpurun_out/ab3_small_16k.json small@16k
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02f_tests.log
tail -5 gpurun_out/r02f_tests.log
