cd $GRAFT_REPO_ROOT
T=$1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
tail -12 gpurun_out/${T}_tests.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - gpurun_out/${T}_bench.json <<'PY'
import json, sys
d=json.load(open(sys.argv[1]))
ex=d["config"].get("extras",{})
print("%d envs %.4f ms %.1f M/s e2e %.1fM |" % (d["config"]["envs_per_gpu"], d["ms_per_step"], d["value"]/1e6, d["e2e"]["value"]/1e6), " | ".join("%s %.4f" % (k[:16], x.get("ms_per_step") or x.get("ms_per_control_step") or 0) for k,x in ex.items()))
PY
