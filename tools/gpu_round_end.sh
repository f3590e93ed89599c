cd $GRAFT_REPO_ROOT
T=$1
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
tail -6 gpurun_out/${T}_tests.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - gpurun_out/${T}_bench.json <<'PY'
import json, sys
d=json.load(open(sys.argv[1]))
ex=d["config"].get("extras",{})
print("%d envs %.4f ms %.1f M/s e2e %.1fM |" % (d["config"]["envs_per_gpu"], d["ms_per_step"], d["value"]/1e6, d["e2e"]["value"]/1e6), " | ".join("%s %.4f" % (k[:24], x.get("ms_per_step") or x.get("ms_per_control_step") or 0) for k,x in ex.items()))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:perform_tc_kernel -s 60 -c 1 -f -o gpurun_out/${T}_perform_tc python tools/dev_perform_tc.py > gpurun_out/${T}_tc.txt 2>&1
python tools/bench_agent.py > gpurun_out/${T}_agent.txt 2>&1; tail -4 gpurun_out/${T}_agent.txt
