cd $GRAFT_REPO_ROOT
T=$1
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${T}_ref.json 2>/dev/null
python tools/bench_agent.py > gpurun_out/${T}_agent.txt 2>&1
python - gpurun_out/${T}_bench.json <<'PY'
import json, sys
d=json.load(open(sys.argv[1]))
ex=d["config"].get("extras",{})
print("%d envs %.4f ms %.1f M/s e2e %.1fM |" % (d["config"]["envs_per_gpu"], d["ms_per_step"], d["value"]/1e6, d["e2e"]["value"]/1e6), " | ".join("%s %.4f" % (k[:16], x.get("ms_per_step") or x.get("ms_per_control_step") or 0) for k,x in ex.items()))
print({k:d[k] for k in d if k in ("roofline","issue","fp32","cpu_baseline","clocks","gpu_launches")})
PY
tail -3 gpurun_out/${T}_ref.json | cut -c1-400; tail -15 gpurun_out/${T}_agent.txt
