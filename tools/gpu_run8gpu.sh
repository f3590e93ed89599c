cd $GRAFT_REPO_ROOT
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N > gpurun_out/r02p_bench_${N}gpu.json 2> gpurun_out/r02p_bench_${N}gpu.err; echo "bench rc=$?"
python - gpurun_out/r02p_bench_${N}gpu.json <<'PY'
import json, sys
s=open(sys.argv[1]).read().strip().split("\n")
print("stdout lines:", len(s))
d=json.loads(s[-1])
ex=d["config"].get("extras",{})
print("n_gpus", d["n_gpus"], "%d envs/gpu %.4f ms %.1f M/s e2e %.1fM" % (d["config"]["envs_per_gpu"], d["ms_per_step"], d["value"]/1e6, d["e2e"]["value"]/1e6))
for k,x in ex.items(): print("  ", k, x.get("envs_per_gpu"), x.get("ms_per_step") or x.get("ms_per_control_step"), "%.1f M/s" % ((x.get("value") or 0)/1e6), x.get("error"))
print(d["timing"]["per_rank_ms_per_step"])
PY
tail -3 gpurun_out/r02p_bench_${N}gpu.err
