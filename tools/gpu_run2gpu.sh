cd $GRAFT_REPO_ROOT
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r02n_bench_2gpu.json 2> gpurun_out/r02n_bench_2gpu.err; echo "bench rc=$?"
python - gpurun_out/r02n_bench_2gpu.json <<'PY'
import json, sys
d=json.load(open(sys.argv[1]))
ex=d["config"].get("extras",{})
print("n_gpus", d["n_gpus"], "%d envs/gpu %.4f ms %.1f M/s e2e %.1fM |" % (d["config"]["envs_per_gpu"], d["ms_per_step"], d["value"]/1e6, d["e2e"]["value"]/1e6), " | ".join("%s %.4f" % (k[:24], x.get("ms_per_step") or x.get("ms_per_control_step") or 0) for k,x in ex.items()))
print(d["timing"])
PY
tail -3 gpurun_out/r02n_bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 10 --warmup 3 2>/dev/null | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_multirank.py -q 2>&1 | tail -3
