cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_4096env.csv python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/r02_launches_bench.json 2>/dev/null
wc -l gpurun_out/r02_launches_4096env.csv
