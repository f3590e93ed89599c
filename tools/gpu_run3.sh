set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02c_tests.log
timeout 600 python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
tail -25 gpurun_out/r02c_tests.log; head -c 600 gpurun_out/r02c_bench.json
