set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02d_tests.log
tail -60 gpurun_out/r02d_tests.log
